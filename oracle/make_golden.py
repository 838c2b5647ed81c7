"""Write cv2's answers for the named parity cases to tests/golden/ (run HERE; cv2 needed).

The golden files hold what OpenCV 4.13.0 -- the library the reference calls at
src/Slam.cpp:1149 and src/LoopCloser.cpp:51 -- returns, followed by the reference's
own filter loops restated in Python:
  top-2 (idx, dist) from cv2.BFMatcher(NORM_L2).knnMatch(q, t, 2)
  good lists at ratio 0.70 / 0.75 / 0.80     (src/Slam.cpp:1151-1158)
  mutual-NN lists                             (knnMatch(t, q, 1) composition)
  per-keyframe survivor counts                (src/LoopCloser.cpp:54-62)
  Slam::track_local_map decisions with cv2.norm as the distance (src/Slam.cpp:380-469, :451)
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases  # noqa: E402
from oracle.check_vs_cv2 import cv2_knn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def filter_loop(idx, dist, ratio, back=None):
    """src/Slam.cpp:1151-1158 on cv2's knn lists; returns queryIdx of survivors."""
    good = []
    r = np.float32(ratio)
    for i in range(idx.shape[0]):
        if idx[i, 1] < 0:
            continue
        if dist[i, 0] < np.float32(r * dist[i, 1]):
            if back is not None and back[idx[i, 0]] != i:
                continue
            good.append(i)
    return np.array(good, np.int32)


def main():
    import cv2
    os.makedirs(OUT, exist_ok=True)
    for name, fn in cases.PAIR_CASES.items():
        q, t = fn()
        idx, dist = cv2_knn(q, t, 2)
        back, _ = cv2_knn(t, q, 1)
        d = {"idx": idx.astype(np.int32), "dist": dist, "back": back[:, 0].astype(np.int32),
             "nq": q.shape[0], "nt": t.shape[0], "cv2": cv2.__version__}
        for r in cases.RATIOS:
            d[f"good_{int(r * 100)}"] = filter_loop(idx, dist, r)
            d[f"mutual_{int(r * 100)}"] = filter_loop(idx, dist, r, back[:, 0])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, q.shape, t.shape, {k: len(v) for k, v in d.items() if k.startswith(("good", "mutual"))})
    q, db, seg_off = cases.db_case()
    gidx, gdist = cv2_knn(q, db, 2)
    counts = []
    seg_idx = -np.ones((len(seg_off) - 1, q.shape[0], 2), np.int32)
    seg_dist = np.full((len(seg_off) - 1, q.shape[0], 2), np.finfo(np.float32).max, np.float32)
    for s in range(len(seg_off) - 1):
        kf = db[seg_off[s]:seg_off[s + 1]]
        i2, d2 = cv2_knn(q, kf, 2)
        seg_idx[s], seg_dist[s] = i2, d2
        counts.append([len(filter_loop(i2, d2, r)) for r in cases.RATIOS])
    np.savez_compressed(os.path.join(OUT, "db_small.npz"), gidx=gidx.astype(np.int32), gdist=gdist,
                        seg_off=seg_off, seg_idx=seg_idx, seg_dist=seg_dist,
                        counts=np.array(counts, np.int32), cv2=cv2.__version__)
    print("db_small", q.shape, db.shape, "counts@0.75 max", np.array(counts)[:, 1].max())
    # Slam::track_local_map (src/Slam.cpp:380-469) with OpenCV's own cv::norm at :451
    from oracle import oracle
    for seed in (1, 2, 3):
        kp, desc, pos, mp_desc, valid, R, t = cases.track_scene(seed)
        ind = -np.ones(len(kp), np.int32)
        ind[7] = 123456
        norm = lambda a, b: cv2.norm(np.ascontiguousarray(a).reshape(1, -1), np.ascontiguousarray(b).reshape(1, -1), cv2.NORM_L2)
        tracked, obs, bk, bd = oracle.track_local_map(kp, desc, pos, mp_desc, valid, R, t, ind, norm_fn=norm)
        np.savez_compressed(os.path.join(OUT, f"track_local_map_s{seed}.npz"), tracked=tracked,
                            obs=np.array(obs, np.int32).reshape(-1, 2), best_ki=bk, best_dist=bd, indices=ind,
                            cv2=cv2.__version__)
        print("track_local_map", seed, "tracked", tracked)


if __name__ == "__main__":
    main()
