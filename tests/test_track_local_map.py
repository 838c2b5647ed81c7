"""vsm_track_local_map against Slam::track_local_map (src/Slam.cpp:380-469).
Pin: tests/golden/track_local_map_s*.npz hold the answers of the loop-for-loop restatement run with
OpenCV's own cv::norm (cv2.norm(a, b, NORM_L2), the call at :451) for every candidate distance
(oracle/make_golden.py).  Bar: identical chosen keypoints, assignments, observation order and tracked
count -- every DECISION equals what cv::norm gives; the distances themselves agree to 1e-12 relative
(cv::norm's fp64 summation order is dispatch-dependent, see include/vsm.h)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import cases, oracle
import vsm_b200


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_track_local_map_matches_reference_loop(seed):
    kp, desc, pos, mp_desc, valid, R, t = cases.track_scene(seed)
    g = np.load(os.path.join(GOLDEN, f"track_local_map_s{seed}.npz"))
    with vsm_b200.Matcher() as m:
        ind_g = -np.ones(len(kp), np.int32)
        ind_o = ind_g.copy()
        ind_g[7] = ind_o[7] = 123456                      # an index set earlier in the frame's life survives unless replaced
        tr_g, obs_g, bk_g, bd_g = m.track_local_map(kp, desc, pos, mp_desc, valid, R, t, ind_g)
        tr_o, obs_o, bk_o, bd_o = oracle.track_local_map(kp, desc, pos, mp_desc, valid, R, t, ind_o)
        assert tr_o > 300
        assert np.array_equal(bk_g, bk_o)
        assert np.allclose(bd_g, bd_o, rtol=1e-12, atol=0)
        assert tr_g == tr_o and obs_g == obs_o
        assert np.array_equal(ind_g, ind_o)
        # ... and against the cv::norm-based golden: decisions identical, distances to 1e-12
        assert np.array_equal(bk_g, g["best_ki"]) and tr_g == int(g["tracked"])
        assert obs_g == [tuple(x) for x in g["obs"].tolist()] and np.array_equal(ind_g, g["indices"])
        assert np.allclose(bd_g, g["best_dist"], rtol=1e-12, atol=0)
        # map-point descriptors resident in the device store instead of passed from the host
        m.add_keyframe(0, mp_desc)
        ind_s = -np.ones(len(kp), np.int32)
        ind_s[7] = 123456
        tr_s, obs_s, bk_s, _ = m.track_local_map(kp, desc, pos, None, valid, R, t, ind_s)
        assert tr_s == tr_o and obs_s == obs_o and np.array_equal(ind_s, ind_o)
        # ... and in the resident map-point table, with the table's own validity flags
        m.clear_store()
        m.add_map_points(mp_desc, 0)
        m.set_map_points_valid(np.nonzero(valid == 0)[0].astype(np.int32), False)
        ind_t = -np.ones(len(kp), np.int32)
        ind_t[7] = 123456
        tr_t, obs_t, bk_t, _ = m.track_local_map(kp, desc, pos, None, None, R, t, ind_t)
        assert tr_t == tr_o and obs_t == obs_o and np.array_equal(ind_t, ind_o) and np.array_equal(bk_t, bk_o)
        # nothing visible / nothing to match
        tr_e, obs_e, _, _ = m.track_local_map(kp, desc, pos + 1000.0, mp_desc, valid, R, t, -np.ones(len(kp), np.int32))
        assert tr_e == 0 and obs_e == []
        tr_z, _, _, _ = m.track_local_map(np.zeros((0, 2), np.float32), np.zeros((0, 256), np.float32), pos, mp_desc,
                                          valid, R, t, np.zeros(0, np.int32))
        assert tr_z == 0


def test_oracle_track_local_map_sanity():
    """CPU: the restated loop tracks the planted correspondences and respects validity."""
    kp, desc, pos, mp_desc, valid, R, t = cases.track_scene(5, nmp=600, nkp=200, dup=False)
    ind = -np.ones(len(kp), np.int32)
    tracked, obs, bk, bd = oracle.track_local_map(kp, desc, pos, mp_desc, valid, R, t, ind)
    assert tracked > 60 and tracked == len(obs)
    assert all(valid[mp] for mp, _ in obs)
    assert np.all(bd[bk >= 0] < 0.5)
    for mp, ki in obs:
        assert bk[mp] == ki


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_track_local_map_equals_cv_norm_golden(seed):
    """CPU: the numpy restatement takes the same decisions as the run with cv2.norm (the golden)."""
    kp, desc, pos, mp_desc, valid, R, t = cases.track_scene(seed)
    g = np.load(os.path.join(GOLDEN, f"track_local_map_s{seed}.npz"))
    ind = -np.ones(len(kp), np.int32)
    ind[7] = 123456
    tracked, obs, bk, bd = oracle.track_local_map(kp, desc, pos, mp_desc, valid, R, t, ind)
    assert tracked == int(g["tracked"]) and np.array_equal(bk, g["best_ki"])
    assert obs == [tuple(x) for x in g["obs"].tolist()] and np.array_equal(ind, g["indices"])
    assert np.allclose(bd, g["best_dist"], rtol=1e-12, atol=0)
