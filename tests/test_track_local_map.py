"""vsm_track_local_map against the loop-for-loop restatement of Slam::track_local_map
(src/Slam.cpp:380-469).  Bar: identical chosen keypoints, assignments, observation order and
tracked count; distances equal to 1e-12 relative (cv::norm's fp64 summation order is
dispatch-dependent, see include/vsm.h)."""
import numpy as np
import pytest

from oracle import gen, oracle
import vsm_b200


def scene(seed, nmp=3000, nkp=800, drop=0.3, dup=True):
    rng = np.random.default_rng(seed)
    # camera pose: small rotation about y + translation (world -> camera)
    a = 0.05
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([0.1, -0.05, 0.2])
    pos = np.stack([rng.uniform(-6, 6, nmp), rng.uniform(-4, 4, nmp), rng.uniform(-1, 12, nmp)], axis=1)
    mp_desc = gen.rows(seed, 0, 0, nmp)
    valid = (rng.random(nmp) > 0.1).astype(np.uint8)
    cam = (R @ pos.T).T + t
    z = cam[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = 525.0 * cam[:, 0] / z + 319.5
        v = 525.0 * cam[:, 1] / z + 239.5
    vis = np.nonzero((z > 0.2) & (u >= 0) & (u < 640) & (v >= 0) & (v < 480))[0]
    pick = rng.permutation(vis)[:int(nkp * (1 - drop))]
    kp = np.zeros((nkp, 2), np.float32)
    desc = gen.rows(seed, 1, 0, nkp).copy()
    k = len(pick)
    kp[:k, 0] = (u[pick] + rng.normal(0, 3.0, k)).astype(np.float32)
    kp[:k, 1] = (v[pick] + rng.normal(0, 3.0, k)).astype(np.float32)
    noisy = mp_desc[pick] + 0.02 * rng.standard_normal((k, 256)).astype(np.float32)
    desc[:k] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
    kp[k:, 0] = rng.uniform(0, 640, nkp - k)
    kp[k:, 1] = rng.uniform(0, 480, nkp - k)
    if dup and k > 40:
        # two map points at the same place with the same descriptor: the later one must NOT replace
        pos[pick[1]] = pos[pick[0]]
        mp_desc[pick[1]] = mp_desc[pick[0]]
        # a keypoint duplicated in a neighbouring cell position: visiting order decides the tie
        kp[k] = kp[5] + np.float32(0.25)
        desc[k] = desc[5]
    kp = np.clip(kp, 0, [639.5, 479.5]).astype(np.float32)
    return kp, desc, pos, mp_desc, valid, R, t


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_track_local_map_matches_reference_loop(seed):
    kp, desc, pos, mp_desc, valid, R, t = scene(seed)
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
        # map-point descriptors resident in the device store instead of passed from the host
        m.add_keyframe(0, mp_desc)
        ind_s = -np.ones(len(kp), np.int32)
        ind_s[7] = 123456
        tr_s, obs_s, bk_s, _ = m.track_local_map(kp, desc, pos, None, valid, R, t, ind_s)
        assert tr_s == tr_o and obs_s == obs_o and np.array_equal(ind_s, ind_o)
        # nothing visible / nothing to match
        tr_e, obs_e, _, _ = m.track_local_map(kp, desc, pos + 1000.0, mp_desc, valid, R, t, -np.ones(len(kp), np.int32))
        assert tr_e == 0 and obs_e == []
        tr_z, _, _, _ = m.track_local_map(np.zeros((0, 2), np.float32), np.zeros((0, 256), np.float32), pos, mp_desc,
                                          valid, R, t, np.zeros(0, np.int32))
        assert tr_z == 0


def test_oracle_track_local_map_sanity():
    """CPU: the restated loop tracks the planted correspondences and respects validity."""
    kp, desc, pos, mp_desc, valid, R, t = scene(5, nmp=600, nkp=200, dup=False)
    ind = -np.ones(len(kp), np.int32)
    tracked, obs, bk, bd = oracle.track_local_map(kp, desc, pos, mp_desc, valid, R, t, ind)
    assert tracked > 60 and tracked == len(obs)
    assert all(valid[mp] for mp, _ in obs)
    assert np.all(bd[bk >= 0] < 0.5)
    for mp, ki in obs:
        assert bk[mp] == ki
