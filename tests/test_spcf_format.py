"""SPCF writer/reader twin (oracle/spcf.py) against a byte-level restatement of the layout in
src/FeatureExtractor.cpp:269-381 (CPU only)."""
import struct

import numpy as np

from oracle import spcf


def test_spcf_layout_bytes(tmp_path):
    d = np.arange(2 * 256, dtype=np.float32).reshape(2, 256)
    kps = np.array([[1.5, 2.5, 8, -1, 0.9, 0, -1], [3, 4, 8, -1, 0.8, 0, -1]], np.float32)
    p = str(tmp_path / "c.bin")
    spcf.write(p, {7: (kps, d)})
    raw = open(p, "rb").read()
    assert struct.unpack("<III", raw[:12]) == (0x53504346, 1, 1)              # magic, version, entries (:281)
    assert struct.unpack("<ii", raw[12:20]) == (7, 2)                         # frame_idx, num_kp
    assert struct.unpack("<fffffii", raw[20:48]) == (1.5, 2.5, 8.0, -1.0, np.float32(0.9), 0, -1)
    off = 20 + 2 * 28
    assert struct.unpack("<iii", raw[off:off + 12]) == (2, 256, 5)            # rows, cols, CV_32F
    assert raw[off + 12:] == d.tobytes()
    back = spcf.read(p)
    assert np.array_equal(back[7][1], d) and len(back[7][0]) == 2
