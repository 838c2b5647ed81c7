"""Writer / reader of the reference's SPCF feature-cache format (TEST INFRASTRUCTURE ONLY).

Byte layout restated from FeatureExtractor::save_cache / load_cache
(src/FeatureExtractor.cpp:269-381): u32 magic 0x53504346, u32 version 1, u32 num_entries; per entry
i32 frame_idx, i32 num_kp, num_kp x {f32 x, y, size, angle, response, i32 octave, class_id},
i32 rows, cols, type (OpenCV type code: CV_32F = 5, CV_8U = 0), raw row-major descriptor bytes.
Entries are written in ascending frame_idx (save_cache sorts the map keys, :343-346)."""
import struct

import numpy as np

MAGIC, VERSION = 0x53504346, 1
CV_8U, CV_32F = 0, 5


def write(path, entries):
    """entries: {frame_idx: (keypoints [n,7] float32-ish or None, descriptors ndarray)}"""
    with open(path, "wb") as f:
        f.write(struct.pack("<III", MAGIC, VERSION, len(entries)))
        for idx in sorted(entries):
            kps, desc = entries[idx]
            n = 0 if kps is None else len(kps)
            f.write(struct.pack("<ii", idx, n))
            for k in range(n):
                x, y, size, angle, resp, octave, cid = kps[k]
                f.write(struct.pack("<fffffii", x, y, size, angle, resp, int(octave), int(cid)))
            desc = np.ascontiguousarray(desc)
            rows, cols = (desc.shape if desc.ndim == 2 else (0, 0))
            typ = CV_32F if desc.dtype == np.float32 else CV_8U
            f.write(struct.pack("<iii", rows, cols, typ))
            if rows > 0 and cols > 0:
                f.write(desc.tobytes())


def read(path):
    out = {}
    with open(path, "rb") as f:
        magic, version, n = struct.unpack("<III", f.read(12))
        assert magic == MAGIC and version == VERSION
        for _ in range(n):
            idx, nkp = struct.unpack("<ii", f.read(8))
            kps = np.frombuffer(f.read(28 * nkp), dtype=np.dtype("<f4,<f4,<f4,<f4,<f4,<i4,<i4"))
            rows, cols, typ = struct.unpack("<iii", f.read(12))
            dt = np.float32 if typ == CV_32F else np.uint8
            desc = np.frombuffer(f.read(rows * cols * np.dtype(dt).itemsize), dtype=dt).reshape(rows, cols) if rows > 0 and cols > 0 else np.zeros((0, 0), dt)
            out[idx] = (kps, desc)
    return out
