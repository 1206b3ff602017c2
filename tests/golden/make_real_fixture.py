"""Generates tests/golden/real_depth_640x480.npz from the one real Kinect depth frame that ships with the
reference as DATA (g2o_frontend/PlaneEx_gui/test_images/image.pgm: P5, 640x480, 16-bit big-endian millimetres,
79.8 % valid pixels, median 0.93 m).  The reference has no depth sequences and no expected outputs; this frame is
only an INPUT with real sensor noise and holes.  Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_real_fixture.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/g2o_frontend/PlaneEx_gui/test_images/image.pgm"


def read_pgm16(path):
    data = open(path, "rb").read()
    parts, i = [], 0
    while len(parts) < 4:
        while data[i:i + 1].isspace():
            i += 1
        if data[i:i + 1] == b"#":
            while data[i:i + 1] != b"\n":
                i += 1
            continue
        j = i
        while not data[j:j + 1].isspace():
            j += 1
        parts.append(data[i:j])
        i = j
    i += 1
    w, h = int(parts[1]), int(parts[2])
    return np.frombuffer(data[i:i + w * h * 2], dtype=">u2").reshape(h, w).astype(np.uint16)


if __name__ == "__main__":
    raw = read_pgm16(SRC)
    out = os.path.join(HERE, "real_depth_640x480.npz")
    np.savez_compressed(out, raw=raw, source=np.array(SRC))
    print(out, os.path.getsize(out), "bytes", raw.shape, float((raw > 0).mean()))
