"""Container I/O for oracle/ref_harness (format in ref_harness.cu) — test infrastructure."""
import struct
import numpy as np


def write_box(path, box):
    with open(path, "wb") as f:
        f.write(b"ZKH1"); f.write(struct.pack("<I", len(box)))
        for name, arr in box.items():
            a = np.ascontiguousarray(arr)
            if a.dtype == np.float32:
                a = a.view(np.uint32)
            a = a.astype(np.uint32, copy=False).reshape(-1)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb))); f.write(nb); f.write(struct.pack("<Q", a.size)); f.write(a.tobytes())


def read_box(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(4) == b"ZKH1"
        (cnt,) = struct.unpack("<I", f.read(4))
        for _ in range(cnt):
            (nl,) = struct.unpack("<I", f.read(4)); name = f.read(nl).decode()
            (nw,) = struct.unpack("<Q", f.read(8))
            out[name] = np.frombuffer(f.read(4 * nw), dtype=np.uint32).copy()
    return out
