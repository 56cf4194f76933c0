#!/usr/bin/env python3
"""Write the inputs of every golden case (tests/test_oracle.py::golden_cases) as case files for oracle/ref_harness — the Rust
binary that runs the UNMODIFIED reference on them (tests/golden/from_reference.sh).  Format: see oracle/ref_harness/src/main.rs."""
import ctypes as C
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from gyroflow_b200 import abi
from tests import cases
from tests.test_oracle import golden_cases


def write_case(path, case):
    p, src, m, mesh, dst0, pix, lens, digital = cases.build(case)
    bw, bh = case.get("in_size", (case["w"], case["h"]))
    obw, obh = case.get("out_size", (case.get("ow", case["w"]), case.get("oh", case["h"])))
    mesh = np.zeros(0, np.float32) if mesh is None else np.ascontiguousarray(mesh, np.float32)
    m = np.ascontiguousarray(m, np.float32)
    with open(path, "wb") as f:
        f.write(b"GFCASE1\0")
        f.write(struct.pack("<4I", abi.PIXEL_TYPES[pix][0], abi.LENS[lens], abi.LENS[digital] if digital else 0, p.interpolation))
        f.write(struct.pack("<6I", bw, bh, p.stride, obw, obh, p.output_stride))
        f.write(C.string_at(C.byref(p), C.sizeof(p)))
        f.write(struct.pack("<Q", m.shape[0])); f.write(m.tobytes())
        f.write(struct.pack("<Q", mesh.size)); f.write(mesh.tobytes())
        f.write(struct.pack("<Q", src.nbytes)); f.write(src.tobytes())
        f.write(struct.pack("<Q", dst0.nbytes)); f.write(dst0.tobytes())


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "oracle", "_ref", "cases")
    os.makedirs(out, exist_ok=True)
    for name, case in golden_cases().items():
        write_case(os.path.join(out, name + ".case"), case)
    print("wrote %d case files to %s" % (len(golden_cases()), out))
