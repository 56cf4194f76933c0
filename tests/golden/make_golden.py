#!/usr/bin/env python3
"""Regenerate tests/golden/oracle_crc.json from the CPU oracle.  The reference has no golden vectors for
this path and cannot be built here (Rust), so these are self-generated known answers: parity stays "unpinned"."""
import json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_oracle import golden_cases, run_oracle

out = {}
for name, case in golden_cases().items():
    rc, p, src, dst = run_oracle(case)
    assert rc == 0, name
    out[name] = zlib.crc32(dst.tobytes())
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_crc.json"), "w"), indent=1, sort_keys=True)
print("wrote %d golden CRCs" % len(out))
