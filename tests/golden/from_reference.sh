#!/bin/sh
# Pin the oracle to the reference itself: needs a Rust toolchain (cargo) and the reference checkout at /root/reference.
# NOT runnable in the build image (no rustc); kept so that parity gets pinned the moment anyone has one.
#   1. case files (inputs of tests/test_oracle.py::golden_cases)      -> oracle/_ref/cases/*.case
#   2. oracle/ref_harness (path dependency on /root/reference/src/core) runs undistort_image_cpu on each
#   3. the reference's output bytes                                   -> tests/golden/ref_<case>.bin   (commit these)
# tests/test_oracle.py::test_reference_fixtures then compares the C oracle with every ref_*.bin byte for byte.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
command -v cargo >/dev/null 2>&1 || { echo "cargo not found: the reference is Rust (edition 2024); oracle/_ref stays unbuilt"; exit 3; }
python3 "$ROOT/tests/golden/make_ref_cases.py" "$ROOT/oracle/_ref/cases"
cargo build --release --manifest-path "$ROOT/oracle/ref_harness/Cargo.toml" --target-dir "$ROOT/oracle/_ref/target"
"$ROOT/oracle/_ref/target/release/gf_ref_harness" "$ROOT/oracle/_ref/cases" "$ROOT/tests/golden"
