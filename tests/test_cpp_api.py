"""The C++ drop-in API (include/BICOS/match.hpp: BICOS::match, BICOS::match_sharded), exercised
by a plain-g++ consumer (tests/cpp/api_check.cpp) and compared with the oracle."""

import os
import struct
import subprocess

import numpy as np
import pytest

from libbicos_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "build", "api_check")
NP_TYPE = {3: np.int16, 5: np.float32, 6: np.float64}


def _write_input(path, left, right, kw):
    n, rows, cols = left.shape
    depth = 2 if left.dtype == np.uint16 else 0

    def opt(v):
        return -1.0 if v is None else float(v)

    with open(path, "wb") as f:
        f.write(struct.pack("<9i", n, rows, cols, depth, int(kw.get("mode_full", False)) | (2 if kw.get("wide_descriptors") else 0),
                            int(kw.get("double", False)),
                            int(kw.get("consistency", False)), int(kw.get("max_lr_diff", 1)),
                            int(kw.get("no_dupes", False))))
        f.write(struct.pack("<3f", opt(kw.get("nxcorr_threshold", 0.5)), opt(kw.get("subpixel_step")),
                            opt(kw.get("min_variance"))))
        f.write(np.ascontiguousarray(left).tobytes())
        f.write(np.ascontiguousarray(right).tobytes())


def _read_output(path):
    raw = open(path, "rb").read()
    dt, ct, rows, cols = struct.unpack("<4i", raw[:16])
    d = np.frombuffer(raw, dtype=NP_TYPE[dt], count=rows * cols, offset=16).reshape(rows, cols)
    c = None
    if ct:
        c = np.frombuffer(raw, dtype=NP_TYPE[ct], count=rows * cols, offset=16 + d.nbytes).reshape(rows, cols)
    return d, c


def test_api_check_binary_is_built():
    """CPU-side: the consumer compiled and linked against the product library with g++ alone."""
    assert os.path.exists(EXE), "run `make -C libbicos_b200/csrc` (or __graft_entry__.build())"


CASES = [
    (33, np.uint8, dict(nxcorr_threshold=0.96, min_variance=2.0)),
    (33, np.uint8, dict(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1, consistency=True, max_lr_diff=1)),
    (16, np.uint16, dict(nxcorr_threshold=0.9, mode_full=True, double=True, min_variance=1.0)),
    (9, np.uint8, dict(nxcorr_threshold=None, consistency=True, max_lr_diff=2, no_dupes=True)),
    (20, np.uint16, dict(nxcorr_threshold=0.9, mode_full=True, wide_descriptors=True, min_variance=1.0)),  # extension
]


@pytest.mark.gpu
@pytest.mark.parametrize("n,dtype,kw", CASES)
@pytest.mark.parametrize("sharded", [False, True])
def test_cpp_match_equals_oracle(tmp_path, oracles, n, dtype, kw, sharded):
    import torch

    left, right, _ = synth.make_stacks(n, 256, 272, dtype, seed=5 * n, row0=64, rows=45)
    okw = dict(kw)
    want_d, want_c = oracles.port.match(left, right, **okw)
    _write_input(tmp_path / "in.bin", left, right, kw)
    env = dict(os.environ, API_CHECK_DEVICES=str(torch.cuda.device_count()))
    args = [EXE, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")] + (["sharded"] if sharded else [])
    res = subprocess.run(args, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr
    got_d, got_c = _read_output(tmp_path / "out.bin")
    assert got_d.dtype == want_d.dtype
    assert np.array_equal(got_d, want_d, equal_nan=True)
    if want_c is None:
        assert got_c is None
    else:
        assert got_c.dtype == want_c.dtype and np.array_equal(got_c, want_c, equal_nan=True)
