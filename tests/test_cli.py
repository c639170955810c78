"""bicos-cli (libbicos_b200/bin/bicos-cli): the reference's command line tool (src/cli.cpp) on the
B200 path. CPU tests cover the option surface and the dependency-free image I/O against OpenCV's
codecs; the GPU test runs a whole folder -> disparity.tiff round trip against the oracle."""

import os
import subprocess

import numpy as np
import pytest

from libbicos_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "libbicos_b200", "bin", "bicos-cli")
cv2 = pytest.importorskip("cv2")


def _run(*args, **kw):
    return subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, timeout=300, **kw)


def test_cli_help_lists_the_reference_options():
    res = _run("--help")
    assert res.returncode == 0
    # reference src/cli.cpp:60-77
    for opt in ("--threshold", "--variance", "--step", "--out", "--stacksize", "--qmatrix", "--allow-negative-z",
                "--lr-maxdiff", "--double", "--limited", "--corrmap", "--no-dupes", "--help"):
        assert opt in res.stdout


def test_cli_argument_errors():
    assert _run().returncode == 1  # folder0 missing
    assert "does not exist" in _run("--bogus", "x").stderr
    res = _run("/nonexistent-folder")
    assert res.returncode == 1 and "bicos-cli:" in res.stderr


def _write_stacks(folder, left, right, paired, ext):
    os.makedirs(folder, exist_ok=True)
    if paired:
        for t in range(left.shape[0]):
            assert cv2.imwrite(os.path.join(folder, f"{t}_left.{ext}"), left[t])
            assert cv2.imwrite(os.path.join(folder, f"{t}_right.{ext}"), right[t])
        return [folder]
    l, r = os.path.join(folder, "l"), os.path.join(folder, "r")
    os.makedirs(l)
    os.makedirs(r)
    for t in range(left.shape[0]):
        assert cv2.imwrite(os.path.join(l, f"{t}.{ext}"), left[t])
        assert cv2.imwrite(os.path.join(r, f"{t}.{ext}"), right[t])
    return [l, r]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,ext,paired,args,kw", [
    (np.uint8, "png", True, ["-t", "0.9", "-v", "2.0", "--limited"],
     dict(nxcorr_threshold=0.9, min_variance=2.0)),
    (np.uint16, "png", False, ["--threshold=0.9", "-s", "0.25", "-m", "1", "--limited", "--corrmap", "-v", "0"],
     dict(nxcorr_threshold=0.9, subpixel_step=0.25, consistency=True, max_lr_diff=1)),
    (np.uint8, "pgm", False, ["-t", "0", "-n", "12", "-v", "0"],
     dict(nxcorr_threshold=None, mode_full=True)),
])
def test_cli_folder_to_disparity(tmp_path, oracles, dtype, ext, paired, args, kw):
    n = 20
    left, right, _ = synth.make_stacks(n, 256, 208, dtype, seed=17, row0=100, rows=40)
    folders = _write_stacks(str(tmp_path / "in"), left, right, paired, ext)
    out = tmp_path / "disp.png"
    res = _run(*folders, "-o", out, *args)
    assert res.returncode == 0, res.stderr
    used = 12 if "-n" in args else n
    assert f"Loaded {2 * used} {8 * np.dtype(dtype).itemsize}-bit images in total" in res.stdout
    want_d, want_c = oracles.port.match(left[:used], right[:used], **kw)
    got = cv2.imread(str(tmp_path / "disp.tiff"), cv2.IMREAD_UNCHANGED)
    assert got is not None and got.shape == want_d.shape
    if want_d.dtype == np.int16:
        assert got.dtype == np.int16 and np.array_equal(got, want_d)
    else:
        invalid = np.isnan(want_d) | (want_d == -32768)
        assert np.array_equal(np.isnan(got), invalid)  # the CLI masks -32768.0 like NaN
        assert np.array_equal(got[~invalid], want_d[~invalid])
    png = cv2.imread(str(out), cv2.IMREAD_UNCHANGED)
    assert png is not None and png.shape == want_d.shape + (3,)
    if "--corrmap" in args:
        corr = cv2.imread(str(tmp_path / "disp-corrmap.tiff"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(corr, want_c, equal_nan=True)


@pytest.mark.gpu
def test_cli_pointcloud(tmp_path):
    left, right, _ = synth.make_stacks(12, 256, 160, np.uint8, seed=3, row0=64, rows=24)
    folders = _write_stacks(str(tmp_path / "in"), left, right, True, "png")
    q = tmp_path / "q.yaml"
    q.write_text("%YAML:1.0\n---\nQ: !!opencv-matrix\n   rows: 4\n   cols: 4\n   dt: d\n"
                 "   data: [ 1., 0., 0., -80., 0., 1., 0., -12., 0., 0., 0., 500., 0., 0., 0.5, 0. ]\n")
    res = _run(*folders, "-o", tmp_path / "d.png", "-t", "0.8", "--limited", "-q", q)
    assert res.returncode == 0, res.stderr
    disp = cv2.imread(str(tmp_path / "d.tiff"), cv2.IMREAD_UNCHANGED)
    pts = np.loadtxt(tmp_path / "d.xyz").reshape(-1, 3)
    valid = ~np.isnan(disp) & (disp > 0)  # W = 0.5 d: positive disparities give finite points with Z > 0
    assert len(pts) == valid.sum() > 100
    rows, cols = np.nonzero(valid)
    w = 0.5 * disp[valid]
    assert np.allclose(pts[:, 0], (cols - 80.0) / w, rtol=1e-5) and np.allclose(pts[:, 2], 500.0 / w, rtol=1e-5)
