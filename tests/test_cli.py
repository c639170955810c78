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


IOCHK = os.path.join(ROOT, "tests", "cpp", "build", "imageio_check")


def _read_with_tool(path, tmp):
    raw = os.path.join(tmp, "img.raw")
    res = subprocess.run([IOCHK, "read", path, raw], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    buf = open(raw, "rb").read()
    rows, cols, bits, colour = np.frombuffer(buf, dtype=np.int32, count=4)
    data = np.frombuffer(buf, dtype=np.uint16 if bits == 16 else np.uint8, offset=16).reshape(rows, cols)
    return data, bool(colour)


@pytest.mark.parametrize("dtype,ext", [(np.uint8, "png"), (np.uint16, "png"), (np.uint8, "pgm"), (np.uint16, "pgm")])
def test_image_reader_matches_opencv_grey(tmp_path, dtype, ext):
    rng = np.random.default_rng(5)
    img = rng.integers(0, np.iinfo(dtype).max + 1, size=(37, 53)).astype(dtype)
    path = str(tmp_path / f"a.{ext}")
    assert cv2.imwrite(path, img)
    got, colour = _read_with_tool(path, str(tmp_path))
    assert not colour and got.dtype == dtype and np.array_equal(got, img)


TIFF_FLAGS = {"none": 1, "lzw": 5, "deflate": 32946, "packbits": 32773}


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("compression", sorted(TIFF_FLAGS))
def test_tiff_reader_matches_opencv(tmp_path, dtype, compression):
    """TIFF input (16-bit camera dumps): every compression OpenCV's writer offers, one strip and several, smooth
    content (long LZW strings, the horizontal predictor OpenCV switches on for LZW / Deflate) and noise."""
    rng = np.random.default_rng(7)
    hi = np.iinfo(dtype).max
    smooth = (np.add.outer(np.arange(301), np.arange(517)) * (hi // 900)).astype(dtype)
    noise = rng.integers(0, hi + 1, size=(64, 129)).astype(dtype)
    flat = np.full((70, 333), hi // 3, dtype)
    for k, img in enumerate((smooth, noise, flat)):
        for rps in (0, 7):
            path = str(tmp_path / f"t{k}_{rps}.tiff")
            params = [cv2.IMWRITE_TIFF_COMPRESSION, TIFF_FLAGS[compression]]
            if rps:
                params += [cv2.IMWRITE_TIFF_ROWSPERSTRIP, rps]
            assert cv2.imwrite(path, img, params)
            got, colour = _read_with_tool(path, str(tmp_path))
            assert not colour and got.dtype == dtype and np.array_equal(got, img), (k, rps)
            assert np.array_equal(cv2.imread(path, cv2.IMREAD_UNCHANGED), img)


def test_tiff_reader_colour_big_endian_and_errors(tmp_path):
    rng = np.random.default_rng(8)
    bgr = rng.integers(0, 256, size=(21, 40, 3)).astype(np.uint8)
    path = str(tmp_path / "rgb.tif")
    assert cv2.imwrite(path, bgr, [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    got, colour = _read_with_tool(path, str(tmp_path))
    want = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    assert colour and np.abs(got.astype(int) - want.astype(int)).max() <= 1
    # hand-made big-endian ("MM"), WhiteIsZero, uncompressed files: whatever cv::imread hands the reference is the truth
    # (it inverts 8-bit WhiteIsZero samples and returns 16-bit ones as stored)
    for dtype, bits in ((np.uint8, 8), (np.uint16, 16)):
        img = rng.integers(0, np.iinfo(dtype).max + 1, size=(5, 9)).astype(dtype)
        data = img.astype(">u2").tobytes() if bits == 16 else img.tobytes()
        tags = [(256, 3, 1, 9), (257, 3, 1, 5), (258, 3, 1, bits), (259, 3, 1, 1), (262, 3, 1, 0), (273, 4, 1, 8),
                (277, 3, 1, 1), (278, 3, 1, 5), (279, 4, 1, len(data))]
        ifd = len(tags).to_bytes(2, "big")
        for tag, typ, cnt, val in tags:
            ifd += tag.to_bytes(2, "big") + typ.to_bytes(2, "big") + cnt.to_bytes(4, "big")
            ifd += (val.to_bytes(2, "big") + b"\0\0") if typ == 3 else val.to_bytes(4, "big")
        ifd += (0).to_bytes(4, "big")
        blob = b"MM" + (42).to_bytes(2, "big") + (8 + len(data)).to_bytes(4, "big") + data + ifd
        mm = tmp_path / f"mm{bits}.tif"
        mm.write_bytes(blob)
        got, colour = _read_with_tool(str(mm), str(tmp_path))
        want = cv2.imread(str(mm), cv2.IMREAD_GRAYSCALE | cv2.IMREAD_ANYDEPTH)  # the reference's flags (fileutils.cpp:72-75)
        assert not colour and got.dtype == dtype and np.array_equal(got, want)
        assert np.array_equal(want, 255 - img if bits == 8 else img)
    # refused with a message, never a crash: truncated file, float samples
    (tmp_path / "cut.tif").write_bytes(blob[: len(blob) // 2])
    res = subprocess.run([IOCHK, "read", str(tmp_path / "cut.tif"), str(tmp_path / "x.raw")], capture_output=True, text=True)
    assert res.returncode == 1 and "cut.tif" in res.stderr
    f32 = str(tmp_path / "f.tiff")
    assert cv2.imwrite(f32, rng.random((4, 4)).astype(np.float32))
    res = subprocess.run([IOCHK, "read", f32, str(tmp_path / "x.raw")], capture_output=True, text=True)
    assert res.returncode == 1 and "TIFF" in res.stderr


def test_image_reader_colour_and_alpha(tmp_path):
    rng = np.random.default_rng(6)
    bgr = rng.integers(0, 256, size=(21, 40, 3)).astype(np.uint8)
    for name, img in (("rgb.png", bgr), ("rgba.png", np.dstack([bgr, np.full((21, 40), 200, np.uint8)]))):
        path = str(tmp_path / name)
        assert cv2.imwrite(path, img)
        got, colour = _read_with_tool(path, str(tmp_path))
        want = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)  # the reference reads with IMREAD_GRAYSCALE
        assert colour and np.abs(got.astype(int) - want.astype(int)).max() <= 1


def test_image_writers_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    disp = rng.normal(40, 5, size=(30, 44)).astype(np.float32)
    disp[3:6] = np.nan
    raw = tmp_path / "d.raw"
    raw.write_bytes(disp.tobytes())
    res = subprocess.run([IOCHK, "write", "30", "44", "5", str(raw), str(tmp_path / "out")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    tiff = cv2.imread(str(tmp_path / "out.tiff"), cv2.IMREAD_UNCHANGED)
    assert tiff.dtype == np.float32 and np.array_equal(tiff, disp, equal_nan=True)
    png = cv2.imread(str(tmp_path / "out.png"), cv2.IMREAD_UNCHANGED)
    assert png.shape == (30, 44, 3) and (png[3:6] == 0).all() and (png[10:] != 0).any()
    # the reference's save_image: min-max normalise the valid pixels to 0..255, COLORMAP_TURBO, invalid black
    valid = ~np.isnan(disp)
    norm = cv2.normalize(disp, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1, valid.astype(np.uint8))
    norm[~valid] = 0
    want = cv2.applyColorMap(norm, cv2.COLORMAP_TURBO)
    want[~valid] = 0
    assert (np.abs(png.astype(int) - want.astype(int)).max(axis=2) <= 12).all()  # at most one grey level apart
    assert (png == want).all(axis=2).mean() > 0.95
    i16 = rng.integers(-50, 300, size=(30, 44)).astype(np.int16)
    i16[0, :5] = -32768
    raw.write_bytes(i16.tobytes())
    assert subprocess.run([IOCHK, "write", "30", "44", "3", str(raw), str(tmp_path / "o2")]).returncode == 0
    assert np.array_equal(cv2.imread(str(tmp_path / "o2.tiff"), cv2.IMREAD_UNCHANGED), i16)


def test_q_matrix_yaml_and_xml(tmp_path):
    q = np.arange(16, dtype=np.float64).reshape(4, 4) * 1.5 - 3
    for ext in ("yaml", "xml"):
        path = str(tmp_path / f"q.{ext}")
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
        fs.write("Q", q)
        fs.release()
        res = subprocess.run([IOCHK, "q", path], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        assert np.allclose(np.array(res.stdout.split(), dtype=np.float64), q.ravel())


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
    (np.uint16, "tiff", True, ["-t", "0.9", "-v", "2.0", "--limited", "-s", "0.5"],
     dict(nxcorr_threshold=0.9, min_variance=2.0, subpixel_step=0.5)),
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
