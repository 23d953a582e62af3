"""Host side of the path (no GPU): scene / camera builders, image geometry, Philox host build,
PNG writer, the CLI contract of source.cpp:190-231, and that the C ABI exports what the headers
declare and refuses to compute without a device."""
import os
import re
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "uecraytracing_b200", "bin", "raytrace")


def test_abi_exports_every_declared_symbol(u):
    import ctypes
    declared = set()
    for h in ("uecrt.h", "uecrt_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(uecrt_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 18
    L = ctypes.CDLL(u.lib_path)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/ but not exported by libuecrt.so"
    from uecraytracing_b200._lib import EXPORTS
    assert declared == set(EXPORTS)
    assert u.lib.uecrt_version() == 2


def test_no_cpu_fallback(u):
    """Without a B200 every compute entry point must fail loudly (never fall back)."""
    if u.device_count() > 0:
        pytest.skip("a GPU is present")
    sc, cam = u.Scene.named("ref4"), u.Camera.named("ref")
    with pytest.raises(u.UecrtError, match="no CUDA device|not sm_100"):
        u.render(sc, cam, 16, 1)
    with pytest.raises(u.UecrtError):
        u.primary_hits(sc, [[0, 0, 0]], [[0, 0, -1]])
    with pytest.raises(u.UecrtError):
        u.philox(0, [[0, 0, 0, 0]])
    with pytest.raises(u.UecrtError):
        u.Context(sc, cam)
    # and the package never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "uecraytracing_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                assert "oracle" not in open(os.path.join(root, f)).read().lower().replace("oracle_free", ""), f


def test_bench_fixture_matches_the_scene_builders(u):
    """tests/golden/bench_scenes.npz (what bench.py's CPU arm renders) == what the product's host builders produce."""
    from oracle import pyoracle
    for name, cam in (("final", "final"), ("ref4", "ref"), ("dense", "dense")):
        fs, fc = pyoracle.bench_scene(name, cam)
        s, c = u.Scene.named(name, seed=1, n=10000), u.Camera.named(cam)
        assert np.array_equal(fs.geom, s.geom) and np.array_equal(fs.mat_type, s.mat_type) and np.array_equal(fs.mat_param, s.mat_param)
        assert bytes(fc.struct()) == bytes(c.struct())


def test_reference_arm_does_not_load_the_product_library():
    """bench.py --impl reference renders through oracle/ only: libuecrt.so must not be mapped (the driver records it)."""
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--workload', 'ref4']\n"
            "import bench\n"
            "bench.cpu_reference_arm.__defaults__ = (0.2,)\n"
            "orig = bench.cpu_reference_arm\n"
            "bench.cpu_reference_arm = lambda *a, **k: orig(*a[:6], budget_s=0.2)\n"
            "bench.main()\n"
            "m = open('/proc/self/maps').read()\n"
            "print('LOADED' if 'libuecrt' in m else 'CLEAN', 'ORACLE' if 'liboracle' in m else 'NOORACLE')\n")
    r = subprocess.run([os.sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert "CLEAN ORACLE" in r.stdout, r.stdout[-300:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][0]
    import json
    j = json.loads(line)
    assert j["impl"] == "reference" and j["gpu_launches"] == 0 and j["cpu_baseline"]["value"] == j["value"]


def test_image_height(u):
    # source.cpp:59-62: uint32(width / (16/9))
    assert [u.image_height(w) for w in (400, 1200, 1920, 3840, 16, 100)] == [225, 675, 1080, 2160, 9, 56]


def test_reference_scene_and_camera(u):
    s = u.Scene.named("ref4")       # source.cpp:103-112
    assert s.n == 4
    assert s.geom.tolist() == [[0, 0, -1, .5], [0, -100.5, -1, 100], [-1, 0, -1, .5], [1, 0, -1, .5]]
    assert s.mat_type.tolist() == [0, 0, 1, 1]
    np.testing.assert_allclose(s.mat_param[:, :3], [[.7, .3, .3], [.8, .8, 0], [.8, .8, .8], [.8, .6, .2]], rtol=1e-7)
    assert u.Scene.named("ref4d").mat_type.tolist() == [0, 0, 2, 1]
    c = u.Camera.named("ref").s     # camera.hpp:16-27
    assert list(c.origin) == [0, 0, 0] and c.lens_radius == 0
    np.testing.assert_allclose(list(c.lower_left), [-16 / 9, -1, -1], rtol=1e-7)
    np.testing.assert_allclose(list(c.horizontal), [32 / 9, 0, 0], rtol=1e-7)
    assert list(c.vertical) == [0, 2, 0]
    # the look-at form with the reference's parameters reproduces the default camera
    d = u.Camera.lookat((0, 0, 0), (0, 0, -1), (0, 1, 0), 90.0, 16 / 9, 0.0, 1.0).s
    for k in ("origin", "lower_left", "horizontal", "vertical"):
        np.testing.assert_allclose(list(getattr(d, k)), list(getattr(c, k)), atol=1e-6)


def test_final_and_dense_scenes(u):
    a, b = u.Scene.named("final", seed=1), u.Scene.named("final", seed=1)
    assert a.n == 484 and np.array_equal(a.geom, b.geom) and np.array_equal(a.mat_param, b.mat_param)
    assert a.geom[0].tolist() == [0, -1000, 0, 1000] and a.mat_type[0] == 0
    assert a.geom[-3:].tolist() == [[0, 1, 0, 1], [-4, 1, 0, 1], [4, 1, 0, 1]] and a.mat_type[-3:].tolist() == [2, 0, 1]
    small = a.geom[1:-3]
    assert np.all(small[:, 3] == np.float32(0.2)) and np.all(small[:, 1] == np.float32(0.2))
    assert np.all(np.hypot(small[:, 0] - 4, small[:, 2]) > 0.9)          # the book's exclusion zone
    assert 0.7 < (a.mat_type[1:-3] == 0).mean() < 0.9
    assert u.Scene.named("final", seed=2).n != 0 and not np.array_equal(u.Scene.named("final", seed=2).geom[:50], a.geom[:50])
    d = u.Scene.named("dense", seed=1, n=10000)
    assert d.n == 10000 and d.geom[0].tolist() == [0, -1000, 0, 1000]
    assert np.all((d.geom[1:, 3] >= 0.1) & (d.geom[1:, 3] <= 0.4)) and np.all(d.geom[1:, 1] == d.geom[1:, 3])
    assert np.all(np.abs(d.geom[1:, [0, 2]]) <= 50)
    assert u.Scene.named("dense", n=0).n == 0 and u.Scene.named("dense", n=1).n == 1
    with pytest.raises(ValueError):
        u.Scene.named("nope")


def test_philox_host_known_answers(u):
    # Random123 kat_vectors (philox4x32-10)
    assert u.philox_host(0, [0, 0, 0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert u.philox_host(0xffffffffffffffff, [0xffffffff] * 4) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert u.philox_host(0x299f31d0a4093822, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # stream independence: distinct (pixel, sample, bounce) -> distinct blocks; uniform mean ~ 0.5
    blocks = {tuple(u.philox_host(9, [p, s, b, 0])) for p in range(8) for s in range(8) for b in range(4)}
    assert len(blocks) == 256
    v = np.array([u.philox_host(1, [i, 0, 0, 0]) for i in range(2000)], np.uint64).ravel()
    assert abs((v >> 8).mean() / 2 ** 24 - 0.5) < 0.01


def test_png_writer_roundtrip(u, tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (9, 16, 3), dtype=np.uint8)
    path = tmp_path / "x.png"
    u.write_png(path, img)
    raw = path.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    # parse chunks by hand: IHDR 16x9 8-bit RGB, one zlib stream, filter byte 0 per row
    pos, chunks = 8, []
    while pos < len(raw):
        n = int.from_bytes(raw[pos:pos + 4], "big"); typ = raw[pos + 4:pos + 8]; data = raw[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + data) == int.from_bytes(raw[pos + 8 + n:pos + 12 + n], "big")
        chunks.append((typ, data)); pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    assert chunks[0][1] == (16).to_bytes(4, "big") + (9).to_bytes(4, "big") + bytes([8, 2, 0, 0, 0])
    rows = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(9, 1 + 16 * 3)
    assert np.all(rows[:, 0] == 0) and np.array_equal(rows[:, 1:].reshape(9, 16, 3), img)
    with pytest.raises(u.UecrtError):
        u.write_png("/nonexistent_dir/x.png", img)
    # tall images are deflated in independent bands on several host threads and stitched into ONE zlib
    # stream (sync-flushed raw deflate pieces + combined Adler-32): any zlib must decode it
    from conftest import read_png_rgb
    big = rng.integers(0, 256, (301, 67, 3), dtype=np.uint8)
    big[100:200] = 7                                    # a compressible band too
    u.write_png(tmp_path / "big.png", big)
    assert np.array_equal(read_png_rgb(tmp_path / "big.png"), big)
    try:
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(tmp_path / "big.png").convert("RGB")), big)
    except ImportError:
        pass


def test_scene_file_roundtrip_and_errors(u, tmp_path):
    """Runtime scene description (the step before the path; the reference hard-codes source.cpp:103-112)."""
    s = u.Scene.named("final", seed=1)
    path = tmp_path / "final.scene"
    u.save_scene_file(path, s)
    s2, cam = u.load_scene_file(path)
    assert cam is None and s2.n == s.n
    assert np.array_equal(s.geom, s2.geom) and np.array_equal(s.mat_type, s2.mat_type) and np.array_equal(s.mat_param, s2.mat_param)
    txt = tmp_path / "mini.scene"
    txt.write_text("# the reference scene, by hand\n"
                   "sphere 0 0 -1 0.5 lambertian 0.7 0.3 0.3\n"
                   "sphere 0 -100.5 -1 100 lambertian 0.8 0.8 0.0   # ground\n"
                   "\n"
                   "sphere -1 0 -1 0.5 metal 0.8 0.8 0.8\n"
                   "sphere 1 0 -1 0.5 metal 0.8 0.6 0.2 0.0\n"
                   "camera 13 2 3  0 0 0  20 0.1 10\n")
    s3, cam3 = u.load_scene_file(txt)
    ref = u.Scene.named("ref4")
    assert np.array_equal(s3.geom, ref.geom) and np.array_equal(s3.mat_type, ref.mat_type) and np.array_equal(s3.mat_param, ref.mat_param)
    fin = u.Camera.named("final").s
    for k in ("origin", "lower_left", "horizontal", "vertical", "u", "v"):
        assert list(getattr(cam3.s, k)) == list(getattr(fin, k))
    assert cam3.s.lens_radius == fin.lens_radius
    for bad, msg in (("sphere 0 0 0 1 plastic 1 1 1\n", "unknown material"), ("sphere 0 0\n", "expected"), ("cube 1\n", "unknown keyword"),
                     ("sphere 0 0 0 1 dielectric\n", "index of refraction")):
        b = tmp_path / "bad.scene"; b.write_text("# ok\n" + bad)
        with pytest.raises(ValueError, match=msg + "|bad.scene:2"):
            u.load_scene_file(b)
    with pytest.raises(ValueError, match="cannot open"):
        u.load_scene_file(tmp_path / "missing.scene")


def test_pfm_writer(u, tmp_path):
    acc = np.zeros((3, 5, 4), np.float32)
    acc[..., 0] = np.arange(15).reshape(3, 5) * 4; acc[..., 1] = 2; acc[..., 2] = 8; acc[..., 3] = 4
    u.write_pfm(tmp_path / "a.pfm", acc, 4)
    raw = (tmp_path / "a.pfm").read_bytes()
    head, body = raw[:raw.index(b"-1.0\n") + 5], raw[raw.index(b"-1.0\n") + 5:]
    assert head == b"PF\n5 3\n-1.0\n"
    img = np.frombuffer(body, "<f4").reshape(3, 5, 3)[::-1]          # PFM stores the bottom row first
    assert np.array_equal(img[..., 0], np.arange(15).reshape(3, 5)) and np.all(img[..., 1] == 0.5) and np.all(img[..., 2] == 2)


def _run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True, timeout=60)


@pytest.mark.skipif(not os.path.exists(BIN), reason="raytrace binary not built")
def test_cli_contract(tmp_path):
    # source.cpp:209-212: no output -> help, exit 0 ; -h likewise
    for args in ((), ("-h",), ("--help",)):
        r = _run(*args)
        assert r.returncode == 0 and "raytracing program" in r.stdout and "-o, --output arg" in r.stdout
        assert "-l, --verbose-level arg" in r.stdout and "--spp arg" in r.stdout
    # unknown option / second positional: uncaught exception -> abort, like the reference (probe C.5)
    assert _run("--bogus", "x.png").returncode == -6
    assert _run(str(tmp_path / "a.png"), "b.png").returncode == -6
    # with no GPU the render fails loudly: "rendering..." then "error", exit 1 (no CPU fallback)
    import uecraytracing_b200 as u
    if u.device_count() == 0:
        r = _run("-o", str(tmp_path / "a.png"), "--width", "16", "--spp", "1")
        assert r.returncode == 1 and r.stdout.split() == ["rendering...", "error"]
        assert "no CPU fallback" in r.stderr
