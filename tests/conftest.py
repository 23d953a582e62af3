import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def unhex(xs, shape=None):
    a = np.array([float.fromhex(x) for x in xs], np.float64)
    return a.reshape(shape) if shape else a


@pytest.fixture(scope="session")
def port():
    from oracle import pyoracle
    return pyoracle.load("port")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference headers (oracle/_ref); skipped where they were never built."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return pyoracle.load("ref")


@pytest.fixture(scope="session")
def oracle():
    """Best available checker: reference harness when prebuilt, else the restatement."""
    from oracle import pyoracle
    return pyoracle.load("best")


@pytest.fixture(scope="session")
def u():
    import uecraytracing_b200
    return uecraytracing_b200


def psnr8(a, b):
    mse = np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / mse) if mse > 0 else 99.0


def primary_rays(cam, W, H, jitter=None):
    """Pixel-centre rays of a flattened camera (float32, as the device sees them)."""
    ys, xs = np.mgrid[0:H, 0:W]
    uu = (xs.ravel() + 0.5) / W
    vv = (H - ys.ravel() - 1 + 0.5) / H
    c = cam.s
    ll, hor, ver, org = (np.array(getattr(c, k), np.float64) for k in ("lower_left", "horizontal", "vertical", "origin"))
    d = ll[None] + uu[:, None] * hor[None] + vv[:, None] * ver[None] - org[None]
    o = np.broadcast_to(org, d.shape)
    return np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)


def read_png_rgb(path):
    """Decode an 8-bit RGB PNG written by uecrt_host_write_png (single IDAT, filter 0 on every row)."""
    import zlib
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = int.from_bytes(raw[16:20], "big"), int.from_bytes(raw[20:24], "big")
    pos, idat = 8, b""
    while pos < len(raw):
        n = int.from_bytes(raw[pos:pos + 4], "big"); typ = raw[pos + 4:pos + 8]
        if typ == b"IDAT":
            idat += raw[pos + 8:pos + 8 + n]
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert np.all(rows[:, 0] == 0)
    return rows[:, 1:].reshape(h, w, 3)
