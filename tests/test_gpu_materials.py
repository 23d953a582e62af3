"""-m gpu: material scatter and Philox on the device against the oracle fed the same numbers."""
import numpy as np
import pytest

from conftest import golden, unhex

pytestmark = pytest.mark.gpu
TOL = 2e-5   # FP32 device vs FP64 oracle on O(1) vectors (north_star: <= 1e-5 relative on hits)


def test_philox_device_equals_host_and_kat(u):
    ctr = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32)
    assert u.philox(0, ctr[:1])[0].tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert u.philox(0xffffffffffffffff, ctr[1:2])[0].tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert u.philox(0x299f31d0a4093822, ctr[2:3])[0].tolist() == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    rng = np.random.default_rng(2)
    c = rng.integers(0, 2 ** 32, (5000, 4), dtype=np.uint64).astype(np.uint32)
    dev = u.philox(0x1234567890abcdef, c)
    for k in range(0, 5000, 97):
        assert dev[k].tolist() == u.philox_host(0x1234567890abcdef, c[k])
    assert abs((dev >> 8).mean() / 2 ** 24 - 0.5) < 0.005


def _compare(u, oracle, cases):
    mt = [c["mat_type"] for c in cases]
    ok, att, oo, do = u.scatter(mt, [c["mat_param"] for c in cases], [c["d_in"] for c in cases], [c["p"] for c in cases],
                                [c["normal"] for c in cases], [c["ff"] for c in cases], [c["uniforms"] for c in cases])
    for k, c in enumerate(cases):
        want = c["want"] if "want" in c else oracle.scatter(c["mat_type"], c["mat_param"], np.array(c["d_in"], np.float32),
                                                          np.array(c["p"], np.float32), np.array(c["normal"], np.float32), c["ff"], c["uniforms"])
        if c["mat_type"] == 1 and abs(np.dot(want[3], np.array(c["normal"], np.float64))) < 1e-4:
            continue                                   # fuzzed reflection grazing the surface: FP32 may absorb
        assert int(ok[k]) == want[0], k
        if want[0]:
            np.testing.assert_allclose(att[k], want[1], rtol=1e-6, atol=0)
            np.testing.assert_allclose(oo[k], want[2], rtol=0, atol=0)
            np.testing.assert_allclose(do[k], want[3], rtol=0, atol=TOL * max(1.0, np.abs(want[3]).max()))


def test_scatter_golden(u, oracle):
    cases = golden("scatter_replay.json")["cases"]
    for c in cases:
        c["want"] = (c["scattered"], unhex(c["att"]), unhex(c["o_out"]), unhex(c["d_out"]))
    _compare(u, oracle, cases)


def test_scatter_random_vs_oracle(u, oracle):
    rng = np.random.default_rng(8)
    cases = []
    for mt, mp in ((0, (.7, .3, .3, 0)), (1, (.8, .8, .8, 0)), (1, (.8, .6, .2, .45)), (2, (1, 1, 1, 1.5)), (2, (1, 1, 1, 2.4))):
        for k in range(200):
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            nrm = nrm.astype(np.float32); nrm /= np.float32(np.linalg.norm(nrm))
            din = rng.normal(size=3) * rng.uniform(0.2, 3)
            if np.dot(din, nrm) > 0:
                din = -din
            un = rng.integers(0, 1 << 24, 4) / float(1 << 24)
            if mt == 2 and k % 4 == 0:
                un[0] = 0.999                    # force the refract branch where it is possible
            cases.append({"mat_type": mt, "mat_param": list(mp), "d_in": din.astype(np.float32).tolist(), "p": rng.uniform(-3, 3, 3).astype(np.float32).tolist(),
                          "normal": nrm.tolist(), "ff": int(k & 1), "uniforms": un.tolist()})
    # drop cases that sit numerically ON a branch (Schlick ~ xi, metal dot ~ 0): FP32 may pick the other side
    keep = []
    for c in cases:
        n = np.array(c["normal"], np.float64); d = np.array(c["d_in"], np.float64); d /= np.linalg.norm(d)
        cos = min(-np.dot(d, n), 1.0)
        if c["mat_type"] == 2:
            eta = 1 / c["mat_param"][3] if c["ff"] else c["mat_param"][3]
            r0 = ((1 - eta) / (1 + eta)) ** 2; refl = r0 + (1 - r0) * (1 - cos) ** 5
            if abs(refl - c["uniforms"][0]) < 1e-4 or abs(eta * np.sqrt(max(0, 1 - cos * cos)) - 1) < 1e-4:
                continue
        keep.append(c)
    assert len(keep) > 900
    _compare(u, oracle, keep)


def test_dielectric_identities(u):
    """Self-checks for the extension with no reference test."""
    n = np.array([[0, 0, 1]] * 3, np.float32)
    d = np.array([[0, 0, -1]] * 3, np.float32)                        # normal incidence: cos = 1 -> R = R0 = 0.04
    ok, att, oo, do = u.scatter([2, 2, 2], [[1, 1, 1, 1.5]] * 3, d, np.zeros((3, 3)), n, [1, 1, 1],
                                [[0.039, 0, 0, 0], [0.041, 0, 0, 0], [0.5, 0, 0, 0]])
    assert ok.tolist() == [1, 1, 1] and np.all(att == 1)
    assert do[0].tolist() == [0, 0, 1]                                # xi < R0 -> reflect
    assert do[1].tolist() == [0, 0, -1] and do[2].tolist() == [0, 0, -1]   # refract straight through
    # total internal reflection iff eta*sin > 1 (inside glass, 60 degrees)
    d = np.array([[np.sin(np.pi / 3), 0, -np.cos(np.pi / 3)]], np.float32)
    ok, att, oo, do = u.scatter([2], [[1, 1, 1, 1.5]], d, np.zeros((1, 3)), n[:1], [0], [[0.999, 0, 0, 0]])
    np.testing.assert_allclose(do[0], [np.sin(np.pi / 3), 0, np.cos(np.pi / 3)], atol=1e-6)
    # refracted / reflected directions of unit inputs are unit
    rng = np.random.default_rng(0)
    v = rng.normal(size=(500, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True); v[:, 2] = -np.abs(v[:, 2])
    ok, att, oo, do = u.scatter([2] * 500, [[1, 1, 1, 1.5]] * 500, v, np.zeros((500, 3)), np.tile(n[:1], (500, 1)), [1] * 500,
                                rng.uniform(size=(500, 4)))
    np.testing.assert_allclose(np.linalg.norm(do, axis=1), 1.0, atol=2e-6)
