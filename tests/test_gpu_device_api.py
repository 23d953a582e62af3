"""-m gpu: the device-resident entry points bench.py uses (torch owns the memory and the stream)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_accumulate_and_resolve_on_torch_memory(u):
    import torch
    sc, cam = u.Scene.named("final", seed=1), u.Camera.named("final")
    W, S = 192, 16
    H = u.image_height(W)
    ctx = u.Context(sc, cam)
    accum = torch.zeros(H, W, 4, device="cuda")
    rgb = torch.zeros(H, W, 3, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ctx.accumulate(accum.data_ptr(), W, H, S, 50, seed=9, sample_begin=0, sample_count=10, stream=stream)
    ctx.accumulate(accum.data_ptr(), W, H, S, 50, seed=9, sample_begin=10, sample_count=6, stream=stream)
    u.resolve_device(accum.data_ptr(), W, H, S, rgb.data_ptr(), stream)
    torch.cuda.synchronize()
    st = ctx.stats()
    assert st["paths"] == W * H * 6 and st["kernel_ms"] > 0 and st["k1_regs"] > 0
    ref_rgb, ref_acc, _ = u.render(sc, cam, W, S, 50, seed=9, return_accum=True)
    a = accum.cpu().numpy()
    assert np.all(a[..., 3] == S)
    np.testing.assert_allclose(a[..., :3], ref_acc[..., :3], rtol=2e-5, atol=1e-5)
    assert (rgb.cpu().numpy() != ref_rgb).mean() < 1e-3
    ctx.close()
