"""GPU: pre / post-processing kernels around the nets against the CPU oracle
(the reference does these with torch.nn.functional.interpolate at test.py:58,71,100-102)."""
import pytest
import torch

from oracle import dvc_oracle as O
from oracle.weights import make_lab

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 3, 32, 48), (2, 3, 432, 768), (1, 1, 6, 4)])
def test_resize_half_matches_interpolate(ctx, shape):
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)) * 40
    ref = O.resize_half(x)
    out = ctx.resize_half(x.cuda()).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max() <= 1e-5 * x.abs().max()


@pytest.mark.parametrize("shape", [(1, 2, 16, 24), (2, 2, 216, 384), (1, 2, 1, 1), (1, 2, 3, 5)])
def test_upsample2_scaled_matches_interpolate(ctx, shape):
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(2)) * 60
    ref = O.upsample2_scaled(x, 1.25)
    out = ctx.upsample2_scaled(x.cuda(), 1.25).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max() <= 1e-5 * x.abs().max()


def test_half_resolution_pipeline_roundtrip(ctx):
    """test.py data flow: 432x768 Lab frame -> 1/2 -> colourise -> x2 * 1.25."""
    big = make_lab(5, 1, 64, 96)
    IB = ctx.resize_half(make_lab(6, 1, 64, 96).cuda())
    ctx.set_exemplar(IB)
    IA = ctx.resize_half(big.cuda())
    ab = ctx.colorize_frames(IA[:, 0:1].contiguous(), torch.zeros(1, 3, 32, 48, device="cuda"))
    up = ctx.upsample2_scaled(ab)
    assert up.shape == (1, 2, 64, 96) and torch.isfinite(up).all()
    assert torch.allclose(up.cpu(), O.upsample2_scaled(ab.cpu()), atol=1e-4)
