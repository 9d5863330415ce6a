"""Numerics of the LayerNorm fold: LN(x) W^T = rstd (x W^T - mean rowsum(W)) with x rounded to fp16
BEFORE the mean is removed and the row statistics kept in fixed point.  CPU emulation of exactly what the two GEMM
epilogues do (PB200_EPI_RESID_LN_F32 / PB200_EPI_F16_LN), against fp32 LayerNorm followed by the same fp16-operand GEMM.

The fold's extra error grows with |row mean| / row std (fp16 rounding acts on x, not on x - mean): this test pins the
bound that DESIGN.md quotes, so a change of the fixed-point scales or of the formula shows up here without a GPU."""
import pytest
import torch


def _folded(x, w16, bias, c, shift=None):
    """shift: per-row fp32 scalar subtracted by the producer before the fp16 copy and the statistics (LayerNorm is
    invariant under it); the executor passes the row mean the previous AttnBlock of the same stream saw."""
    if shift is not None:
        x = x - shift[:, None]
    x16 = x.half()
    s = torch.round(x.double().sum(1) * 2 ** 20) / 2 ** 20            # per-row sum, 2^-20 fixed point
    q = torch.round((x.double() ** 2).sum(1) * 2 ** 16) / 2 ** 16     # per-row sum of squares, 2^-16 fixed point
    mean = (s / c).float()
    ex2 = (q / c).float()
    rstd = 1.0 / torch.sqrt(torch.clamp(ex2 - mean * mean, min=0.0) + 1e-6)
    acc = x16.float() @ w16.float().t()
    wsum = w16.float().sum(1)
    return (rstd[:, None] * (acc - mean[:, None] * wsum[None, :]) + bias).half().float()


def _plain(x, w16, bias, c):
    xn = torch.nn.functional.layer_norm(x, (c,), eps=1e-6).half()
    return (xn.float() @ w16.float().t() + bias).half().float()


@pytest.mark.parametrize("ratio,bound", [(0.0, 3e-3), (1.0, 4e-3), (4.0, 1e-2), (16.0, 4e-2)])
def test_fold_error_grows_with_mean_over_std(ratio, bound):
    g = torch.Generator().manual_seed(0)
    M, C, N = 512, 1280, 384
    x = torch.randn(M, C, generator=g) * 1.7 + ratio * 1.7
    w16 = (torch.randn(N, C, generator=g) / C ** 0.5).half()
    bias = torch.randn(N, generator=g) * 0.1
    exact = torch.nn.functional.layer_norm(x.double(), (C,), eps=1e-6) @ w16.double().t() + bias.double()
    e_fold = float((_folded(x, w16, bias, C).double() - exact).abs().max())
    e_plain = float((_plain(x, w16, bias, C).double() - exact).abs().max())
    # outputs are O(1): the plain path's error is the fp16 rounding of LN(x) and of the result (~2e-3)
    assert e_plain < 4e-3
    assert e_fold < bound, (ratio, e_fold, e_plain)


@pytest.mark.parametrize("ratio", [4.0, 16.0, 64.0])
def test_shifted_fold_is_as_accurate_as_the_plain_path(ratio):
    """With the per-row shift (the previous block's row mean: here the true mean perturbed by 30 % of a std, far more than
    one block changes it) the fold's error no longer depends on |mean| / std."""
    g = torch.Generator().manual_seed(1)
    M, C, N = 512, 1280, 384
    x = torch.randn(M, C, generator=g) * 1.7 + ratio * 1.7
    w16 = (torch.randn(N, C, generator=g) / C ** 0.5).half()
    bias = torch.randn(N, generator=g) * 0.1
    shift = x.mean(1) + 0.3 * 1.7 * torch.randn(M, generator=g)
    exact = torch.nn.functional.layer_norm(x.double(), (C,), eps=1e-6) @ w16.double().t() + bias.double()
    e_fold = float((_folded(x, w16, bias, C, shift).double() - exact).abs().max())
    e_plain = float((_plain(x, w16, bias, C).double() - exact).abs().max())
    assert e_plain < 4e-3 and e_fold < 4e-3, (ratio, e_fold, e_plain)


def test_fixed_point_statistics_cover_the_residual_stream_range():
    """|x| up to 1e3 over 2560 columns stays far inside int64 at the 2^20 / 2^16 scales."""
    c, big = 2560, 1e3
    assert c * big * 2 ** 20 < 2 ** 62 and c * big * big * 2 ** 16 < 2 ** 62
