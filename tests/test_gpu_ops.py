"""Parity of single CUDA kernels (through the C ABI) against torch fp32 on the same fp16-rounded operands."""

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from gpu_ops import OpHandle
    h = OpHandle(0)
    yield h
    h.close()


def r16(t):
    return t.to(torch.float16).to(torch.float32)


def _check(out, ref, tol=2e-3):
    # fp32 accumulation of fp16 operands, fp16 output rounding (2^-11 relative)
    scale = float(ref.abs().max())
    err = float((out.cpu() - ref).abs().max())
    assert err <= tol * max(scale, 1.0), f"max err {err} vs scale {scale}"


CONV_CASES = [
    # kind, N, H, W, C0, C1, Cout
    (0, 1, 8, 16, 64, 0, 64),
    (0, 2, 10, 24, 64, 0, 128),      # BH=2, ragged W
    (0, 2, 20, 40, 128, 0, 256),     # BN=256
    (0, 1, 10, 32, 512, 0, 1024),    # 4 N tiles, long K (72 k-steps)
    (0, 2, 40, 56, 128, 128, 128),   # two K sources (skip concat)
    (0, 3, 80, 64, 128, 0, 128),     # many tiles per CTA wave (persistent loop): 3*5*8=120.. plus next
    (0, 8, 80, 96, 64, 0, 64),       # 8*5*12=480 tiles > 148 CTAs
    (1, 2, 16, 32, 64, 0, 64),       # stride 2
    (1, 2, 80, 48, 128, 0, 128),
    (1, 1, 20, 24, 512, 0, 512),
    (2, 2, 8, 16, 64, 0, 384),       # 1x1 (to_qkv shape)
    (2, 2, 20, 24, 256, 256, 128),   # 1x1 res_conv over the concat
    (3, 2, 10, 12, 64, 0, 64),       # transposed 4x4 s2
    (3, 1, 20, 64, 256, 0, 256),
    # Cout == 128 with H % 4 == 0 runs the swapped-operand kernel (channels on the MMA M axis, 256-pixel tiles)
    (0, 1, 20, 40, 128, 0, 128),     # BH=4, BW=64, ragged W
    (0, 2, 80, 72, 256, 0, 128),     # BH=16, BW=16, ragged W, long K
    (3, 2, 20, 24, 128, 0, 128),     # transposed conv, W < BW
    (3, 1, 40, 64, 128, 0, 128),
    (1, 2, 40, 96, 128, 0, 128),     # stride 2 -> tile space 20 x 48
]


@pytest.mark.parametrize("kind,N,H,W,C0,C1,Cout", CONV_CASES)
def test_conv_matches_torch(ops, kind, N, H, W, C0, C1, Cout):
    g = torch.Generator().manual_seed(kind * 1000 + H * 10 + W + C0)
    Cin = C0 + C1
    x0 = r16(torch.randn(N, C0, H, W, generator=g))
    x1 = r16(torch.randn(N, C1, H, W, generator=g)) if C1 else None
    k = {0: 3, 1: 3, 2: 1, 3: 4}[kind]
    wshape = (Cin, Cout, k, k) if kind == 3 else (Cout, Cin, k, k)
    w = r16(torch.randn(wshape, generator=g) / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, generator=g)
    xin = torch.cat([x0, x1], 1) if C1 else x0
    if kind == 0:
        ref = F.conv2d(xin, w, b, padding=1)
    elif kind == 1:
        ref = F.conv2d(xin, w, b, stride=2, padding=1)
    elif kind == 2:
        ref = F.conv2d(xin, w, b)
    else:
        ref = F.conv_transpose2d(xin, w, b, stride=2, padding=1)
    out, _ = ops.conv(kind, x0, w, b, x1=x1)
    assert torch.isfinite(out).all()
    _check(out, ref)


def test_conv_epilogue_stats_mask_residual(ops):
    g = torch.Generator().manual_seed(5)
    N, C, H, W, Cout = 2, 128, 20, 40, 256
    x = r16(torch.randn(N, C, H, W, generator=g))
    w = r16(torch.randn(Cout, C, 3, 3, generator=g) / (C * 9) ** 0.5)
    b = torch.randn(Cout, generator=g)
    conv = F.conv2d(x, w, b, padding=1)
    # stats of conv+bias per (sample, group)
    out, st = ops.conv(0, x, w, b, stats_groups=8)
    cg = conv.reshape(N, 8, -1).double()
    ref_st = torch.stack([cg.sum(-1), (cg * cg).sum(-1)], -1)
    assert torch.allclose(st.cpu(), ref_st, rtol=2e-5, atol=1e-2), (st.cpu() - ref_st).abs().max()
    _check(out, conv)
    # mask + residual blend: (conv*g + res) * mask
    mask = (torch.arange(W).unsqueeze(0) < torch.tensor([W, 27]).unsqueeze(1)).float()
    res = r16(torch.randn(N, Cout, H, W, generator=g))
    out2, _ = ops.conv(0, x, w, b, mask=mask, residual=res, res_scale=0.37)
    ref2 = (conv * 0.37 + res) * mask.view(N, 1, 1, W)
    _check(out2, ref2)


def test_swapped_conv_stats_and_mask(ops):
    """Cout == 128: statistics and output mask through the swapped-operand kernel."""
    g = torch.Generator().manual_seed(9)
    N, C, H, W, Cout = 2, 128, 40, 56, 128
    x = r16(torch.randn(N, C, H, W, generator=g))
    w = r16(torch.randn(Cout, C, 3, 3, generator=g) / (C * 9) ** 0.5)
    b = torch.randn(Cout, generator=g)
    conv = F.conv2d(x, w, b, padding=1)
    out, st = ops.conv(0, x, w, b, stats_groups=8)
    cg = conv.reshape(N, 8, -1).double()
    ref_st = torch.stack([cg.sum(-1), (cg * cg).sum(-1)], -1)
    assert torch.allclose(st.cpu(), ref_st, rtol=2e-5, atol=1e-2), (st.cpu() - ref_st).abs().max()
    _check(out, conv)
    mask = (torch.arange(W).unsqueeze(0) < torch.tensor([W, 31]).unsqueeze(1)).float()
    out2, _ = ops.conv(0, x, w, b, mask=mask)
    _check(out2, conv * mask.view(N, 1, 1, W))
    # transposed conv with the (finer) output mask
    wt = r16(torch.randn(C, Cout, 4, 4, generator=g) / (C * 4) ** 0.5)
    mask2 = (torch.arange(2 * W).unsqueeze(0) < torch.tensor([2 * W, 61]).unsqueeze(1)).float()
    out3, _ = ops.conv(3, x, wt, b, mask=mask2)
    _check(out3, F.conv_transpose2d(x, wt, b, stride=2, padding=1) * mask2.view(N, 1, 1, 2 * W))


@pytest.mark.parametrize("C,cpg_note", [(64, "cpg8"), (128, "cpg16"), (512, "cpg64"), (1024, "cpg128")])
def test_conv_stats_group_widths(ops, C, cpg_note):
    g = torch.Generator().manual_seed(C)
    N, H, W = 2, 10, 16
    x = r16(torch.randn(N, 64, H, W, generator=g))
    w = r16(torch.randn(C, 64, 3, 3, generator=g) / 24.0)
    b = torch.randn(C, generator=g)
    conv = F.conv2d(x, w, b, padding=1)
    _, st = ops.conv(0, x, w, b, stats_groups=8)
    cg = conv.reshape(N, 8, -1).double()
    ref_st = torch.stack([cg.sum(-1), (cg * cg).sum(-1)], -1)
    assert torch.allclose(st.cpu(), ref_st, rtol=2e-5, atol=1e-2)


def _mish(x):
    return x * torch.tanh(F.softplus(x))


@pytest.mark.parametrize("C", [64, 128, 1024])
def test_gn_apply_matches_torch(ops, C):
    g = torch.Generator().manual_seed(C + 1)
    N, H, W = 2, 10, 24
    raw = r16(torch.randn(N, C, H, W, generator=g) * 3 + 0.5)
    raw[0, 0, 0, 0] = 30.0  # exercises the softplus threshold branch
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    addvec = torch.randn(N, C, generator=g)
    res = r16(torch.randn(N, C, H, W, generator=g))
    mask = (torch.arange(W).unsqueeze(0) < torch.tensor([W, 17]).unsqueeze(1)).float()
    rg = raw.reshape(N, 8, -1).double()
    stats = torch.stack([rg.sum(-1), (rg * rg).sum(-1)], -1)
    ref = (_mish(F.group_norm(raw, 8, gamma, beta, eps=1e-5)) + addvec[:, :, None, None] + res) * mask.view(N, 1, 1, W)
    out = ops.gn_apply(raw, stats, gamma, beta, addvec, res, mask)
    _check(out, ref, tol=1.5e-3)
    out2 = ops.gn_apply(raw, stats, gamma, beta, None, None, mask)
    ref2 = _mish(F.group_norm(raw, 8, gamma, beta, eps=1e-5)) * mask.view(N, 1, 1, W)
    _check(out2, ref2, tol=1.5e-3)


@pytest.mark.parametrize("H,W", [(10, 16), (40, 64), (80, 40)])
def test_attention_context_matches_torch(ops, H, W):
    g = torch.Generator().manual_seed(H * W)
    N, C, heads, dh = 2, 256, 4, 32
    qkv = r16(torch.randn(N, 3 * heads * dh, H, W, generator=g) * 1.5)
    wo = torch.randn(C, heads * dh, generator=g) / (heads * dh) ** 0.5
    q, k, v = qkv.reshape(N, 3, heads, dh, H * W).unbind(1)
    ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(-1), v)
    # Weff[b, co, h*32+d] = sum_e wo[co, h*32+e] * ctx[b,h,d,e]
    ref = torch.einsum("chE,bhdE->bchd", wo.reshape(C, heads, dh), ctx).reshape(N, C, heads * dh)
    out = ops.attn_context(qkv, wo)
    _check(out, ref, tol=2e-3)


def test_cluster_multicast_conv_variant_in_a_subprocess():
    """USB_MC=1 (read once per process) launches the swapped-operand conv as clusters of two CTAs that TMA-multicast the
    activation patches to each other; the same conv cases must pass with it."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, USB_MC="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "test_conv_matches_torch or test_conv_stats_group_widths"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
