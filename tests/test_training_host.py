"""Host side of the fine-tune step, no GPU: parameter inventory, flat-buffer layout and the ABI call sequence of one
training step (the graph walk of unitspeech_b200/training.py recorded instead of launched)."""

import collections

import pytest
import torch

from oracle import unitspeech_oracle as O
from unitspeech_b200 import abi
from unitspeech_b200.training import FineTuner, param_shapes
from train_cases import case_inputs


@pytest.mark.parametrize("dim,mults", [(64, (1, 2)), (128, (1, 2, 4, 8))])
def test_parameter_inventory_matches_reference_state_dict(dim, mults):
    mine, ref = param_shapes(80, dim, mults, 256), O.param_shapes(80, dim, mults, 256)
    assert set(mine) == set(ref) and all(tuple(mine[k]) == tuple(ref[k]) for k in ref)


def test_training_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(abi.UsbError):
        FineTuner(dim=64, dim_mults=(1, 2))


@pytest.mark.parametrize("dim,mults,B,T", [(64, (1, 2), 2, 16), (128, (1, 2, 4, 8), 2, 24)])
def test_step_call_sequence(dim, mults, B, T):
    trace = []
    ft = FineTuner(dim=dim, dim_mults=mults, _trace_calls=trace)
    ft.load_state_dict(O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=4.0))
    # flat layout: 256-byte aligned views, the 16 ResnetBlock.mlp Linears contiguous (one stacked Linear)
    assert all(v.data_ptr() % 256 == ft.P.data_ptr() % 256 for v in ft.params.values())
    K = dim + 256
    rows = 0
    for r in ft.resnets:
        w = ft.params[r["pre"] + ".mlp.1.weight"]
        assert w.data_ptr() == ft.wcat.data_ptr() + rows * K * 4 and r["emb_off"] == rows
        rows += r["cout"]
    assert rows == ft.J
    x0, mask, cond, spk = case_inputs(B, T, [T, T - 5])
    ft.train_step(x0, mask, cond, torch.tensor([0.3, 0.7]), spk, torch.randn(x0.shape))
    c = collections.Counter(n for n, _ in trace)
    L, n_res, n_attn = len(mults), 4 * len(mults), 2 * len(mults)
    n_conv_params = sum(1 for k in ft.shapes if k.endswith(".weight") and len(ft.shapes[k]) == 4) - 3  # first conv pair, final 1x1
    assert c["usb_t_gn_apply"] == 2 * n_res and c["usb_t_gn_bwd"] == 2 * n_res + 1
    assert c["usb_t_attn_context"] == n_attn == c["usb_t_attn_bwd_small"] == c["usb_t_attn_bwd_dkv"]
    assert c["usb_t_first_conv"] == c["usb_t_first_conv_wgrad"] == c["usb_t_embed"] == c["usb_t_embed_bwd"] == 1
    assert c["usb_t_sumsq"] == c["usb_t_adam"] == 1
    # every conv parameter gets a weight gradient: skip-concat convs and to_qkv take two launches, to_out is handled by
    # usb_t_attn_bwd_small from the per-sample G matrix (one usb_t_wgrad per attention block)
    n_split = 2 * (L - 1)             # block1 conv + res_conv of each ups.k.0
    assert c["usb_t_wgrad"] == (n_conv_params - n_attn) + n_split + n_attn + n_attn
    names = [n for n, _ in trace]
    assert names.index("usb_t_loss") < names.index("usb_t_loss_grad") < names.index("usb_t_embed_bwd") < names.index("usb_t_adam")


def test_training_layout_round_trip_and_operand_order():
    """Conv master weights live in the forward operand layout; the conversion is exact, invertible and matches the order
    the kernels address (tap-major rows of contiguous input channels; ConvTranspose as four phase matrices)."""
    from unitspeech_b200.training import K1, K3S1, K3S2, KT4, to_reference_layout, to_train_layout
    g = torch.Generator().manual_seed(0)
    w = torch.randn(6, 4, 3, 3, generator=g)
    for kind in (K3S1, K3S2):
        t = to_train_layout(kind, w)
        assert torch.equal(to_reference_layout(kind, t.reshape(-1), w.shape), w)
        assert t.reshape(6, 9, 4)[2, 5, 3] == w[2, 3, 5 // 3, 5 % 3]
    w1 = torch.randn(5, 7, 1, 1, generator=g)
    assert torch.equal(to_reference_layout(K1, to_train_layout(K1, w1).reshape(-1), w1.shape), w1)
    wt = torch.randn(5, 7, 4, 4, generator=g)                       # ConvTranspose2d weight: (Cin, Cout, 4, 4)
    t = to_train_layout(KT4, wt)
    assert torch.equal(to_reference_layout(KT4, t.reshape(-1), wt.shape), wt)
    flat = t.reshape(-1)
    for ph in range(2):
        for pw in range(2):
            for a in range(2):
                for b in range(2):
                    kh = (1 if a == 0 else 3) if ph == 0 else (0 if a == 0 else 2)     # engine.cu pack_conv_host
                    kw = (1 if b == 0 else 3) if pw == 0 else (0 if b == 0 else 2)
                    assert flat[(((ph * 2 + pw) * 7 + 3) * 4 + (a * 2 + b)) * 5 + 2] == wt[2, 3, kh, kw]


def test_state_dict_round_trip_through_the_training_layout():
    trace = []
    ft = FineTuner(dim=64, dim_mults=(1, 2), _trace_calls=trace)
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=7, out_scale=1.0)
    ft.load_state_dict(p)
    back = ft.state_dict()
    assert list(back) == list(ft.shapes) and set(back) == set(p)
    assert all(torch.equal(back[k], p[k]) for k in p)
    with pytest.raises(KeyError):
        ft.load_state_dict({k: v for k, v in p.items() if k != "text_uncon"})


def test_every_trainable_parameter_receives_a_gradient_writer():
    """Graph-walk completeness without a GPU: in the recorded ABI call sequence of one backward pass, the gradient view
    of every parameter the objective depends on is handed to some kernel (directly, or as part of the stacked
    ResnetBlock.mlp buffers); only the CFG unconditionals (unused by loss_t, unitspeech.py:393-405) get none."""
    import ctypes
    trace = []
    ft = FineTuner(dim=128, dim_mults=(1, 2, 4, 8), _trace_calls=trace)
    ft.load_state_dict(O.harness_params(seed=3, out_scale=1.0))
    x0, mask, cond, spk = case_inputs(2, 16, [16, 11])
    ft.zero_grad()
    ft.forward(x0, mask, cond, torch.tensor([0.3, 0.7]), spk, torch.randn(x0.shape))
    n_fwd = len(trace)
    ft.backward()
    ptrs = set()
    for _, args in trace[n_fwd:]:
        for a in args:
            if isinstance(a, ctypes.c_void_p) and a.value:
                ptrs.add(a.value)
    stacked = [(ft.dwcat.data_ptr(), ft.dwcat.numel() * 4), (ft.dbcat.data_ptr(), ft.dbcat.numel() * 4)]
    assert ft.dwcat.data_ptr() in ptrs and ft.dbcat.data_ptr() in ptrs
    missing = []
    for k, g in ft.grads.items():
        a = g.data_ptr()
        if a in ptrs or any(lo <= a < lo + n for lo, n in stacked):
            continue
        missing.append(k)
    assert sorted(missing) == ["spk_uncon", "text_uncon"], missing
