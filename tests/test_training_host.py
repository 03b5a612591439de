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
