"""Graph surgery on the CPU: pass match counts, and semantics of the rewritten graph (with the eager
test double of tests/fake_kernels.py standing in for the CUDA kernels)."""
import importlib.util
import os
import sys

import pytest
import torch
import torch.nn as nn

import fake_kernels
from conftest import parity
from stabletriton_b200 import UNet2DConditionModel, UNetConfig, optimize_model, synth
from stabletriton_b200 import fx_passes as P
from stabletriton_b200 import wrappers as W

REF_FILE = "/root/reference/src/stabletriton/optimizers/unet_pt.py"

# per UNet forward, SURVEY section 3.1 / 8a
SDXL_COUNTS = {"remove_dropout": 227, "fuse_attention": 140, "fuse_linear_geglu": 70,
               "replace_group_norm_activation": 35, "replace_group_norm": 11, "replace_layer_norm": 210}


def _compiled_tiny(seed=3):
    cfg = UNetConfig.tiny()
    model = synth.build_unet(cfg, seed=seed, device="cpu", dtype=torch.float32)
    return cfg, model, optimize_model(model, cuda_graph=False, check_device=False)


def test_tiny_rewrite_leaves_no_module_calls_and_preserves_output():
    cfg, model, gm = _compiled_tiny()
    left = P.census(gm)
    assert not [k for k in left if k.startswith("module:")], left
    assert left["attention_wrapper"] == 34 and left["linear_geglu_wrapper"] == 17
    assert gm.pass_report["fuse_shared_input_projections"] == 1  # all cross-attention K/V projections: one GEMM
    # 17 cross-attention layers (5 at 128 channels, 12 at 256): K and V rows stacked, context width 128
    assert gm.get_buffer("_st_shared_proj_0").shape == (2 * (5 * 128 + 12 * 256), 128)
    assert not [k for k, _ in gm.named_buffers() if k.startswith("_st_fused_proj_")
                and gm.get_buffer(k).shape[1] == 128 and gm.get_buffer(k).shape[0] in (256, 512)]
    # every GroupNorm takes its statistics from its producer's epilogue: 39 convs + 9 proj_out GEMMs emit them
    assert gm.pass_report["fuse_group_norm_statistics"] == 46
    assert left["group_norm_wrapper"] == 46 and left["concat_wrapper"] == 9
    assert left["conv2d_wrapper"] + left["conv2d_stats_wrapper"] == 51 and left["linear_stats_wrapper"] == 9
    inp = synth.synth_inputs(2, 16, cfg, seed=5)
    with torch.no_grad():
        ref = model(**inp)[0]
        with fake_kernels.installed() as calls:
            out = gm(**inp)[0]
    rel, cos = parity(out, ref)
    assert rel < 1e-5 and cos > 1 - 1e-9, (rel, cos)
    assert calls["attention"] == 34 and calls["conv2d"] == 51 and calls["groupnorm"] == 46
    # at a 16x16 latent only the top level (256 pixels per image) has whole 128-row tiles per image; the fake GroupNorm
    # checks every set of partials it receives against the statistics of its actual input
    assert calls["groupnorm_from_partials"] >= 10
    # keeps the forward signature and the Diffusers config shim
    assert gm.config.in_channels == 4 and gm.config.addition_time_embed_dim == cfg.addition_time_embed_dim
    assert isinstance(gm(**inp) if False else [out], list)


def test_prompt_constant_split_is_exact():
    """prepare() + step_forward() == forward(), bit for bit; the prologue holds exactly the prompt-only work."""
    cfg, model, gm = _compiled_tiny()
    pro, body, n = P.split_prompt_constants(gm)
    assert n == gm.num_prompt_constants == 1 + 2 * 17  # embedding MLP hidden + K and V of 17 cross-attention layers
    assert P.census(pro) == {"timestep_wrapper": 1, "linear_wrapper": 1, "linear_wrapper_functional": 1}
    assert [x.target for x in pro.graph.nodes if x.op == "placeholder"][:2] == ["encoder_hidden_states", "added_cond_kwargs"]
    body_inputs = [x.target for x in body.graph.nodes if x.op == "placeholder"]
    assert body_inputs[:2] == ["sample", "timesteps"] and len(body_inputs) == 2 + n
    inp = synth.synth_inputs(2, 16, cfg, seed=5)
    with torch.no_grad(), fake_kernels.installed():
        ref = gm(**inp)[0]
        consts = gm.prepare(inp["encoder_hidden_states"], inp["added_cond_kwargs"])
        out = gm.step_forward(inp["sample"], inp["timesteps"], *consts)[0]
        # a different step with the same prompt reuses the constants
        inp2 = dict(inp, sample=inp["sample"] * 0.5, timesteps=torch.tensor(17.0))
        ref2 = gm(**inp2)[0]
        out2 = gm.step_forward(inp2["sample"], inp2["timesteps"], *consts)[0]
    assert torch.equal(out, ref) and torch.equal(out2, ref2)


def test_pass_counts_on_sdxl_architecture():
    with torch.device("meta"):
        model = UNet2DConditionModel(UNetConfig.sdxl())
    gm = P.trace(model)
    report = {}
    from stabletriton_b200.optimization import replace_backend
    # qkv fusion concatenates real weights: skip it on the meta device
    orig = P.fuse_qkv_projection
    P.fuse_qkv_projection = lambda g: 0
    try:
        replace_backend(gm, report)
    finally:
        P.fuse_qkv_projection = orig
    for k, v in SDXL_COUNTS.items():
        assert report[k] == v, (k, report[k], v)
    assert report["fuse_conv_epilogues"] == 34 and report["replace_conv"] == 17  # 51 Conv2d
    assert report["fuse_proj_out_residual"] == 11 and report["replace_cat"] == 9 and report["replace_timesteps"] == 2
    census = P.census(gm)
    assert not [k for k in census if k.startswith("module:")], census
    # 743 Linears: 70 GEGLU, the 17 resnet time-embedding projections batched into one GEMM, the rest 1:1
    assert report["fuse_time_embedding_projections"] == 1 and census["linear_wrapper_functional"] == 1
    assert census["linear_geglu_wrapper"] + census["linear_wrapper"] + census["linear_stats_wrapper"] == 743 - 17
    # all 46 GroupNorms take their statistics from the epilogue of the conv / proj_out GEMM that wrote their input
    assert report["fuse_group_norm_statistics"] == 46 and census["linear_stats_wrapper"] == 9
    assert census["conv2d_stats_wrapper"] + census["conv2d_wrapper"] == 51


@pytest.mark.skipif(not os.path.exists(REF_FILE), reason="reference not mounted")
def test_passes_apply_to_the_reference_model_definition():
    """Drop-in: the same passes rewrite the reference's own unet_pt.UNet2DConditionModel."""
    sys.setrecursionlimit(10000)
    spec = importlib.util.spec_from_file_location("reference_unet_pt", REF_FILE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    with torch.device("meta"):
        model = ref.UNet2DConditionModel()
    gm = P.trace(model)
    assert P.remove_dropout(gm) == 227
    assert P.fuse_attention(gm) == 140
    assert P.fuse_linear_geglu(gm) == 70
    assert P.fuse_proj_out_residual(gm) == 11
    assert P.fuse_linear_residual(gm) == 211
    assert P.replace_linear_activ(gm) == 19
    assert P.replace_group_norm_activation(gm) == 35
    assert P.replace_group_norm(gm) == 11
    assert P.replace_layer_norm(gm) == 210
    P.replace_linear(gm)
    assert P.fuse_conv_epilogues(gm) == 34
    assert P.replace_conv(gm) == 17
    assert P.replace_cat(gm) == 9
    assert P.replace_timesteps(gm) == 2
    assert not [k for k in P.census(gm) if k.startswith("module:")]


class _Toy(nn.Module):
    """The reference's own toy (remove_dropout.py:8-20): Linear x3 -> SiLU -> Dropout."""

    def __init__(self):
        super().__init__()
        self.lin1, self.lin2, self.lin3 = nn.Linear(5, 5), nn.Linear(5, 5), nn.Linear(5, 5)
        self.nonlin, self.dropout = nn.SiLU(), nn.Dropout(p=0.0)

    def forward(self, x):
        return self.dropout(self.nonlin(self.lin3(self.lin2(self.lin1(x)))))


def test_toy_dropout_and_linear_activation_like_reference_selftests():
    m = _Toy().eval()
    gm = P.trace(m)
    before = gm.code
    assert P.remove_dropout(gm) == 1
    assert P.replace_linear_activ(gm, nn.SiLU()) == 1
    assert P.replace_linear(gm) == 2
    assert gm.code != before  # the reference's own success criterion
    x = torch.rand(5, 5)
    with fake_kernels.installed():
        assert (gm(x) - m(x)).abs().max() < 1e-6


def test_attention_pattern_binds_literals():
    from stabletriton_b200.unet import Attention
    class SelfAttn(nn.Module):  # call it the way BasicTransformerBlock does: no context argument
        def __init__(self):
            super().__init__()
            self.attn = Attention(128, None, 64)

        def forward(self, x):
            return self.attn(x)

    m = SelfAttn().eval()
    gm = P.trace(m)
    assert P.fuse_attention(gm) == 1
    node = next(n for n in gm.graph.nodes if n.op == "call_function" and n.target is W.attention_wrapper)
    assert node.args[3] is None and node.args[4] == 0.125 and node.args[5] == 2 and node.args[6] == 64
    assert P.fuse_qkv_projection(gm) == 1
    assert gm.get_buffer("_st_fused_proj_0").shape == (384, 128)
    assert "_st_fused_proj_0" not in gm.state_dict()  # non-persistent: Diffusers keys unchanged
    x = torch.randn(2, 10, 128)
    P.replace_linear(gm)
    with fake_kernels.installed(), torch.no_grad():
        assert (gm(x) - m(x)).abs().max() < 1e-5


def test_geglu_fallback_pattern_matches_reference_toy():
    class G(nn.Module):
        def forward(self, state, gate):
            return state * torch.nn.functional.gelu(gate)
    gm = P.trace(G())
    assert P.fuse_linear_geglu(gm) == 0 and P.fuse_geglu(gm) == 1
    a, b = torch.rand(5, 5), torch.rand(5, 5)
    with fake_kernels.installed():
        assert (gm(a, b) - a * torch.nn.functional.gelu(b)).abs().max() < 1e-6


def test_compile_refuses_cpu_models():
    import stabletriton_b200 as st
    model = synth.build_unet(UNetConfig.tiny(), seed=1, device="cpu", dtype=torch.float32)
    with pytest.raises(AssertionError):
        st.compile(model)


def test_fused_projection_weights_stay_live_after_compile():
    """the fused QKV / cross-attention K/V / time-embedding buffers must not freeze the weights at
    compile() time.  The parameters are views of the fused buffers, so an in-place update reaches the rewritten graph;
    a REPLACED parameter is picked up by refresh_fused_weights()."""
    cfg, model, gm = _compiled_tiny()
    inp = synth.synth_inputs(2, 16, cfg, seed=5)
    sources = gm._st_fused_sources
    assert any(k.startswith("_st_shared_proj_") for k in sources) and any(k.startswith("_st_temb_proj_w_") for k in sources)
    # every q/k/v, cross-attention K/V and time_emb_proj parameter aliases its rows of a fused buffer
    n_alias = 0
    for buffer, parts in sources.items():
        buf = gm.get_buffer(buffer)
        for qualname, attr, off, rows in parts:
            p = getattr(gm.get_submodule(qualname), attr)
            assert p.data_ptr() == buf[off:off + rows].data_ptr() and p.shape == buf[off:off + rows].shape
            n_alias += 1
    assert n_alias == 3 * 17 + 2 * 17 + 2 * len([m for m in gm.modules() if hasattr(m, "time_emb_proj")])

    def run(g):
        with torch.no_grad(), fake_kernels.installed():
            return g(**inp)[0]

    before = run(gm)
    targets = [n for n, _ in model.named_parameters()
               if n.endswith(("attn1.to_q.weight", "attn2.to_k.weight", "time_emb_proj.weight", "time_emb_proj.bias"))]
    assert len(targets) > 20
    with torch.no_grad():  # in-place update through the ORIGINAL model object (compile() shares its parameters)
        for name in targets:
            p = model.get_parameter(name)
            p.mul_(1.25).add_(0.01)
        ref = model(**inp)[0]
    after = run(gm)
    assert not torch.allclose(after, before)
    rel, cos = parity(after, ref)
    assert rel < 1e-5, (rel, cos)

    # replacing parameter tensors breaks the aliasing until refresh_fused_weights()
    sd = {k: v * 0.5 for k, v in model.state_dict().items()}
    model.load_state_dict(sd, strict=True, assign=True)
    assert gm.refresh_fused_weights() > 0
    with torch.no_grad():
        ref2 = model(**inp)[0]
    rel, cos = parity(run(gm), ref2)
    assert rel < 1e-5, (rel, cos)
    assert gm.refresh_fused_weights() == 0  # idempotent: everything aliases again
