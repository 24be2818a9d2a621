"""Pin the oracle (oracle/unet_oracle.py) and the model definition (stabletriton_b200/unet.py) against
the reference's own eager model: fixtures in tests/golden/ were produced by oracle/make_golden.py, which
imports /root/reference/src/stabletriton/optimizers/unet_pt.py and records its fp32 outputs.  Weights
are regenerated here from `stabletriton_b200.synth` (a pure function of parameter name + seed)."""
import importlib.util
import os

import pytest
import torch

from conftest import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 7  # oracle/make_golden.py:SEED


def _oracle():
    spec = importlib.util.spec_from_file_location("unet_oracle", os.path.join(ROOT, "oracle", "unet_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


O = _oracle()


def rnd(name, shape, scale=1.0):
    from stabletriton_b200 import synth
    return synth.synth_tensor("golden." + name, shape, SEED) * (3.0 ** 0.5) * scale


@pytest.fixture(scope="module")
def blocks(golden_dir):
    return torch.load(os.path.join(golden_dir, "blocks.pt"))


def _cfg():
    from stabletriton_b200 import UNetConfig
    # the reference hard-codes temb 1280, 32 groups, head_dim 64, cross-attention dim 2048
    return UNetConfig(block_out_channels=(320, 640, 1280), cross_attention_dim=2048, norm_num_groups=32)


def _mine(ctor):
    """(module of stabletriton_b200.unet with synthetic weights, its fp32 state dict)."""
    from stabletriton_b200 import synth
    with torch.device("meta"):
        m = ctor()
    sd = synth.synth_state_dict(m, seed=SEED)
    m.load_state_dict(sd, strict=True, assign=True)
    return m.eval(), sd


def _close(got, ref, tol=2e-5):
    rel, cos = parity(got, ref)
    assert got.shape == ref.shape
    assert rel <= tol and cos >= 1 - 1e-6, (rel, cos)


@torch.no_grad()
def test_resnet_blocks(blocks):
    from stabletriton_b200 import unet as U
    temb = rnd("temb", (2, 1280))
    for key, cin, cout, xname in (("resnet_shortcut", 64, 128, "resnet.x"), ("resnet_plain", 64, 64, "resnet2.x")):
        m, sd = _mine(lambda: U.ResnetBlock2D(cin, cout, 1280, 32))
        x = rnd(xname, (2, 64, 16, 16))
        _close(O.resnet_block(sd, x, temb, 32), blocks[key]["out"])
        _close(m(x, temb), blocks[key]["out"])


@torch.no_grad()
def test_attention_geglu_ff_transformer(blocks):
    from stabletriton_b200 import unet as U
    h = rnd("attn.x", (2, 96, 128))
    ctx = rnd("ctx", (2, 77, 2048))
    m, sd = _mine(lambda: U.Attention(128, None, 64))
    _close(O.attention(sd, h, None, 64), blocks["attention_self"]["out"])
    _close(m(h), blocks["attention_self"]["out"])
    m, sd = _mine(lambda: U.Attention(128, 2048, 64))
    _close(O.attention(sd, h, ctx, 64), blocks["attention_cross"]["out"])
    _close(m(h, ctx), blocks["attention_cross"]["out"])
    m, sd = _mine(lambda: U.GEGLU(128, 512))
    s, g = torch.nn.functional.linear(h, sd["proj.weight"], sd["proj.bias"]).chunk(2, dim=-1)
    _close(O.geglu(s, g), blocks["geglu"]["out"])
    _close(m(h), blocks["geglu"]["out"])
    m, sd = _mine(lambda: U.FeedForward(128))
    _close(O.feed_forward(sd, h), blocks["feed_forward"]["out"])
    _close(m(h), blocks["feed_forward"]["out"])
    m, sd = _mine(lambda: U.BasicTransformerBlock(128, 2048, 64))
    _close(O.transformer_block(sd, h, ctx, 64), blocks["transformer_block"]["out"])
    _close(m(h, ctx), blocks["transformer_block"]["out"])
    x = rnd("tf2d.x", (2, 128, 8, 12))
    m, sd = _mine(lambda: U.Transformer2DModel(128, 2, 2048, 64, 32))
    _close(O.transformer_2d(sd, x, ctx, 32, 64), blocks["transformer_2d"]["out"])
    _close(m(x, ctx), blocks["transformer_2d"]["out"])


@torch.no_grad()
def test_embeddings_and_samplers(blocks):
    from stabletriton_b200 import unet as U
    t = torch.tensor([999.0, 1.0, 500.0])
    _close(O.timesteps_embedding(t, 320), blocks["timesteps_320"]["out"], tol=1e-6)
    _close(O.timesteps_embedding(t, 256), blocks["timesteps_256"]["out"], tol=1e-6)
    _close(U.Timesteps(320)(t), blocks["timesteps_320"]["out"], tol=1e-6)
    e = rnd("temb_in", (2, 320))
    m, sd = _mine(lambda: U.TimestepEmbedding(320, 1280))
    _close(O.timestep_mlp(sd, e), blocks["timestep_embedding"]["out"])
    _close(m(e), blocks["timestep_embedding"]["out"])
    x = rnd("samp.x", (2, 64, 16, 16))
    m, sd = _mine(lambda: U.Downsample2D(64))
    _close(m(x), blocks["downsample"]["out"])
    m, sd = _mine(lambda: U.Upsample2D(64))
    _close(m(x), blocks["upsample"]["out"])


@torch.no_grad()
def test_down_up_mid_blocks(blocks):
    from stabletriton_b200 import unet as U
    from stabletriton_b200 import UNetConfig
    cfg = _cfg()
    temb = rnd("temb", (2, 1280))
    ctx = rnd("ctx", (2, 77, 2048))
    x = rnd("cadb.x", (2, 64, 16, 16))

    m, sd = _mine(lambda: U.DownBlock(cfg, 64, 128, 1, True))
    hs, outs = O.down_block(sd, x, temb, ctx, 32, 64)
    _close(hs, blocks["cross_attn_down_block"]["out"])
    for a, b in zip(outs, blocks["cross_attn_down_block"]["outs"]):
        _close(a, b)
    hs2, outs2 = m(x, temb, ctx)
    _close(hs2, blocks["cross_attn_down_block"]["out"])
    assert len(outs2) == len(blocks["cross_attn_down_block"]["outs"]) == 3

    m, sd = _mine(lambda: U.DownBlock(cfg, 64, 64, 0, True))
    hs, outs = O.down_block(sd, x, temb, None, 32, 64)
    _close(hs, blocks["down_block"]["out"])
    _close(m(x, temb)[0], blocks["down_block"]["out"])

    x = rnd("caub.x", (2, 128, 8, 8))
    skips = [rnd(f"caub.s{i}", (2, c, 8, 8)) for i, c in enumerate((64, 128, 128))]
    m, sd = _mine(lambda: U.UpBlock(cfg, 64, 128, 128, 1, True))
    _close(O.up_block(sd, x, skips, temb, ctx, 32, 64), blocks["cross_attn_up_block"]["out"])
    _close(m(x, list(skips), temb, ctx), blocks["cross_attn_up_block"]["out"])
    skips = [rnd(f"ub.s{i}", (2, 64, 8, 8)) for i in range(3)]
    m, sd = _mine(lambda: U.UpBlock(cfg, 64, 64, 128, 0, False))
    _close(O.up_block(sd, x, skips, temb, None, 32, 64), blocks["up_block"]["out"])
    _close(m(x, list(skips), temb), blocks["up_block"]["out"])

    x = rnd("mid.x", (1, 64, 8, 8))
    mid_cfg = UNetConfig(block_out_channels=(320, 640, 64), transformer_layers_per_block=(0, 2, 10))
    m, sd = _mine(lambda: U.UNetMidBlock2DCrossAttn(mid_cfg, 64, 10))
    _close(O.mid_block(sd, x, temb[:1], ctx[:1], 32, 64), blocks["mid_block"]["out"])
    _close(m(x, temb[:1], ctx[:1]), blocks["mid_block"]["out"])


@torch.no_grad()
def test_oracle_matches_model_definition_tiny():
    """Whole UNet, tiny config: oracle.unet_forward == stabletriton_b200.unet.UNet2DConditionModel (fp32)."""
    from stabletriton_b200 import UNetConfig, synth
    cfg = UNetConfig.tiny()
    model = synth.build_unet(cfg, seed=5, device="cpu", dtype=torch.float32)
    inp = synth.synth_inputs(2, 16, cfg, seed=9)
    ref = model(**inp)[0]
    got = O.unet_forward(model.state_dict(), inp["sample"], inp["timesteps"], inp["encoder_hidden_states"],
                         inp["added_cond_kwargs"], groups=cfg.norm_num_groups, head_dim=cfg.attention_head_dim,
                         addition_time_embed_dim=cfg.addition_time_embed_dim)[0]
    _close(got, ref)


def test_state_dict_keys_are_diffusers_sdxl():
    """1680 tensors / 2 567 463 684 parameters, the SDXL-base UNet."""
    from stabletriton_b200 import UNet2DConditionModel, UNetConfig
    with torch.device("meta"):
        m = UNet2DConditionModel(UNetConfig.sdxl())
    sd = m.state_dict()
    assert len(sd) == 1680
    assert sum(p.numel() for p in m.parameters()) == 2_567_463_684
    for key in ("conv_in.weight", "time_embedding.linear_1.weight", "add_embedding.linear_2.bias",
                "down_blocks.1.attentions.0.transformer_blocks.1.attn2.to_k.weight",
                "down_blocks.0.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_out.bias",
                "up_blocks.0.attentions.2.transformer_blocks.9.ff.net.0.proj.weight",
                "up_blocks.1.upsamplers.0.conv.bias", "up_blocks.2.resnets.2.conv_shortcut.weight", "conv_out.bias"):
        assert key in sd, key
    assert sd["down_blocks.2.attentions.0.transformer_blocks.0.attn2.to_k.weight"].shape == (1280, 2048)
    assert sd["add_embedding.linear_1.weight"].shape == (1280, 2816)
    assert sd["up_blocks.0.resnets.2.conv1.weight"].shape == (1280, 1920, 3, 3)


def test_euler_schedule_matches_oracle():
    from stabletriton_b200.pipeline import euler_schedule
    for n in (30, 50, 8):
        t, s, init = euler_schedule(n)
        to, so, inito = O.euler_sigmas(n)
        assert torch.equal(t, to)
        assert torch.allclose(s, so, rtol=1e-6, atol=1e-7)
        assert abs(init - inito) < 1e-6 * inito
    t, s, _ = euler_schedule(30)
    assert t[0].item() == 958.0 and t[-1].item() == 1.0 and s[-1].item() == 0.0
    assert 11.0 < s[0].item() < 12.0  # sigma at t = 958 (sigma_max = 14.6 at t = 999)


@pytest.mark.slow
@torch.no_grad()
def test_oracle_whole_unet_matches_reference_golden(golden_dir):
    """Full SDXL UNet, BASELINE config 1 (B=1, 4x64x64, t=999), fp32: oracle and model definition vs the
    reference's own output.  ~2 minutes on 8 cores (10 GB of synthetic weights)."""
    from stabletriton_b200 import UNet2DConditionModel, UNetConfig, synth
    fx = torch.load(os.path.join(golden_dir, "unet_sdxl_b1_64.pt"))
    cfg = UNetConfig.sdxl()
    with torch.device("meta"):
        m = UNet2DConditionModel(cfg)
    sd = synth.synth_state_dict(m, seed=fx["weight_seed"])
    inp = synth.synth_inputs(fx["batch"], fx["latent"], cfg, seed=fx["input_seed"], timestep=fx["timestep"])
    got = O.unet_forward(sd, inp["sample"], inp["timesteps"], inp["encoder_hidden_states"], inp["added_cond_kwargs"])[0]
    rel, cos = parity(got, fx["out"])
    print(f"oracle vs reference (fp32, full SDXL): rel={rel:.3e} cos={cos:.9f}")
    assert rel <= 1e-4 and cos >= 1 - 1e-7
    m.load_state_dict(sd, strict=True, assign=True)
    got2 = m.eval()(**inp)[0]
    rel, cos = parity(got2, fx["out"])
    assert rel <= 1e-4 and cos >= 1 - 1e-7
