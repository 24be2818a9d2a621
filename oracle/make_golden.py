"""Generate tests/golden/*.pt by running the REFERENCE's own eager model (test infrastructure only).

Run in the build container, where /root/reference is mounted:

    python oracle/make_golden.py            # block-level fixtures + whole-UNet (config 1) fixture

The reference's `optimizers/unet_pt.py` depends on torch only and is imported from where it lies (it
is never copied into this repository).  Weights are not stored: both this script and the tests fill
parameters with `stabletriton_b200.synth` (a pure function of parameter name and seed), so a fixture
holds only the inputs' recipe and the reference's fp32 output.

Fixtures:
  blocks.pt   reference building blocks at reduced sizes (ResnetBlock2D with and without shortcut,
              Attention self/cross, GEGLU, FeedForward, BasicTransformerBlock, Transformer2DModel,
              Timesteps, TimestepEmbedding, Down/Upsample2D, CrossAttnDownBlock2D, CrossAttnUpBlock2D,
              UpBlock2D, UNetMidBlock2DCrossAttn) -> pins every oracle function.
  unet_sdxl_b1_64.pt  whole reference UNet2DConditionModel, BASELINE config 1 (B=1, 4x64x64 latent,
              fp32, t=999) -> pins `unet_forward` and the engine's full-size parity test.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabletriton_b200 import synth  # noqa: E402
from stabletriton_b200.unet import UNetConfig  # noqa: E402

REF_FILE = "/root/reference/src/stabletriton/optimizers/unet_pt.py"
OUT_DIR = os.path.join(ROOT, "tests", "golden")
SEED = 7


def load_reference():
    spec = importlib.util.spec_from_file_location("reference_unet_pt", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def rnd(name, shape, scale=1.0):
    return synth.synth_tensor("golden." + name, shape, SEED) * (3.0 ** 0.5) * scale


def build(ctor, *args, **kwargs):
    with torch.device("meta"):
        m = ctor(*args, **kwargs)
    sd = synth.synth_state_dict(m, seed=SEED)
    m = m.to_empty(device="cpu")
    m.load_state_dict(sd, strict=True)
    return m.eval()


@torch.no_grad()
def block_fixtures(ref):
    fx = {}
    temb = rnd("temb", (2, 1280))
    ctx = rnd("ctx", (2, 77, 2048))

    x = rnd("resnet.x", (2, 64, 16, 16))
    fx["resnet_shortcut"] = dict(ctor=("ResnetBlock2D", (64, 128), {}), out=build(ref.ResnetBlock2D, 64, 128)(x, temb))
    x = rnd("resnet2.x", (2, 64, 16, 16))
    fx["resnet_plain"] = dict(ctor=("ResnetBlock2D", (64, 64), {"conv_shortcut": False}),
                              out=build(ref.ResnetBlock2D, 64, 64, conv_shortcut=False)(x, temb))

    h = rnd("attn.x", (2, 96, 128))
    fx["attention_self"] = dict(ctor=("Attention", (128,), {}), out=build(ref.Attention, 128)(h))
    fx["attention_cross"] = dict(ctor=("Attention", (128, 2048), {}), out=build(ref.Attention, 128, 2048)(h, ctx))
    fx["geglu"] = dict(ctor=("GEGLU", (128, 512), {}), out=build(ref.GEGLU, 128, 512)(h))
    fx["feed_forward"] = dict(ctor=("FeedForward", (128, 128), {}), out=build(ref.FeedForward, 128, 128)(h))
    fx["transformer_block"] = dict(ctor=("BasicTransformerBlock", (128,), {}),
                                   out=build(ref.BasicTransformerBlock, 128)(h, ctx))

    x = rnd("tf2d.x", (2, 128, 8, 12))
    fx["transformer_2d"] = dict(ctor=("Transformer2DModel", (128, 128, 2), {}),
                                out=build(ref.Transformer2DModel, 128, 128, 2)(x, ctx))

    t = torch.tensor([999.0, 1.0, 500.0])
    fx["timesteps_320"] = dict(out=ref.Timesteps(320)(t))
    fx["timesteps_256"] = dict(out=ref.Timesteps(256)(t))
    e = rnd("temb_in", (2, 320))
    fx["timestep_embedding"] = dict(ctor=("TimestepEmbedding", (320, 1280), {}),
                                    out=build(ref.TimestepEmbedding, 320, 1280)(e))

    x = rnd("samp.x", (2, 64, 16, 16))
    fx["downsample"] = dict(ctor=("Downsample2D", (64, 64), {}), out=build(ref.Downsample2D, 64, 64)(x))
    fx["upsample"] = dict(ctor=("Upsample2D", (64, 64), {}), out=build(ref.Upsample2D, 64, 64)(x))

    x = rnd("cadb.x", (2, 64, 16, 16))
    hs, outs = build(ref.CrossAttnDownBlock2D, 64, 128, 1)(x, temb, ctx)
    fx["cross_attn_down_block"] = dict(ctor=("CrossAttnDownBlock2D", (64, 128, 1), {}), out=hs, outs=outs)
    hs, outs = build(ref.DownBlock2D, 64, 64)(x, temb)
    fx["down_block"] = dict(ctor=("DownBlock2D", (64, 64), {}), out=hs, outs=outs)

    x = rnd("caub.x", (2, 128, 8, 8))
    skips = [rnd(f"caub.s{i}", (2, c, 8, 8)) for i, c in enumerate((64, 128, 128))]
    fx["cross_attn_up_block"] = dict(
        ctor=("CrossAttnUpBlock2D", (), dict(in_channels=64, out_channels=128, prev_output_channel=128, n_layers=1)),
        out=build(ref.CrossAttnUpBlock2D, in_channels=64, out_channels=128, prev_output_channel=128, n_layers=1)(
            x, list(skips), temb, ctx))
    skips = [rnd(f"ub.s{i}", (2, 64, 8, 8)) for i in range(3)]
    fx["up_block"] = dict(
        ctor=("UpBlock2D", (), dict(in_channels=64, out_channels=64, prev_output_channel=128)),
        out=build(ref.UpBlock2D, in_channels=64, out_channels=64, prev_output_channel=128)(x, list(skips), temb))

    x = rnd("mid.x", (1, 64, 8, 8))
    mid = build(ref.UNetMidBlock2DCrossAttn, 64)
    fx["mid_block"] = dict(ctor=("UNetMidBlock2DCrossAttn", (64,), {}), out=mid(x, temb[:1], ctx[:1]))
    return fx


@torch.no_grad()
def unet_fixture(ref):
    t0 = time.time()
    model = build(ref.UNet2DConditionModel)
    print(f"reference UNet built + filled in {time.time() - t0:.1f}s", flush=True)
    cfg = UNetConfig.sdxl()
    inp = synth.synth_inputs(1, 64, cfg, seed=1234)
    t0 = time.time()
    out = model(inp["sample"], inp["timesteps"], inp["encoder_hidden_states"], inp["added_cond_kwargs"])[0]
    print(f"reference forward {time.time() - t0:.1f}s  mean {out.mean():.4f} std {out.std():.4f} "
          f"absmax {out.abs().max():.4f}", flush=True)
    return dict(batch=1, latent=64, input_seed=1234, weight_seed=SEED, timestep=999.0, out=out.clone())


def main():
    if not os.path.exists(REF_FILE):
        raise SystemExit(f"{REF_FILE} not found: golden fixtures can only be generated where the reference is mounted")
    os.makedirs(OUT_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = load_reference()
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "blocks"):
        fx = block_fixtures(ref)
        torch.save(fx, os.path.join(OUT_DIR, "blocks.pt"))
        print("blocks.pt:", {k: tuple(v["out"].shape) for k, v in fx.items()})
    if what in ("all", "unet"):
        torch.save(unet_fixture(ref), os.path.join(OUT_DIR, "unet_sdxl_b1_64.pt"))
    for f in sorted(os.listdir(OUT_DIR)):
        print(f, os.path.getsize(os.path.join(OUT_DIR, f)), "bytes")


if __name__ == "__main__":
    main()
