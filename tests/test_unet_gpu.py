"""Whole-UNet parity on the GPU: compile() -> every hot op through the sm_100a kernels -> compare with the
fp32 oracle (tiny config, oracle run on the CPU here) and with the reference's own output (full SDXL,
golden fixture produced by oracle/make_golden.py from /root/reference's eager model).

Tolerance (BASELINE.json north_star): per-step UNet output vs fp32 eager: max|d|/max|ref| <= 2e-2 and
cosine >= 0.9999.
"""
import os

import pytest
import torch

from conftest import parity

pytestmark = pytest.mark.gpu

REL_TOL = 2e-2
COS_TOL = 0.9999


def _oracle():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("unet_oracle", os.path.join(root, "oracle", "unet_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _to_dev(inp, device, dtype):
    return dict(
        sample=inp["sample"].to(device, dtype),
        timesteps=inp["timesteps"].to(device),
        encoder_hidden_states=inp["encoder_hidden_states"].to(device, dtype),
        added_cond_kwargs={k: v.to(device, dtype) for k, v in inp["added_cond_kwargs"].items()},
    )


def _build(cfg, seed, device="cuda"):
    from stabletriton_b200 import synth
    return synth.build_unet(cfg, seed=seed, device=device, dtype=torch.bfloat16)


@pytest.mark.parametrize("batch,latent", [(2, 32), (1, 16), (16, 16)])  # 16 = 8 prompts x CFG
def test_tiny_unet_matches_oracle(built_lib, batch, latent):
    import stabletriton_b200 as st
    from stabletriton_b200 import UNetConfig, synth, _cabi

    cfg = UNetConfig.tiny()
    model = _build(cfg, seed=3)
    inp = synth.synth_inputs(batch, latent, cfg, seed=11)
    # the oracle sees the same bf16-rounded weights and inputs, in fp32
    sd32 = {k: v.float().cpu() for k, v in model.state_dict().items()}
    inp32 = _to_dev(_to_dev(inp, "cpu", torch.bfloat16), "cpu", torch.float32)
    ref = _oracle().unet_forward(sd32, inp32["sample"], inp32["timesteps"], inp32["encoder_hidden_states"],
                                 inp32["added_cond_kwargs"], groups=cfg.norm_num_groups,
                                 head_dim=cfg.attention_head_dim, addition_time_embed_dim=cfg.addition_time_embed_dim)[0]

    compiled = st.compile(model, cuda_graph=False)
    left = [n for n in compiled.graph.nodes if n.op == "call_module"]
    assert not left, f"modules left un-replaced: {left[:5]}"
    _cabi.reset_launch_count()
    with torch.no_grad():
        out = compiled(**_to_dev(inp, "cuda", torch.bfloat16))[0]
    torch.cuda.synchronize()
    assert _cabi.launch_count() > 100, "the CUDA kernels did not run"
    assert out.shape == ref.shape and out.dtype == torch.bfloat16
    rel, cos = parity(out.float(), ref)
    print(f"tiny B={batch} {latent}x{latent}: rel={rel:.3e} cos={cos:.6f} launches={_cabi.launch_count()}")
    assert rel <= REL_TOL and cos >= COS_TOL, (rel, cos)


def test_tiny_unet_cuda_graph_replay_is_deterministic(built_lib):
    import stabletriton_b200 as st
    from stabletriton_b200 import UNetConfig, synth

    cfg = UNetConfig.tiny()
    model = _build(cfg, seed=3)
    compiled = st.compile(model, cuda_graph=True)
    a = _to_dev(synth.synth_inputs(2, 32, cfg, seed=11), "cuda", torch.bfloat16)
    b = _to_dev(synth.synth_inputs(2, 32, cfg, seed=12), "cuda", torch.bfloat16)
    eager_a = compiled.eager_forward(**a)[0].clone()
    out_a1 = compiled(**a)[0]
    out_b = compiled(**b)[0]
    out_a2 = compiled(**a)[0]
    torch.cuda.synchronize()
    assert len(compiled.forward._cached) == 1, "same signature must reuse one captured graph"
    assert torch.equal(out_a1, out_a2), "replay with the same inputs must be bit-identical"
    assert torch.equal(out_a1, eager_a), "graph replay must equal the eager launch sequence"
    assert not torch.equal(out_a1, out_b)
    # a new signature (batch 1) captures a second graph
    c = _to_dev(synth.synth_inputs(1, 32, cfg, seed=11), "cuda", torch.bfloat16)
    compiled(**c)
    assert len(compiled.forward._cached) == 2


def test_sdxl_unet_matches_reference_golden(built_lib, golden_dir):
    """Full-size SDXL UNet, BASELINE config 1 inputs (B=1, 4x64x64, t=999): engine (bf16) vs the
    reference's own fp32 eager output on identical synthetic weights."""
    import stabletriton_b200 as st
    from stabletriton_b200 import UNetConfig, synth, _cabi

    fx = torch.load(os.path.join(golden_dir, "unet_sdxl_b1_64.pt"))
    cfg = UNetConfig.sdxl()
    model = _build(cfg, seed=fx["weight_seed"])
    compiled = st.compile(model, cuda_graph=True)
    inp = _to_dev(synth.synth_inputs(fx["batch"], fx["latent"], cfg, seed=fx["input_seed"], timestep=fx["timestep"]),
                  "cuda", torch.bfloat16)
    _cabi.reset_launch_count()
    out = compiled(**inp)[0]
    torch.cuda.synchronize()
    rel, cos = parity(out.float(), fx["out"])
    print(f"SDXL B=1 64x64 vs reference fp32: rel={rel:.3e} cos={cos:.6f} launches(capture+warmup)={_cabi.launch_count()}")
    print("pass report:", compiled.pass_report)
    assert rel <= REL_TOL and cos >= COS_TOL, (rel, cos)


def test_prepare_plus_step_forward_equals_forward(built_lib):
    """Prompt-constant hoisting on the GPU: prepare() + step_forward() is bit-identical to forward()."""
    import stabletriton_b200 as st
    from stabletriton_b200 import UNetConfig, synth

    cfg = UNetConfig.tiny()
    compiled = st.compile(_build(cfg, seed=3), cuda_graph=False)
    inp = _to_dev(synth.synth_inputs(2, 32, cfg, seed=11), "cuda", torch.bfloat16)
    with torch.no_grad():
        ref = compiled(**inp)[0]
        consts = compiled.prepare(inp["encoder_hidden_states"], inp["added_cond_kwargs"])
        assert len(consts) == compiled.num_prompt_constants == 35
        out = compiled.step_forward(inp["sample"], inp["timesteps"], *consts)[0]
        again = compiled.step_forward(inp["sample"] * 0.5, torch.tensor(17.0, device="cuda"), *consts)[0]
        ref2 = compiled(inp["sample"] * 0.5, torch.tensor(17.0, device="cuda"), inp["encoder_hidden_states"],
                        inp["added_cond_kwargs"])[0]
    assert torch.equal(out, ref) and torch.equal(again, ref2)


def test_checkpoint_round_trip_through_the_engine(built_lib, tmp_path):
    """Diffusers-layout safetensors -> load_diffusers_unet -> compile(): same output as the model that wrote the file."""
    import stabletriton_b200 as st
    from stabletriton_b200 import UNetConfig, synth

    cfg = UNetConfig.tiny()
    model = _build(cfg, seed=9)
    st.save_diffusers_unet(model, os.path.join(tmp_path, "unet"), cfg)
    loaded = st.load_diffusers_unet(os.path.join(tmp_path, "unet"))
    assert next(loaded.parameters()).dtype == torch.bfloat16 and next(loaded.parameters()).is_cuda
    inp = _to_dev(synth.synth_inputs(2, 16, cfg, seed=5), "cuda", torch.bfloat16)
    with torch.no_grad():
        a = st.compile(model, cuda_graph=False)(**inp)[0]
        b = st.compile(loaded, cuda_graph=False)(**inp)[0]
    assert torch.equal(a, b)
