"""The Diffusers drop-in flow of `implementations/diffusers_sdxl.py` (reference:
implementations/Diffusers/load_sdxl_pipeline.py:24-46) on the CPU, with a stand-in pipeline that touches the UNet the way
`StableDiffusionXLPipeline.__call__` does: reads `unet.config.*`, calls
`unet(x, t, encoder_hidden_states=..., cross_attention_kwargs=None, added_cond_kwargs=..., return_dict=False)[0]`
once per step on the CFG pair, and is called twice."""
import importlib.util
import os

import torch

import fake_kernels
from conftest import parity
from stabletriton_b200 import UNet2DConditionModel, UNetConfig, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _script():
    spec = importlib.util.spec_from_file_location("diffusers_sdxl", os.path.join(ROOT, "implementations", "diffusers_sdxl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class FakeSDXLPipeline:
    """The slice of StableDiffusionXLPipeline that involves the UNet (diffusers 0.21.2 pipeline_stable_diffusion_xl.py:
    prepare_latents / _get_add_time_ids read `unet.config`, the denoising loop calls the UNet with keyword arguments)."""

    def __init__(self, unet, cfg):
        self.unet, self.cfg = unet, cfg
        self.calls = 0

    def __call__(self, prompt_seed: int, num_inference_steps: int = 2, guidance_scale: float = 5.0):
        cfg = self.cfg
        h = w = 16
        assert self.unet.config.in_channels == cfg.in_channels                      # prepare_latents
        passed = self.unet.config.addition_time_embed_dim * cfg.num_time_ids + cfg.text_embed_dim
        assert passed == cfg.addition_time_embed_dim * cfg.num_time_ids + cfg.text_embed_dim   # _get_add_time_ids
        assert self.unet.config.sample_size > 0
        s = synth.synth_inputs(2, h, cfg, seed=prompt_seed)
        latents = synth.synth_tensor("latents", (1, cfg.in_channels, h, w), prompt_seed) * 3.0 ** 0.5
        for i in range(num_inference_steps):
            latent_model_input = torch.cat([latents] * 2)
            t = torch.tensor(999.0 - 400.0 * i)
            noise_pred = self.unet(
                latent_model_input, t, encoder_hidden_states=s["encoder_hidden_states"], cross_attention_kwargs=None,
                added_cond_kwargs=s["added_cond_kwargs"], return_dict=False)[0]
            self.calls += 1
            u, c = noise_pred.chunk(2)
            latents = latents - 0.1 * (u + guidance_scale * (c - u))
        return latents


def test_attach_engine_is_a_drop_in_for_pipe_unet():
    mod = _script()
    cfg = UNetConfig.tiny()
    ref_unet = synth.build_unet(cfg, seed=3, device="cpu", dtype=torch.float32)
    pipe = FakeSDXLPipeline(ref_unet, cfg)
    with torch.no_grad():
        want = [pipe(7), pipe(8)]

    # load_sdxl_pipeline.py:24-28 with this repo's model definition: strict load of the pipeline UNet's state dict
    fresh = UNet2DConditionModel(cfg)
    fresh.load_state_dict(ref_unet.state_dict(), strict=True)
    import stabletriton_b200 as st
    engine = st.optimize_model(fresh.eval(), cuda_graph=False, check_device=False)
    mod.attach_engine(pipe, engine)
    assert pipe.unet is engine and not [n for n in engine.graph.nodes if n.op == "call_module"]
    assert engine.config.in_channels == 4 and engine.config.addition_time_embed_dim == cfg.addition_time_embed_dim
    assert engine.config.sample_size == cfg.sample_size
    with torch.no_grad(), fake_kernels.installed() as calls:
        got = [pipe(7), pipe(8)]          # "call twice" (load_sdxl_pipeline.py:39,46)
    assert pipe.calls == 8 and calls["attention"] > 0
    for g, w in zip(got, want):
        rel, cos = parity(g, w)
        assert rel < 1e-5, (rel, cos)


def test_build_engine_unet_rejects_a_foreign_state_dict():
    mod = _script()
    import pytest

    class NotSDXL(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_in = torch.nn.Conv2d(4, 320, 3, padding=1)

    with pytest.raises(RuntimeError):  # strict load: missing / unexpected keys
        mod.build_engine_unet(NotSDXL(), device="meta", dtype=torch.float32, check_device=False, cuda_graph=False)
