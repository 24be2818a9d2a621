#!/usr/bin/env python
"""SDXL through Diffusers with the B200 engine as `pipe.unet` -- the counterpart of the reference's
`implementations/Diffusers/load_sdxl_pipeline.py:17-46`:

    reference                                                  here
    ---------------------------------------------------------  ------------------------------------------------------
    pipe = DiffusionPipeline.from_pretrained(..., fp16)  :17   same call, torch_dtype=bfloat16
    unet_new = UNet2DConditionModelPT().half().cuda()    :24   st.UNet2DConditionModel().to(bfloat16).cuda()
    unet_new.load_state_dict(pipe.unet.state_dict())     :25   same (strict: identical 1680 Diffusers keys)
    unet_new = optimize_model(unet_new, cuda_graph=True) :28   st.compile(unet_new, cuda_graph=True)
    unet_new.config = namedtuple(...) (manual patch)  :29-34   not needed: compile() carries `.config` over
    pipe.unet = unet_new                                 :35   same
    image = pipe(prompt).images[0]  (twice)           :39,46   same: the first call warms up + captures the CUDA graph

Diffusers / a checkpoint are not part of this repository's environment (no network), so the imports are guarded and the
steps are functions: `tests/test_integration.py` drives `attach_engine` with a stand-in pipeline object that calls the
UNet exactly the way `StableDiffusionXLPipeline.__call__` does.

    python implementations/diffusers_sdxl.py [--model stabilityai/stable-diffusion-xl-base-1.0 | /path/to/pipeline]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.setrecursionlimit(10000)  # deep fx graphs (load_sdxl_pipeline.py:12)

import stabletriton_b200 as st  # noqa: E402


def build_engine_unet(reference_unet: torch.nn.Module, device="cuda", dtype=torch.bfloat16, *,
                      check_device: bool = True, cuda_graph: bool = True):
    """load_sdxl_pipeline.py:24-28: a fresh UNet of this repo's definition, strict `load_state_dict` of the pipeline's
    UNet weights, then compile().  `reference_unet` only has to provide `state_dict()` with the Diffusers keys."""
    unet = st.UNet2DConditionModel().to(dtype).to(device)
    state = {k: v.to(device=device, dtype=dtype) for k, v in reference_unet.state_dict().items()}
    unet.load_state_dict(state, strict=True)
    unet.eval().requires_grad_(False)
    return st.optimize_model(unet, cuda_graph=cuda_graph, check_device=check_device)


def attach_engine(pipe, engine_unet=None, **build_kwargs):
    """load_sdxl_pipeline.py:24-35: swap the pipeline's UNet for the compiled engine.  The pipeline keeps calling
    `pipe.unet(latent_model_input, t, encoder_hidden_states=..., cross_attention_kwargs=None,
    added_cond_kwargs={...}, return_dict=False)[0]` and reading `pipe.unet.config.{in_channels,
    addition_time_embed_dim, sample_size}`; both work on the compiled module without a patch."""
    if engine_unet is None:
        engine_unet = build_engine_unet(pipe.unet, **build_kwargs)
    for field in ("in_channels", "addition_time_embed_dim", "sample_size"):
        assert hasattr(engine_unet.config, field), f"compiled UNet lost config.{field}"
    pipe.unet = engine_unet
    return pipe


def attach_vae(pipe, device="cuda", dtype=torch.bfloat16):
    """Optional: decode on the B200 kernels too (the reference leaves the pipeline's eager VAE in place).
    The pipeline keeps calling `pipe.vae.decode(latents / scaling_factor, return_dict=False)[0]`."""
    from stabletriton_b200.vae import AutoencoderKLDecoder, VAEConfig, compile_vae

    cfg = VAEConfig(scaling_factor=float(pipe.vae.config.scaling_factor))
    dec = AutoencoderKLDecoder(cfg).to(dtype).to(device)
    dec.load_state_dict({k: v.to(device=device, dtype=dtype) for k, v in pipe.vae.state_dict().items()
                         if k.startswith(("decoder.", "post_quant_conv."))}, strict=True)
    engine = compile_vae(dec.eval().requires_grad_(False))

    def decode(z, return_dict: bool = True, **_):
        image = engine.decode(z.to(dtype), pre_scaled=True)
        if not return_dict:
            return (image,)
        from diffusers.models.autoencoders.vae import DecoderOutput  # only reached inside a Diffusers process
        return DecoderOutput(sample=image)

    pipe.vae.decode = decode
    return pipe


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="stabilityai/stable-diffusion-xl-base-1.0")
    ap.add_argument("--prompt", default="a photo of an astronaut riding a horse on mars")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--out", default="sdxl_b200.png")
    ap.add_argument("--engine-vae", action="store_true", help="also run the VAE decode on the B200 kernels")
    args = ap.parse_args()
    try:
        from diffusers import DiffusionPipeline  # third party; not installed in the build image
    except ImportError as e:
        raise SystemExit(f"diffusers is not installed ({e}); this script is the Diffusers integration -- see "
                         f"tests/test_integration.py for the same flow with a stand-in pipeline") from None

    pipe = DiffusionPipeline.from_pretrained(args.model, torch_dtype=torch.bfloat16, use_safetensors=True).to("cuda")
    attach_engine(pipe)
    if args.engine_vae:
        attach_vae(pipe)

    image = pipe(args.prompt, num_inference_steps=args.steps).images[0]  # warm-up: kernels loaded, CUDA graph captured
    del image
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    image = pipe(args.prompt, num_inference_steps=args.steps).images[0]  # load_sdxl_pipeline.py:46: the measured call
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{args.steps} steps in {dt:.2f} s = {args.steps / dt:.2f} it/s (whole pipeline: text encoders + UNet + VAE)")
    image.save(args.out)


if __name__ == "__main__":
    main()
