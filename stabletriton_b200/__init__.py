"""stabletriton_b200 -- a B200-native (sm_100a) SDXL UNet denoise engine behind StableTriton's drop-in
surface: `model = stabletriton_b200.compile(model)`.

Layout:
    optimization.py   compile / optimize_model / replace_backend        (reference: optimization.py)
    fx_passes.py      torch.fx rewrite passes                            (reference: optimizers/*.py, utils/)
    wrappers.py       fx.wrap'ed op seam                                 (reference: *_wrapper functions)
    kernels.py        tensor-level launchers over the C ABI              (reference: kernels/*.py)
    cuda_graphs.py    signature-keyed CUDA-graph replay                  (reference: optimizers/cuda/graphs.py)
    unet.py           SDXL UNet definition (Diffusers keys)              (reference: optimizers/unet_pt.py)
    pipeline.py       Euler + CFG denoise loop, data-parallel launcher   (reference: implementations/Diffusers)
    weights.py        Diffusers safetensors / config.json I/O            (reference: load_sdxl_pipeline.py:24-35)
    vae.py            VAE decode on the same kernels (Diffusers AutoencoderKL decoder)
    comfy.py          ComfyUI call convention + LDM <-> Diffusers UNet key renaming  (reference: README.md:5)
    csrc/             CUDA kernels + C ABI (include/stabletriton_b200.h)
"""
from .optimization import compile, optimize_model, replace_backend, run_compiler  # noqa: F401
from .unet import UNet2DConditionModel, UNetConfig  # noqa: F401
from .weights import load_diffusers_unet, save_diffusers_unet  # noqa: F401
from .vae import AutoencoderKLDecoder, VAEConfig, build_vae_decoder, compile_vae  # noqa: F401

__version__ = "0.1.0"
