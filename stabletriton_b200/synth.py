"""Deterministic synthetic weights and inputs, bit-identical on CPU and CUDA.

BASELINE.json asks for random-init weights of the SDXL architecture (no checkpoints: there is no
network).  `torch.manual_seed` streams differ between CPU and CUDA generators and depend on module
construction order, so instead every tensor is filled from an integer hash of (seed, parameter name,
element index), evaluated with exact int64 arithmetic: the oracle on the CPU and the engine on the
GPU see the same bits.  Distributions follow PyTorch's defaults (U(-1/sqrt(fan_in), 1/sqrt(fan_in))
for Linear / Conv weights and biases); norm affines are perturbed (w = 1 + 0.1 z, b = 0.1 z) because
the default w = 1, b = 0 would hide affine bugs.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict

import torch

_MASK = 0xFFFFFFFF


def hash_uniform(numel: int, seed: int, device="cpu", chunk: int = 1 << 24) -> torch.Tensor:
    """fp32 tensor of `numel` values in [-1, 1), a pure function of (seed, index)."""
    out = torch.empty(numel, dtype=torch.float32, device=device)
    salt = (seed * 0x9E3779B1 + 0x7F4A7C15) & _MASK
    for start in range(0, numel, chunk):
        stop = min(start + chunk, numel)
        x = torch.arange(start, stop, dtype=torch.int64, device=device)
        x = (x + salt) & _MASK
        x = ((x ^ (x >> 16)) * 0x45D9F3B) & _MASK
        x = ((x ^ (x >> 16)) * 0x45D9F3B) & _MASK
        x = x ^ (x >> 16)
        out[start:stop] = (x.to(torch.float64) * (2.0 / 4294967296.0) - 1.0).to(torch.float32)
    return out


def _name_seed(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 2654435761)) & _MASK


def synth_tensor(name: str, shape, seed: int, device="cpu") -> torch.Tensor:
    """The synthetic fp32 value of parameter `name` (Diffusers key) with the given shape."""
    shape = tuple(shape)
    numel = 1
    for s in shape:
        numel *= s
    u = hash_uniform(numel, _name_seed(name, seed), device).reshape(shape)
    leaf = name.rsplit(".", 1)[-1]
    owner = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
    is_norm = "norm" in owner
    if is_norm:
        z = u * math.sqrt(3.0)  # unit variance
        return (1.0 + 0.1 * z) if leaf == "weight" else 0.1 * z
    return u  # scaled by the caller, who knows fan_in


def fill_module_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Overwrite every parameter of `module` in place with its synthetic value (in the parameter's own
    dtype and device; values are generated in fp32 and rounded once)."""
    fan_in: Dict[str, int] = {}
    for mod_name, mod in module.named_modules():
        if isinstance(mod, torch.nn.Linear):
            fan_in[mod_name] = mod.in_features
        elif isinstance(mod, torch.nn.Conv2d):
            fan_in[mod_name] = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1] // mod.groups
    with torch.no_grad():
        for name, p in module.named_parameters():
            owner = name.rsplit(".", 1)[0]
            val = synth_tensor(name, p.shape, seed, device=p.device)
            if owner in fan_in:
                val = val * (1.0 / math.sqrt(fan_in[owner]))
            p.copy_(val.to(p.dtype))
    return module


def synth_state_dict(module_on_meta: torch.nn.Module, seed: int = 0, device="cpu",
                     dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """State dict with synthetic values for a module built on the meta device (no default init cost)."""
    fan_in: Dict[str, int] = {}
    for mod_name, mod in module_on_meta.named_modules():
        if isinstance(mod, torch.nn.Linear):
            fan_in[mod_name] = mod.in_features
        elif isinstance(mod, torch.nn.Conv2d):
            fan_in[mod_name] = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1] // mod.groups
    sd = {}
    for name, p in module_on_meta.state_dict().items():
        owner = name.rsplit(".", 1)[0]
        val = synth_tensor(name, p.shape, seed, device=device)
        if owner in fan_in:
            val = val * (1.0 / math.sqrt(fan_in[owner]))
        sd[name] = val.to(dtype)
    return sd


def build_unet(cfg=None, seed: int = 0, device="cuda", dtype=torch.bfloat16, adm_input: bool = False):
    """A `UNet2DConditionModel` with synthetic weights, built on the meta device (no default-init cost)
    and materialised directly on `device`; values are generated in fp32 and rounded once to `dtype`."""
    from .unet import UNet2DConditionModel, UNetConfig

    cfg = cfg or UNetConfig.sdxl()
    with torch.device("meta"):
        model = UNet2DConditionModel(cfg, adm_input=adm_input)
    sd = synth_state_dict(model, seed=seed, device=device, dtype=dtype)
    model.load_state_dict(sd, strict=True, assign=True)
    for p in model.parameters():
        p.requires_grad_(False)
    return model.eval()


def synth_inputs(batch: int, latent: int, cfg, seed: int = 1234, device="cpu", dtype=torch.float32,
                 timestep: float = 999.0):
    """UNet inputs of SURVEY section 8d: sample ~ U-hash scaled to unit variance, ctx (B, 77, ctx_dim),
    text_embeds (B, text_dim), time_ids = [H, W, 0, 0, H, W] in pixels, t = 999."""
    s3 = math.sqrt(3.0)
    sample = synth_tensor("input.sample", (batch, cfg.in_channels, latent, latent), seed, device) * s3
    ctx = synth_tensor("input.encoder_hidden_states", (batch, 77, cfg.cross_attention_dim), seed, device) * s3
    text = synth_tensor("input.text_embeds", (batch, cfg.text_embed_dim), seed, device) * s3
    px = float(latent * 8)
    ids = [px, px, 0.0, 0.0, px, px][: cfg.num_time_ids]
    time_ids = torch.tensor([ids] * batch, dtype=torch.float32, device=device)
    t = torch.tensor(timestep, dtype=torch.float32, device=device)
    return {
        "sample": sample.to(dtype),
        "timesteps": t,
        "encoder_hidden_states": ctx.to(dtype),
        "added_cond_kwargs": {"text_embeds": text.to(dtype), "time_ids": time_ids.to(dtype)},
    }
