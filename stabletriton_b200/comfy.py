"""ComfyUI-side adapter.

The reference advertises `model = compile(model)` for ComfyUI next to Diffusers (README.md:5) and ships an empty file for
it (implementations/ComfyUI/example.py, 0 bytes), so there is no reference code to mirror -- only ComfyUI's own calling
convention.  ComfyUI is not installed in this image either; what is here is the part that does not need it:

  * `convert_ldm_unet_state_dict`: ComfyUI / sgm checkpoints store the SDXL UNet under `model.diffusion_model.*` with the
    original LDM module names (`input_blocks.4.1.transformer_blocks.0.attn1.to_q.weight`, `label_emb.0.0.weight`, ...).
    The engine keeps the reference's Diffusers names (strict `load_state_dict`, load_sdxl_pipeline.py:24-26); this is the
    renaming, derived from the UNet configuration (no per-key table), invertible (`diffusers_to_ldm_unet_state_dict`).
  * `ComfyUNetAdapter`: the call ComfyUI's samplers make on `model.diffusion_model` --
    `forward(x, timesteps, context=..., y=..., control=None, transformer_options={})` -- mapped onto the compiled UNet.
    ComfyUI hands the SDXL micro-conditioning over already embedded (`y` = [pooled text | Fourier features of the six
    size / crop ids], 2816 wide, model_base.SDXL.encode_adm); a UNet built with `adm_input=True` takes that vector as
    `added_cond_kwargs["adm"]` instead of embedding `time_ids` itself.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .unet import UNetConfig

LDM_PREFIX = "model.diffusion_model."

_RESNET = {"in_layers.0": "norm1", "in_layers.2": "conv1", "emb_layers.1": "time_emb_proj", "out_layers.0": "norm2",
           "out_layers.3": "conv2", "skip_connection": "conv_shortcut"}


def ldm_to_diffusers_module_map(cfg: Optional[UNetConfig] = None) -> Dict[str, str]:
    """{LDM module prefix: Diffusers module prefix} for every container whose name differs.  Transformer internals
    (`norm`, `proj_in`, `transformer_blocks.N.attn1.to_q`, `ff.net.0.proj`, ...) carry the same names on both sides."""
    cfg = cfg or UNetConfig.sdxl()
    levels = len(cfg.block_out_channels)
    n_res = cfg.layers_per_block
    m: Dict[str, str] = {
        "time_embed.0": "time_embedding.linear_1", "time_embed.2": "time_embedding.linear_2",
        "label_emb.0.0": "add_embedding.linear_1", "label_emb.0.2": "add_embedding.linear_2",
        "input_blocks.0.0": "conv_in", "out.0": "conv_norm_out", "out.2": "conv_out",
        "middle_block.0": "mid_block.resnets.0", "middle_block.1": "mid_block.attentions.0",
        "middle_block.2": "mid_block.resnets.1",
    }
    idx = 1
    for lvl in range(levels):  # encoder: [res (+ transformer)] x layers_per_block, then a stride-2 conv between levels
        has_attn = cfg.transformer_layers_per_block[lvl] > 0
        for i in range(n_res):
            m[f"input_blocks.{idx}.0"] = f"down_blocks.{lvl}.resnets.{i}"
            if has_attn:
                m[f"input_blocks.{idx}.1"] = f"down_blocks.{lvl}.attentions.{i}"
            idx += 1
        if lvl < levels - 1:
            m[f"input_blocks.{idx}.0.op"] = f"down_blocks.{lvl}.downsamplers.0.conv"
            idx += 1
    idx = 0
    for j, lvl in enumerate(reversed(range(levels))):  # decoder: layers_per_block + 1 blocks per level, upsample on the last
        has_attn = cfg.transformer_layers_per_block[lvl] > 0
        for i in range(n_res + 1):
            m[f"output_blocks.{idx}.0"] = f"up_blocks.{j}.resnets.{i}"
            if has_attn:
                m[f"output_blocks.{idx}.1"] = f"up_blocks.{j}.attentions.{i}"
            if i == n_res and lvl > 0:
                m[f"output_blocks.{idx}.{2 if has_attn else 1}.conv"] = f"up_blocks.{j}.upsamplers.0.conv"
            idx += 1
    return m


def _rename(key: str, table: Dict[str, str], resnet: Dict[str, str], resnet_side: str) -> str:
    """Longest-prefix module renaming, then the resnet-internal names when the module is a resnet."""
    best = ""
    for src in table:
        if (key == src or key.startswith(src + ".")) and len(src) > len(best):
            best = src
    if not best:
        raise KeyError(f"no UNet module matches checkpoint key '{key}'")
    dst, rest = table[best], key[len(best):]
    is_resnet = ".resnets." in (dst if resnet_side == "dst" else best)
    if is_resnet:
        for a, b in resnet.items():
            if rest.startswith("." + a + "."):
                rest = "." + b + rest[len(a) + 1:]
                break
    return dst + rest


def convert_ldm_unet_state_dict(sd: Dict[str, torch.Tensor], cfg: Optional[UNetConfig] = None) -> Dict[str, torch.Tensor]:
    """ComfyUI / sgm UNet weights (`model.diffusion_model.` prefix optional; other sub-models ignored) -> the Diffusers
    key names the engine loads strictly.  Tensors are passed through untouched."""
    table = ldm_to_diffusers_module_map(cfg)
    out = {}
    has_prefix = any(k.startswith(LDM_PREFIX) for k in sd)
    for k, v in sd.items():
        if has_prefix:
            if not k.startswith(LDM_PREFIX):
                continue  # first_stage_model.*, conditioner.*, ...
            k = k[len(LDM_PREFIX):]
        out[_rename(k, table, _RESNET, "dst")] = v
    return out


def diffusers_to_ldm_unet_state_dict(sd: Dict[str, torch.Tensor], cfg: Optional[UNetConfig] = None,
                                     prefix: str = "") -> Dict[str, torch.Tensor]:
    """The inverse renaming (writing a ComfyUI-style checkpoint; round-trip tests)."""
    table = {v: k for k, v in ldm_to_diffusers_module_map(cfg).items()}
    inv_res = {v: k for k, v in _RESNET.items()}
    return {prefix + _rename(k, table, inv_res, "src"): v for k, v in sd.items()}


class ComfyUNetAdapter(torch.nn.Module):
    """`model.diffusion_model` as ComfyUI calls it, backed by a compiled engine UNet built with `adm_input=True`.

    forward(x, timesteps, context, y) -> eps tensor (not a list).  ControlNet residuals (`control`) and attention patches
    (`transformer_options["patches"]`) would have to be part of the captured graph and are refused rather than ignored."""

    def __init__(self, compiled_unet, dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        self.unet = compiled_unet
        self.dtype = dtype

    def forward(self, x, timesteps=None, context=None, y=None, control=None, transformer_options=None, **kwargs):
        if control is not None:
            raise NotImplementedError("ComfyUNetAdapter: ControlNet residuals are not part of the compiled graph")
        if transformer_options and transformer_options.get("patches"):
            raise NotImplementedError("ComfyUNetAdapter: attention patches are not part of the compiled graph")
        if timesteps is None or context is None or y is None:
            raise ValueError("ComfyUNetAdapter: timesteps, context and y (the SDXL adm vector) are required")
        out_dtype = x.dtype
        eps = self.unet(x.to(self.dtype), timesteps.to(torch.float32), context.to(self.dtype),
                        {"adm": y.to(self.dtype)})[0]
        return eps.to(out_dtype)
