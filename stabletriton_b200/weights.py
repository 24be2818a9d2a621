"""Diffusers-compatible weight I/O for the SDXL UNet.

The reference loads a Diffusers checkpoint by building its own `UNet2DConditionModelPT` and calling a strict
`load_state_dict(pipe.unet.state_dict())` (implementations/Diffusers/load_sdxl_pipeline.py:24-26), then patches the
`.config` namedtuple back by hand (:29-34).  Here the same 1680 Diffusers key names are read straight from the
on-disk format -- `unet/diffusion_pytorch_model[.fp16].safetensors` (or its sharded index) + `unet/config.json` --
into bf16 parameters on the target device, tensor by tensor (no fp32 staging copy of the 10 GB checkpoint), and
`compile()` carries `.config` over.  Nothing here touches the network: paths are local files.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Optional, Tuple

import torch

from .unet import UNet2DConditionModel, UNetConfig

_WEIGHT_FILES = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors",
                 "diffusion_pytorch_model.bf16.safetensors")
_INDEX_FILES = tuple(f + ".index.json" for f in _WEIGHT_FILES)


def config_from_diffusers(cfg_json: dict) -> UNetConfig:
    """UNetConfig from a Diffusers `unet/config.json` dict; rejects architectures this engine does not implement."""
    def need(key, expect):
        got = cfg_json.get(key, expect)
        same = list(got) == list(expect) if isinstance(expect, (list, tuple)) else got == expect
        if not same:
            raise ValueError(f"unsupported UNet config: {key} = {got!r} (this engine implements {expect!r})")

    blocks = tuple(cfg_json.get("block_out_channels", (320, 640, 1280)))
    need("down_block_types", ["DownBlock2D"] + ["CrossAttnDownBlock2D"] * (len(blocks) - 1))
    need("up_block_types", ["CrossAttnUpBlock2D"] * (len(blocks) - 1) + ["UpBlock2D"])
    need("addition_embed_type", "text_time")
    need("use_linear_projection", True)
    tl = cfg_json.get("transformer_layers_per_block", (1, 2, 10))
    tl = (tl,) * len(blocks) if isinstance(tl, int) else tuple(tl)
    if len(tl) != len(blocks):
        raise ValueError("transformer_layers_per_block / block_out_channels length mismatch")
    head = cfg_json.get("attention_head_dim", (5, 10, 20))
    # Diffusers stores the number of heads per block here (SDXL: 5/10/20 -> 64 channels per head everywhere)
    heads = (head,) * len(blocks) if isinstance(head, int) else tuple(head)
    dims = {c // h for c, h in zip(blocks, heads)}
    if len(dims) != 1:
        raise ValueError(f"heads {heads} over channels {blocks} do not give one head dim")
    add_dim = cfg_json.get("addition_time_embed_dim", 256)
    proj_in = cfg_json.get("projection_class_embeddings_input_dim", 2816)
    num_ids = 6
    return UNetConfig(
        in_channels=cfg_json.get("in_channels", 4), out_channels=cfg_json.get("out_channels", 4),
        block_out_channels=blocks, layers_per_block=cfg_json.get("layers_per_block", 2),
        transformer_layers_per_block=(0,) + tl[1:],  # block 0 is a plain DownBlock2D: its entry is ignored
        attention_head_dim=dims.pop(), cross_attention_dim=cfg_json.get("cross_attention_dim", 2048),
        norm_num_groups=cfg_json.get("norm_num_groups", 32), addition_time_embed_dim=add_dim,
        text_embed_dim=proj_in - num_ids * add_dim, num_time_ids=num_ids, sample_size=cfg_json.get("sample_size", 128))


def _weight_files(unet_dir: str) -> Tuple[Iterable[str], Optional[Dict[str, str]]]:
    for idx in _INDEX_FILES:
        p = os.path.join(unet_dir, idx)
        if os.path.exists(p):
            with open(p) as f:
                wm = json.load(f)["weight_map"]
            return sorted({os.path.join(unet_dir, v) for v in wm.values()}), wm
    for w in _WEIGHT_FILES:
        p = os.path.join(unet_dir, w)
        if os.path.exists(p):
            return [p], None
    raise FileNotFoundError(f"no diffusion_pytorch_model*.safetensors under {unet_dir}")


def load_diffusers_unet(path: str, cfg: Optional[UNetConfig] = None, device="cuda", dtype=torch.bfloat16,
                        strict: bool = True) -> UNet2DConditionModel:
    """Build the UNet and fill it from a local Diffusers checkpoint.

    `path` is the `unet/` directory of a pipeline (or the pipeline directory, or one `.safetensors` file).  Key names
    must match the model's state dict exactly (1680 keys for SDXL-base), as the reference's strict load does."""
    from safetensors import safe_open

    if os.path.isdir(path) and os.path.isdir(os.path.join(path, "unet")):
        path = os.path.join(path, "unet")
    if os.path.isdir(path):
        files, _ = _weight_files(path)
        cj = os.path.join(path, "config.json")
        if cfg is None and os.path.exists(cj):
            with open(cj) as f:
                cfg = config_from_diffusers(json.load(f))
    else:
        files = [path]
    cfg = cfg or UNetConfig.sdxl()
    with torch.device("meta"):
        model = UNet2DConditionModel(cfg)
    model = model.to_empty(device=device).to(dtype)
    params = dict(model.state_dict())
    seen = set()
    for fn in files:
        with safe_open(fn, framework="pt", device="cpu") as f:
            for key in f.keys():
                if key not in params:
                    if strict:
                        raise KeyError(f"unexpected key in checkpoint: {key}")
                    continue
                t = f.get_tensor(key)
                if tuple(t.shape) != tuple(params[key].shape):
                    raise ValueError(f"{key}: checkpoint shape {tuple(t.shape)} != model shape {tuple(params[key].shape)}")
                params[key].copy_(t)  # casts fp16 / fp32 -> dtype on the way to the device
                seen.add(key)
    missing = sorted(set(params) - seen)
    if missing and strict:
        raise KeyError(f"{len(missing)} keys missing from checkpoint, first: {missing[:3]}")
    return model.eval().requires_grad_(False)


def save_diffusers_unet(model: torch.nn.Module, unet_dir: str, cfg: Optional[UNetConfig] = None) -> str:
    """Write `diffusion_pytorch_model.safetensors` + `config.json` in the Diffusers layout (state-dict keys as is)."""
    from safetensors.torch import save_file

    os.makedirs(unet_dir, exist_ok=True)
    sd = {k: v.detach().to("cpu").contiguous() for k, v in model.state_dict().items()}
    out = os.path.join(unet_dir, _WEIGHT_FILES[0])
    save_file(sd, out, metadata={"format": "pt"})
    cfg = cfg or getattr(model, "cfg", None)
    if cfg is not None:
        n = len(cfg.block_out_channels)
        heads = [c // cfg.attention_head_dim for c in cfg.block_out_channels]
        with open(os.path.join(unet_dir, "config.json"), "w") as f:
            json.dump({
                "_class_name": "UNet2DConditionModel", "in_channels": cfg.in_channels, "out_channels": cfg.out_channels,
                "block_out_channels": list(cfg.block_out_channels), "layers_per_block": cfg.layers_per_block,
                "transformer_layers_per_block": [max(1, t) for t in cfg.transformer_layers_per_block],
                "attention_head_dim": heads, "cross_attention_dim": cfg.cross_attention_dim,
                "norm_num_groups": cfg.norm_num_groups, "addition_embed_type": "text_time",
                "addition_time_embed_dim": cfg.addition_time_embed_dim, "use_linear_projection": True,
                "projection_class_embeddings_input_dim": cfg.add_embed_in_dim, "sample_size": cfg.sample_size,
                "down_block_types": ["DownBlock2D"] + ["CrossAttnDownBlock2D"] * (n - 1),
                "up_block_types": ["CrossAttnUpBlock2D"] * (n - 1) + ["UpBlock2D"],
            }, f, indent=1)
    return out
