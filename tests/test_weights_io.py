"""Diffusers-layout weight I/O: safetensors + config.json round trip with the Diffusers key names,
strict key / shape checking as the reference's `load_state_dict(pipe.unet.state_dict())`
(implementations/Diffusers/load_sdxl_pipeline.py:24-26)."""
import dataclasses
import json
import os

import pytest
import torch

from stabletriton_b200 import UNet2DConditionModel, UNetConfig, synth
from stabletriton_b200.weights import config_from_diffusers, load_diffusers_unet, save_diffusers_unet

SDXL_CONFIG_JSON = {  # the fields of stabilityai/stable-diffusion-xl-base-1.0 unet/config.json this engine reads
    "in_channels": 4, "out_channels": 4, "block_out_channels": [320, 640, 1280], "layers_per_block": 2,
    "transformer_layers_per_block": [1, 2, 10], "attention_head_dim": [5, 10, 20], "cross_attention_dim": 2048,
    "norm_num_groups": 32, "addition_embed_type": "text_time", "addition_time_embed_dim": 256,
    "projection_class_embeddings_input_dim": 2816, "use_linear_projection": True, "sample_size": 128,
    "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"],
    "up_block_types": ["CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"],
}


def test_sdxl_config_json_maps_to_the_sdxl_config():
    assert dataclasses.asdict(config_from_diffusers(SDXL_CONFIG_JSON)) == dataclasses.asdict(UNetConfig.sdxl())
    bad = dict(SDXL_CONFIG_JSON, use_linear_projection=False)
    with pytest.raises(ValueError):
        config_from_diffusers(bad)
    bad = dict(SDXL_CONFIG_JSON, down_block_types=["CrossAttnDownBlock2D"] * 3)
    with pytest.raises(ValueError):
        config_from_diffusers(bad)


def test_round_trip_tiny_unet(tmp_path):
    cfg = UNetConfig.tiny()
    model = synth.build_unet(cfg, seed=5, device="cpu", dtype=torch.float32)
    unet_dir = os.path.join(tmp_path, "pipe", "unet")
    save_diffusers_unet(model, unet_dir, cfg)
    with open(os.path.join(unet_dir, "config.json")) as f:
        assert dataclasses.asdict(config_from_diffusers(json.load(f))) == dataclasses.asdict(cfg)
    # the pipeline directory, the unet directory and the bare file are all accepted
    for path in (os.path.join(tmp_path, "pipe"), unet_dir, os.path.join(unet_dir, "diffusion_pytorch_model.safetensors")):
        loaded = load_diffusers_unet(path, cfg=None if os.path.isdir(path) else cfg, device="cpu", dtype=torch.float32)
        a, b = model.state_dict(), loaded.state_dict()
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)
    # bf16 load = the rounding of the stored values
    loaded = load_diffusers_unet(unet_dir, device="cpu", dtype=torch.bfloat16)
    k = "mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj.weight"
    assert torch.equal(loaded.state_dict()[k], model.state_dict()[k].to(torch.bfloat16))
    assert loaded.config.in_channels == cfg.in_channels and loaded.config.sample_size == cfg.sample_size


def test_strict_loading_rejects_wrong_checkpoints(tmp_path):
    from safetensors.torch import save_file

    cfg = UNetConfig.tiny()
    with torch.device("meta"):
        meta = UNet2DConditionModel(cfg)
    sd = {k: torch.zeros(v.shape) for k, v in meta.state_dict().items()}
    extra = dict(sd, **{"not.a.unet.key": torch.zeros(1)})
    p = os.path.join(tmp_path, "extra.safetensors")
    save_file(extra, p)
    with pytest.raises(KeyError):
        load_diffusers_unet(p, cfg=cfg, device="cpu", dtype=torch.float32)
    missing = dict(sd)
    missing.pop("conv_in.weight")
    p = os.path.join(tmp_path, "missing.safetensors")
    save_file(missing, p)
    with pytest.raises(KeyError):
        load_diffusers_unet(p, cfg=cfg, device="cpu", dtype=torch.float32)
    wrong = dict(sd, **{"conv_in.weight": torch.zeros(3, 3)})
    p = os.path.join(tmp_path, "shape.safetensors")
    save_file(wrong, p)
    with pytest.raises(ValueError):
        load_diffusers_unet(p, cfg=cfg, device="cpu", dtype=torch.float32)
