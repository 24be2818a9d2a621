"""ComfyUI-side adapter, CPU: LDM <-> Diffusers UNet key renaming, the `adm` input variant of the
UNet, the rewritten graph of that variant, and the call convention ComfyUI's samplers use on `model.diffusion_model`."""
import pytest
import torch

import fake_kernels
from conftest import parity
from stabletriton_b200 import UNet2DConditionModel, UNetConfig, optimize_model, synth
from stabletriton_b200.comfy import (LDM_PREFIX, ComfyUNetAdapter, convert_ldm_unet_state_dict,
                                     diffusers_to_ldm_unet_state_dict, ldm_to_diffusers_module_map)


def _sdxl_shapes():
    with torch.device("meta"):
        return {k: v.shape for k, v in UNet2DConditionModel(UNetConfig.sdxl()).state_dict().items()}


def test_sdxl_key_renaming_is_a_bijection_onto_the_ldm_names():
    shapes = _sdxl_shapes()
    sd = {k: torch.empty(0) for k in shapes}
    ldm = diffusers_to_ldm_unet_state_dict(sd, prefix=LDM_PREFIX)
    assert len(ldm) == len(sd) == 1680
    names = {k[len(LDM_PREFIX):] for k in ldm}
    # spot checks against the module names of the original LDM / sgm SDXL UNet (9 input blocks, 9 output blocks)
    for k in ("time_embed.0.weight", "time_embed.2.bias", "label_emb.0.0.weight", "label_emb.0.2.bias",
              "input_blocks.0.0.weight", "input_blocks.1.0.in_layers.0.weight", "input_blocks.1.0.emb_layers.1.bias",
              "input_blocks.2.0.out_layers.3.weight", "input_blocks.3.0.op.weight", "input_blocks.4.0.skip_connection.weight",
              "input_blocks.4.1.proj_in.weight", "input_blocks.4.1.transformer_blocks.1.attn1.to_q.weight",
              "input_blocks.6.0.op.bias", "input_blocks.8.1.transformer_blocks.9.ff.net.0.proj.weight",
              "middle_block.0.in_layers.2.weight", "middle_block.1.transformer_blocks.9.attn2.to_k.weight",
              "middle_block.2.out_layers.0.bias", "output_blocks.0.0.skip_connection.weight",
              "output_blocks.2.1.transformer_blocks.0.norm3.weight", "output_blocks.2.2.conv.weight",
              "output_blocks.5.2.conv.bias", "output_blocks.5.1.proj_out.weight", "output_blocks.8.0.emb_layers.1.weight",
              "out.0.weight", "out.2.bias"):
        assert k in names, k
    assert not any(k.startswith("input_blocks.9.") or k.startswith("output_blocks.9.") for k in names)
    assert not any(k.startswith("output_blocks.8.1") or k.startswith("input_blocks.1.1") for k in names)  # no attention at 320
    back = convert_ldm_unet_state_dict({**ldm, "first_stage_model.decoder.conv_in.weight": torch.empty(0)})
    assert list(back) == list(sd)
    # the same without the checkpoint prefix (a bare diffusion_model state dict)
    assert list(convert_ldm_unet_state_dict(diffusers_to_ldm_unet_state_dict(sd))) == list(sd)
    with pytest.raises(KeyError):
        convert_ldm_unet_state_dict({"input_blocks.12.0.in_layers.0.weight": torch.empty(0)})
    table = ldm_to_diffusers_module_map()
    assert table["output_blocks.2.2.conv"] == "up_blocks.0.upsamplers.0.conv" and len(set(table.values())) == len(table)


def test_adm_variant_and_comfy_call_convention():
    cfg = UNetConfig.tiny()
    model = synth.build_unet(cfg, seed=3, device="cpu", dtype=torch.float32)
    adm_model = synth.build_unet(cfg, seed=3, device="cpu", dtype=torch.float32, adm_input=True)
    inp = synth.synth_inputs(2, 16, cfg, seed=5)
    added = inp["added_cond_kwargs"]
    # what ComfyUI's SDXL.encode_adm hands over: [pooled text | Fourier features of the six ids], cos half first
    y = torch.cat([added["text_embeds"], model.add_time_proj(added["time_ids"].flatten()).reshape(2, -1)], dim=-1)
    assert y.shape == (2, cfg.add_embed_in_dim)
    t = inp["timesteps"].expand(2).clone()
    with torch.no_grad():
        ref = model(**inp)[0]
        eager = ComfyUNetAdapter(adm_model, dtype=torch.float32)(inp["sample"], timesteps=t, context=inp["encoder_hidden_states"],
                                                                   y=y, control=None, transformer_options={})
        gm = optimize_model(adm_model, cuda_graph=False, check_device=False)
        with fake_kernels.installed():
            rewritten = ComfyUNetAdapter(gm, dtype=torch.float32)(inp["sample"], t, context=inp["encoder_hidden_states"], y=y)
    assert torch.is_tensor(eager) and eager.shape == ref.shape
    assert parity(eager, ref)[0] < 1e-6
    assert parity(rewritten, ref)[0] < 1e-5
    assert gm.pass_report["replace_timesteps"] == 1  # only the timestep embedding is left in the graph
    adapter = ComfyUNetAdapter(adm_model, dtype=torch.float32)
    with pytest.raises(NotImplementedError):
        adapter(inp["sample"], t, context=inp["encoder_hidden_states"], y=y, control={"output": []})
    with pytest.raises(ValueError):
        adapter(inp["sample"], t, context=inp["encoder_hidden_states"])
