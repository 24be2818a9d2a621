"""VAE decode path: model definition vs the fp32 oracle on the CPU; on a B200 the new kernels
(fp32-output GEMM, row softmax, tiny 1x1 conv, token transpose) and the whole decoder on the sm_100a kernels vs the
oracle -- tiny topology against the CPU oracle, the full SDXL VAE decoder at 512^2 and 1024^2 against the oracle in
fp32 on the GPU (TF32 off).  Bar as for the UNet: max|d| / max|ref| <= 2e-2, cosine >= 0.9999."""
import importlib.util
import math
import os

import pytest
import torch

from conftest import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle():
    spec = importlib.util.spec_from_file_location("vae_oracle", os.path.join(ROOT, "oracle", "vae_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _latents(batch, hw, seed, device="cpu", dtype=torch.float32):
    from stabletriton_b200 import synth
    z = synth.synth_tensor("input.latents", (batch, 4, hw, hw), seed, device) * math.sqrt(3.0)  # unit variance
    return (z * 0.13025 * 4.0).to(dtype)  # the scale of SDXL latents after the denoise loop (|z / 0.13025| of a few units)


# ------------------------------------------------------------------------------------------------- CPU
def test_sdxl_vae_decoder_has_the_diffusers_keys_and_size():
    from stabletriton_b200.vae import AutoencoderKLDecoder, VAEConfig
    with torch.device("meta"):
        m = AutoencoderKLDecoder(VAEConfig.sdxl())
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 49_490_179 + 20  # Diffusers' SDXL decoder + post_quant_conv
    assert len(sd) == 140
    assert tuple(sd["decoder.conv_in.weight"].shape) == (512, 4, 3, 3)
    assert tuple(sd["decoder.mid_block.attentions.0.to_q.weight"].shape) == (512, 512)
    assert tuple(sd["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"].shape) == (256, 512, 1, 1)
    assert tuple(sd["decoder.up_blocks.3.resnets.0.conv_shortcut.weight"].shape) == (128, 256, 1, 1)
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in sd and "decoder.up_blocks.2.upsamplers.0.conv.weight" in sd
    assert tuple(sd["decoder.conv_out.weight"].shape) == (3, 128, 3, 3)


@pytest.mark.parametrize("cfg_name,hw", [("tiny", 16), ("mid", 8)])
def test_model_definition_matches_oracle_fp32(cfg_name, hw):
    from stabletriton_b200.vae import VAEConfig, build_vae_decoder
    cfg = VAEConfig.tiny() if cfg_name == "tiny" else VAEConfig(block_out_channels=(64, 64, 128, 128), layers_per_block=2)
    model = build_vae_decoder(cfg, seed=5, device="cpu", dtype=torch.float32)
    z = _latents(2, hw, seed=9)
    O = _oracle()
    with torch.no_grad():
        got = model(z)
        ref = O.vae_decode(model.state_dict(), z, groups=cfg.norm_num_groups, eps=cfg.norm_eps,
                           scaling_factor=cfg.scaling_factor)
        chunked = O.vae_decode(model.state_dict(), z, groups=cfg.norm_num_groups, eps=cfg.norm_eps,
                               scaling_factor=cfg.scaling_factor, query_chunk=24)
    assert got.shape == (2, 3, hw * 2 ** (len(cfg.block_out_channels) - 1), hw * 2 ** (len(cfg.block_out_channels) - 1))
    rel, cos = parity(got, ref)
    assert rel <= 1e-5 and cos >= 0.999999, (rel, cos)
    rel, cos = parity(chunked, ref)
    assert rel <= 1e-5, (rel, cos)


def test_kernel_path_composition_with_the_eager_test_double():
    """The launch sequence of CompiledVAEDecoder (which producer feeds which GroupNorm its statistics, the q / k / v
    slices of the fused projection, score blocks of ATTN_ROWS query rows written into row slices of the output) run on
    the CPU with tests/fake_kernels.py standing in for the CUDA kernels: must reproduce the model definition."""
    import fake_kernels
    from stabletriton_b200.vae import CompiledVAEDecoder, VAEConfig, build_vae_decoder
    cfg = VAEConfig.tiny()
    model = build_vae_decoder(cfg, seed=5, device="cpu", dtype=torch.float32)
    z = _latents(2, 16, seed=9)
    vae = CompiledVAEDecoder.__new__(CompiledVAEDecoder)  # no device / dtype gate: this is the CPU double
    vae.model, vae.cfg, vae.cuda_graph, vae._graphs = model, cfg, False, {}
    att = model.decoder.mid_block.attentions[0]
    vae._wqkv = torch.cat([att.to_q.weight, att.to_k.weight, att.to_v.weight], dim=0)
    vae._bqkv = torch.cat([att.to_q.bias, att.to_k.bias, att.to_v.bias], dim=0)
    vae.ATTN_ROWS = 96  # 256 tokens -> three score blocks, the last one ragged
    with torch.no_grad(), fake_kernels.installed() as calls:
        got = vae._forward(z)
        pre = vae._forward(z / cfg.scaling_factor, pre_scaled=True)
        ref = model(z)
    assert parity(got, ref)[0] <= 1e-5 and parity(pre, ref)[0] <= 1e-5
    assert calls["softmax_rows"] == 2 * 2 * 3 and calls["matmul_nt_f32"] == 12 and calls["pointwise_conv_small"] == 2
    assert calls["groupnorm_from_partials"] >= 10  # the fake GroupNorm verifies every set of partials it is handed


def test_compile_vae_refuses_cpu_models():
    from stabletriton_b200.vae import VAEConfig, build_vae_decoder, compile_vae
    model = build_vae_decoder(VAEConfig.tiny(), device="cpu", dtype=torch.bfloat16)
    with pytest.raises(AssertionError):
        compile_vae(model)


# ------------------------------------------------------------------------------------------------- GPU
class _Fp32Exact:
    def __enter__(self):
        self.saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32,
                      torch.get_float32_matmul_precision())
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch.set_float32_matmul_precision("highest")

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.saved[:2]
        torch.set_float32_matmul_precision(self.saved[2])


def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k", [(256, 256, 128), (300, 1000, 512), (4096, 16384, 512)])
def test_gemm_fp32_output(built_lib, m, n, k):
    from stabletriton_b200 import kernels as K
    a, b = _rnd(m, k, seed=1).cuda(), _rnd(n, k, seed=2).cuda()
    n8 = n - n % 8
    got = K.matmul_nt_f32(a, b[:n8])
    assert got.dtype == torch.float32 and got.shape == (m, n8)
    with _Fp32Exact():
        ref = a.float() @ b[:n8].float().t()
    rel, cos = parity(got, ref)
    assert rel <= 1e-5 and cos >= 0.9999999, (rel, cos)  # fp32 accumulation of exact bf16 products on both sides


@pytest.mark.gpu
@pytest.mark.parametrize("m,n", [(7, 256), (33, 4096), (16, 4100), (5, 16384), (3, 20000)])
def test_softmax_rows(built_lib, m, n):
    from stabletriton_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    s = (torch.randn(m, n, generator=g) * 30.0).cuda()
    s[0, 5] = 400.0  # one dominant score: the others must underflow to 0, not NaN
    got = K.softmax_rows(s, 0.044)
    ref = torch.softmax(s.double() * 0.044, dim=-1)
    assert got.dtype == torch.bfloat16 and bool(torch.isfinite(got.float()).all())
    assert (got.double() - ref).abs().max().item() <= 4e-3 * ref.max().item() + 1e-6  # bf16 rounding of the probabilities
    assert (got.double().sum(-1) - 1.0).abs().max().item() <= 5e-3


@pytest.mark.gpu
def test_small_pointwise_conv_transpose_and_linear_into_a_slice(built_lib):
    from stabletriton_b200 import kernels as K
    x, w, b = _rnd(2, 4, 16, 24, seed=4).cuda(), _rnd(4, 4, 1, 1, seed=5).cuda(), _rnd(4, seed=6).cuda()
    got = K.pointwise_conv_small(x, w, b, in_scale=1.0 / 0.13025)
    ref = torch.nn.functional.conv2d(x.float() / 0.13025, w.float(), b.float())
    rel, _ = parity(got, ref)
    assert rel <= 8e-3, rel
    t = _rnd(2, 100, 72, seed=7).cuda()
    assert torch.equal(K.transpose_tokens(t), t.transpose(1, 2).contiguous())
    wide = _rnd(2, 100, 216, seed=8).cuda()
    assert torch.equal(K.transpose_tokens(wide[..., 72:144]), wide[..., 72:144].transpose(1, 2).contiguous())
    a, wt = _rnd(300, 128, seed=9).cuda(), _rnd(64, 128, seed=10, scale=0.1).cuda()
    buf = torch.full((2, 300, 64), -7.0, dtype=torch.bfloat16, device="cuda")
    K.linear(a, wt, out=buf[1])
    assert torch.equal(buf[1], K.linear(a, wt)) and bool((buf[0] == -7.0).all())


@pytest.mark.gpu
def test_tiny_vae_decode_matches_oracle_and_graph_replay_is_exact(built_lib):
    from stabletriton_b200.vae import VAEConfig, build_vae_decoder, compile_vae
    cfg = VAEConfig.tiny()
    model = build_vae_decoder(cfg, seed=5, device="cuda", dtype=torch.bfloat16)
    sd32 = {k: v.float().cpu() for k, v in model.state_dict().items()}
    vae = compile_vae(model)
    z = _latents(2, 16, seed=9, dtype=torch.bfloat16)
    ref = _oracle().vae_decode(sd32, z.float(), groups=cfg.norm_num_groups, eps=cfg.norm_eps, scaling_factor=cfg.scaling_factor)
    eager = vae.eager_decode(z.cuda())
    got = vae.decode(z.cuda())
    again = vae.decode(z.cuda())
    assert got.shape == ref.shape and got.dtype == torch.bfloat16
    assert torch.equal(got, eager) and torch.equal(got, again)
    rel, cos = parity(got, ref)
    print(f"tiny VAE decode B=2 16x16 -> 32x32: rel={rel:.3e} cos={cos:.6f}")
    assert rel <= 2e-2 and cos >= 0.9999, (rel, cos)
    with pytest.raises(ValueError):
        vae.decode(z.float().cuda())


@pytest.fixture(scope="module")
def sdxl_vae(built_lib):
    from stabletriton_b200.vae import VAEConfig, build_vae_decoder, compile_vae
    cfg = VAEConfig.sdxl()
    model = build_vae_decoder(cfg, seed=11, device="cuda", dtype=torch.bfloat16)
    sd32 = {k: v.float() for k, v in model.state_dict().items()}
    vae = compile_vae(model)
    yield vae, sd32, cfg
    del vae, sd32, model
    torch.cuda.empty_cache()


@pytest.mark.gpu
@pytest.mark.parametrize("batch,latent", [(2, 64), (1, 128)])
def test_sdxl_vae_decode_matches_fp32_oracle(sdxl_vae, batch, latent):
    """Full SDXL VAE decoder: 512^2 (B = 2) and 1024^2 (the image size of the benchmarked UNet configuration: 16 384
    tokens through the single 512-wide attention head)."""
    vae, sd32, cfg = sdxl_vae
    z = _latents(batch, latent, seed=21, device="cuda", dtype=torch.bfloat16)
    with _Fp32Exact(), torch.no_grad():
        ref = _oracle().vae_decode(sd32, z.float(), groups=cfg.norm_num_groups, eps=cfg.norm_eps,
                                   scaling_factor=cfg.scaling_factor, query_chunk=2048)
    got = vae.decode(z)
    torch.cuda.synchronize()
    assert got.shape == (batch, 3, 8 * latent, 8 * latent)
    rel, cos = parity(got, ref)
    print(f"SDXL VAE decode B={batch} latent {latent} -> {8 * latent}^2: rel={rel:.3e} cos={cos:.6f}")
    assert rel <= 2e-2 and cos >= 0.9999, (rel, cos)
    del ref, got
    torch.cuda.empty_cache()
