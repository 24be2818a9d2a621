"""Public entry point: `model = compile(model)`.

Same contract as the reference's `optimize_model(model, cuda_graph) -> GraphModule`
(`src/stabletriton/optimization.py:27-38`, advertised as `compile(model)` in README.md:5):
trace the nn.Module with torch.fx, run an ordered list of rewrite passes that swap eager sub-graphs for
fused ops (`replace_backend`, optimization.py:10-22), optionally wrap `forward` in a CUDA-graph replayer.
The returned module keeps the forward signature
`forward(sample, timesteps, encoder_hidden_states, added_cond_kwargs, **kwargs) -> [eps]`.

Differences, all deliberate:
  * the device must be a Blackwell B200 (compute capability 10.x) -- the kernels are sm_100a only;
    RuntimeError otherwise (the reference required >= 8.0, optimization.py:30-32);
  * parameters must already be bfloat16 (the reference assumed `.half()`): nothing is cast or mutated
    behind the caller's back; Conv2d weights are re-laid-out once to KRSC (channels-last), values intact;
  * every Linear and Conv2d is swapped too (the reference disabled its Linear pass and never had a
    working conv), with bias / SiLU / GEGLU / time-embedding / residual epilogues fused;
  * `.config` is carried over, so `pipe.unet = compile(unet)` needs no manual patch
    (cf. implementations/Diffusers/load_sdxl_pipeline.py:29-34);
  * weights stay live, as in the reference (wrappers read them through the original submodules): the fused projection
    buffers (QKV, cross-attention K/V, time-embedding) ARE the storage of the parameters they were built from, so
    in-place updates (`load_state_dict`, `weight.copy_`, a LoRA merge) reach every kernel and every captured graph;
    after an update that replaces parameter tensors, or that touches conv_in / conv_out, call
    `model.refresh_fused_weights()`;
  * besides forward(), the returned module offers `prepare(encoder_hidden_states, added_cond_kwargs)` and
    `step_forward(sample, timesteps, *prompt_constants)`: the prompt-constant part of the graph (cross-attention K/V
    projections, text / time-ids embedding) computed once per prompt instead of once per step (pipeline.DenoiseLoop
    uses them).
"""
from __future__ import annotations

import sys
from typing import Dict

import torch
import torch.fx as fx

from . import _cabi, fx_passes as P
from .cuda_graphs import make_dynamic_graphed_callable


def replace_backend(gm: fx.GraphModule, report: Dict[str, int] | None = None) -> fx.GraphModule:
    """Run every rewrite pass, in order.  `report`, if given, receives the match count of each pass."""
    passes = [
        ("remove_dropout", P.remove_dropout),
        ("fuse_attention", P.fuse_attention),
        ("fuse_qkv_projection", P.fuse_qkv_projection),
        ("fuse_shared_input_projections", P.fuse_shared_input_projections),
        ("fuse_time_embedding_projections", P.fuse_time_embedding_projections),
        ("fuse_linear_geglu", P.fuse_linear_geglu),
        ("fuse_geglu", P.fuse_geglu),
        ("fuse_proj_out_residual", P.fuse_proj_out_residual),
        ("fuse_linear_residual", P.fuse_linear_residual),
        ("replace_linear_activ", P.replace_linear_activ),
        ("replace_group_norm_activation", P.replace_group_norm_activation),
        ("replace_group_norm", P.replace_group_norm),
        ("replace_layer_norm", P.replace_layer_norm),
        ("replace_linear", P.replace_linear),
        ("fuse_conv_epilogues", P.fuse_conv_epilogues),
        ("replace_conv", P.replace_conv),
        ("replace_cat", P.replace_cat),
        ("replace_timesteps", P.replace_timesteps),
        ("keep_channels_last", P.keep_channels_last),
        ("fuse_group_norm_statistics", P.fuse_group_norm_statistics),
    ]
    for name, fn in passes:
        count = fn(gm)
        if report is not None:
            report[name] = count
    return gm


def run_compiler(gm: fx.GraphModule) -> fx.GraphModule:
    return replace_backend(gm)


def _check_device(model: torch.nn.Module) -> None:
    assert torch.cuda.is_available(), "CUDA capacity is required to use stabletriton_b200"
    first = next(model.parameters())
    assert first.device.type == "cuda", "Model must be on GPU"
    major, minor = torch.cuda.get_device_capability(first.device)
    if major != 10:
        raise RuntimeError(
            f"stabletriton_b200 kernels are built for sm_100a (B200, compute capability 10.x); "
            f"found {major}.{minor}")
    bad = {p.dtype for p in model.parameters()} - {torch.bfloat16}
    if bad:
        raise TypeError(f"Model parameters must be bfloat16 (call model.to(torch.bfloat16)); found {sorted(map(str, bad))}")
    _cabi.lib()  # fail loudly now if the CUDA library is missing


def pack_weights_(model: torch.nn.Module) -> None:
    """One-time weight layout change: Conv2d weights to dense KRSC storage (the B operand of the
    implicit GEMM reads (K, 3*3*C) rows).  Logical shape and values are unchanged."""
    from .kernels import pack_conv_weight

    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.Conv2d):
                m.weight.data = pack_conv_weight(m.weight.data)


def optimize_model(model: torch.nn.Module, cuda_graph: bool = True, *, check_device: bool = True) -> fx.GraphModule:
    """Trace, rewrite, optionally graph.  `check_device=False` is for CPU-side tests of the graph surgery."""
    if check_device:
        _check_device(model)
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))  # deep fx graphs (load_sdxl_pipeline.py:12)
    config = getattr(model, "config", None)
    cfg = getattr(model, "cfg", None)
    gm = P.trace(model)
    report: Dict[str, int] = {}
    replace_backend(gm, report)
    if check_device:
        pack_weights_(gm)
    # fused projection buffers become the storage of the parameters they were built from: weights stay live
    P.alias_fused_parameters_(gm)

    def refresh_fused_weights() -> int:
        """Re-sync every derived weight copy with the current parameters, in place (captured graphs keep working):
        fused projection buffers whose source parameter was REPLACED (load_state_dict(assign=True), `.to()`), and the
        zero-padded conv_in / conv_out operands.  In-place updates of Linear weights need no call: the parameters are
        views of the fused buffers.  Returns the number of tensors rewritten."""
        from . import kernels as K
        return P.refresh_fused_parameters_(gm) + K.refresh_padded_()

    gm.refresh_fused_weights = refresh_fused_weights
    gm.pass_report = report
    if config is not None:
        gm.config = config
    if cfg is not None:
        gm.cfg = cfg
    gm.eval()
    # Prompt-constant hoisting: `prepare` computes what depends on the prompt only (all cross-attention
    # K/V projections, the text / time-ids embedding), `step_forward` is the rest.  forward() itself is unchanged.
    prologue, body, n_consts = P.split_prompt_constants(gm)
    prologue.eval()
    body.eval()

    def prepare(encoder_hidden_states, added_cond_kwargs):
        with torch.no_grad():
            return tuple(prologue(encoder_hidden_states, added_cond_kwargs))

    def step_forward(sample, timesteps, *prompt_constants):
        with torch.no_grad():
            return body(sample, timesteps, *prompt_constants)

    gm.prepare, gm.step_forward, gm.num_prompt_constants = prepare, step_forward, n_consts
    if cuda_graph:
        eager_forward = gm.forward

        def _no_grad_forward(*args, **kwargs):
            with torch.no_grad():
                return eager_forward(*args, **kwargs)

        gm.eager_forward = _no_grad_forward
        gm.forward = make_dynamic_graphed_callable(_no_grad_forward)
    return gm


def compile(model: torch.nn.Module, cuda_graph: bool = True) -> fx.GraphModule:  # noqa: A001 (reference's name)
    return optimize_model(model, cuda_graph)
