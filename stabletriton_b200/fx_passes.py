"""torch.fx graph surgery: swap eager sub-graphs of the traced UNet for calls into `wrappers`.

Role of the reference's L3 layer (`optimizers/remove_dropout.py`, `replace_*.py`) and of its forked
SubgraphMatcher (`optimizers/utils/util.py:56-524`, `utils/fx.py:21-38`).  The reference matches by
running a generic backtracking sub-graph isomorphism over (pattern, replacement) module pairs; here
each pass anchors on one distinctive node (softmax, gelu, a module type, an `add`) and walks its
producers/consumers explicitly -- a few lines per pattern, linear time, and it can match the
multi-node epilogue patterns (conv + time-embedding add, linear + residual, proj_out + image residual)
that make the fused kernels possible.  Like the reference (util.py:459-468), a replaced module is
handed to its wrapper through a `get_attr` node on the *original* submodule, so weights stay live.

Pass order matters (cf. optimization.py:10-22): dropout first (Dropout nodes split patterns), fused
variants (+GEGLU, +residual, +activation) before the plain ones.
"""
from __future__ import annotations

import inspect
import operator
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.fx as fx
import torch.nn as nn
import torch.nn.functional as F

from . import wrappers as W

Node = fx.Node


# ------------------------------------------------------------------------------------------------
# tracing
# ------------------------------------------------------------------------------------------------
class UNetTracer(fx.Tracer):
    """symbolic_trace, except that sinusoidal `Timesteps` modules (matched by class name, so the
    reference's own unet_pt.Timesteps qualifies) stay leaves and can be swapped as a unit."""

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        if type(m).__name__ == "Timesteps" and hasattr(m, "num_channels"):
            return True
        return super().is_leaf_module(m, module_qualified_name)


def trace(model: nn.Module) -> fx.GraphModule:
    tracer = UNetTracer()
    graph = tracer.trace(model)
    return fx.GraphModule(tracer.root, graph, type(model).__name__)


# ------------------------------------------------------------------------------------------------
# small matching vocabulary
# ------------------------------------------------------------------------------------------------
def _module_of(gm: fx.GraphModule, node: Node, cls) -> Optional[nn.Module]:
    if isinstance(node, Node) and node.op == "call_module":
        m = gm.get_submodule(node.target)
        if isinstance(m, cls):
            return m
    return None


def _is_method(node, *names) -> bool:
    return isinstance(node, Node) and node.op == "call_method" and node.target in names


def _is_function(node, *fns) -> bool:
    return isinstance(node, Node) and node.op == "call_function" and node.target in fns


def _only_user(node: Node) -> Optional[Node]:
    users = list(node.users)
    return users[0] if len(users) == 1 else None


def _is_silu(gm, node) -> bool:
    return _module_of(gm, node, nn.SiLU) is not None or _is_function(node, F.silu)


def _module_attr(gm: fx.GraphModule, before: Node, target: str) -> Node:
    """`get_attr` node resolving to the submodule `target`, inserted before `before`."""
    with gm.graph.inserting_before(before):
        return gm.graph.get_attr(target)


def _call(gm: fx.GraphModule, before: Node, fn: Callable, args: tuple, kwargs: Optional[dict] = None) -> Node:
    with gm.graph.inserting_before(before):
        return gm.graph.call_function(fn, args, kwargs or {})


def _finish(gm: fx.GraphModule) -> None:
    gm.graph.eliminate_dead_code()
    gm.graph.lint()
    gm.recompile()


_ADD = (operator.add, torch.add)
_MUL = (operator.mul, torch.mul)


# ------------------------------------------------------------------------------------------------
# fused parameter buffers stay LIVE
# ------------------------------------------------------------------------------------------------
# The fusion passes below row-concatenate the weights of several Linears into one buffer (QKV, all cross-attention K/V
# projections, all resnet time-embedding projections).  A plain copy would freeze the weights at compile() time: a
# later in-place update (load_state_dict, a LoRA merge, `weight.copy_`) would change the un-fused Linears but not the
# attention / time-embedding projections.  So every fused buffer records where its rows came from
# (`gm._st_fused_sources`), and `alias_fused_parameters_` makes the buffer the storage of those parameters: each
# module's weight / bias becomes a row-slice view of it, so in-place updates reach the kernels with no copy -- also
# through an already captured CUDA graph, which reads the same addresses.  Updates that REPLACE a parameter tensor
# (load_state_dict(assign=True), `module.weight = ...`) break the aliasing; `refresh_fused_parameters_` copies the
# current parameter values back into the buffers (in place) and re-establishes it.
def _record_fused(gm: fx.GraphModule, buffer: str, parts: List[Tuple[str, str, int, int]]) -> None:
    """parts: (module qualified name, 'weight' | 'bias', first row in the buffer, rows)."""
    if not hasattr(gm, "_st_fused_sources"):
        gm._st_fused_sources = {}
    gm._st_fused_sources[buffer] = list(parts)


def alias_fused_parameters_(gm: fx.GraphModule) -> int:
    """Re-point every parameter that was copied into a fused buffer at its rows of that buffer.  Returns the number
    of parameters aliased."""
    n = 0
    with torch.no_grad():
        for buffer, parts in getattr(gm, "_st_fused_sources", {}).items():
            buf = gm.get_buffer(buffer)
            for qualname, attr, off, rows in parts:
                param = getattr(gm.get_submodule(qualname), attr)
                view = buf[off:off + rows]
                if param.data_ptr() != view.data_ptr() or param.device != view.device:
                    if param.shape != view.shape or param.dtype != view.dtype:
                        raise ValueError(f"fused buffer {buffer}: {qualname}.{attr} changed shape or dtype")
                    param.data = view
                n += 1
    return n


def refresh_fused_parameters_(gm: fx.GraphModule) -> int:
    """Copy the CURRENT value of every fused parameter into its buffer (in place: captured graphs keep reading the same
    addresses) and re-alias it.  Needed only after a parameter tensor was replaced instead of updated in place."""
    n = 0
    with torch.no_grad():
        for buffer, parts in getattr(gm, "_st_fused_sources", {}).items():
            buf = gm.get_buffer(buffer)
            for qualname, attr, off, rows in parts:
                param = getattr(gm.get_submodule(qualname), attr)
                view = buf[off:off + rows]
                if param.data_ptr() != view.data_ptr():
                    view.copy_(param.detach().to(device=view.device, dtype=view.dtype))
                    n += 1
    alias_fused_parameters_(gm)
    return n


# ------------------------------------------------------------------------------------------------
# passes
# ------------------------------------------------------------------------------------------------
def remove_dropout(gm: fx.GraphModule) -> int:
    """Erase nn.Dropout / F.dropout nodes: p = 0 at inference (reference: remove_dropout.py:19-33)."""
    n = 0
    for node in list(gm.graph.nodes):
        if _module_of(gm, node, nn.Dropout) is not None or _is_function(node, F.dropout):
            node.replace_all_uses_with(node.args[0])
            gm.graph.erase_node(node)
            n += 1
    _finish(gm)
    return n


def _split_heads_source(node) -> Optional[Tuple[Node, int, int]]:
    """node == X.view(_, _, H, D).transpose(1, 2)  ->  (X, H, D)."""
    if not (_is_method(node, "transpose") and tuple(node.args[1:]) in ((1, 2), (2, 1))):
        return None
    view = node.args[0]
    if not (_is_method(view, "view", "reshape") and len(view.args) == 5):
        return None
    h, d = view.args[3], view.args[4]
    if not (isinstance(h, int) and isinstance(d, int)):
        return None
    return view.args[0], h, d


def fuse_attention(gm: fx.GraphModule) -> int:
    """view/transpose -> QK^T * scale -> softmax -> .V -> transpose/contiguous/view  ==>  attention_wrapper.
    Same pattern as the reference (replace_attention.py:76-86); the literals bound to `sm_scale`,
    `num_heads`, `head_dim` are read off the matched nodes."""
    n = 0
    for sm in list(gm.graph.nodes):
        if not (_is_function(sm, torch.softmax, F.softmax) or _is_method(sm, "softmax")):
            continue
        scaled = sm.args[0]
        if not (_is_function(scaled, *_MUL) and len(scaled.args) == 2):
            continue
        a, b = scaled.args
        scores, scale = (a, b) if isinstance(b, (int, float)) else (b, a)
        if not isinstance(scale, (int, float)) or not _is_function(scores, torch.matmul):
            continue
        q_t, k_tt = scores.args
        if not (_is_method(k_tt, "transpose") and tuple(k_tt.args[1:]) in ((-2, -1), (-1, -2), (2, 3), (3, 2))):
            continue
        qs = _split_heads_source(q_t)
        ks = _split_heads_source(k_tt.args[0])
        pv = _only_user(sm)
        if qs is None or ks is None or pv is None or not _is_function(pv, torch.matmul) or pv.args[0] is not sm:
            continue
        vs = _split_heads_source(pv.args[1])
        if vs is None or not (qs[1:] == ks[1:] == vs[1:]):
            continue
        merge = _only_user(pv)
        if not (_is_method(merge, "transpose") and tuple(merge.args[1:]) in ((1, 2), (2, 1))):
            continue
        contig = _only_user(merge)
        final = _only_user(contig) if _is_method(contig, "contiguous") else None
        if final is None:
            final = contig if _is_method(contig, "reshape") else None
        if not _is_method(final, "view", "reshape"):
            continue
        fused = _call(gm, final, W.attention_wrapper, (qs[0], ks[0], vs[0], None, float(scale), qs[1], qs[2]))
        final.replace_all_uses_with(fused)
        n += 1
    _finish(gm)
    return n


def fuse_qkv_projection(gm: fx.GraphModule) -> int:
    """Bias-free q/k/v Linears that share an input become one GEMM over row-concatenated weights
    (the author's planned fusion, optimizations.txt:22; kernels/attention_proj.py:52-155).  The fused
    weight is a non-persistent buffer on the GraphModule, so `state_dict()` is unchanged; q/k/v are
    column slices of the fused output, consumed by the attention kernel through strides."""
    n = 0
    for att in list(gm.graph.nodes):
        if not _is_function(att, W.attention_wrapper):
            continue
        qn, kn, vn = att.args[:3]
        mods = [_module_of(gm, x, nn.Linear) for x in (qn, kn, vn)]
        if any(m is None or m.bias is not None for m in mods):
            continue
        if any(len(x.users) != 1 for x in (qn, kn, vn)):
            continue
        group = None
        if qn.args[0] is kn.args[0] is vn.args[0]:
            group = [(qn, mods[0]), (kn, mods[1]), (vn, mods[2])]
        elif kn.args[0] is vn.args[0]:
            group = [(kn, mods[1]), (vn, mods[2])]
        if group is None or len({m.in_features for _, m in group}) != 1:
            continue
        name = f"_st_fused_proj_{n}"
        with torch.no_grad():
            gm.register_buffer(name, torch.cat([m.weight.detach() for _, m in group], dim=0).contiguous(),
                               persistent=False)
        parts, row = [], 0
        for node, m in group:
            parts.append((node.target, "weight", row, m.out_features))
            row += m.out_features
        _record_fused(gm, name, parts)
        first = group[0][0]
        with gm.graph.inserting_before(att):
            w = gm.graph.get_attr(name)
            fused = gm.graph.call_function(W.linear_wrapper_functional, (first.args[0], w, None, False))
            off = 0
            for node, m in group:
                piece = gm.graph.call_function(operator.getitem, (fused, (Ellipsis, slice(off, off + m.out_features))))
                off += m.out_features
                node.replace_all_uses_with(piece)
        n += 1
    _finish(gm)
    return n


def fuse_shared_input_projections(gm: fx.GraphModule) -> int:
    """Every bias-free projection of one and the same tensor becomes a single GEMM.  In the SDXL UNet this
    collapses the 70 cross-attention K/V projections of `encoder_hidden_states` (M = 154 rows: far too small
    to fill the machine one by one) into one (154 x 2048) . (2048 x 166 400) GEMM
    whose output every attention layer reads through strided column slices."""
    groups: Dict[Node, List[Node]] = {}
    for node in gm.graph.nodes:
        if _is_function(node, W.linear_wrapper_functional) and node.args[2] is None and node.args[3] is False \
                and isinstance(node.args[1], Node) and node.args[1].op == "get_attr":
            groups.setdefault(node.args[0], []).append(node)
    n = 0
    for src, nodes in groups.items():
        if len(nodes) < 2:
            continue
        weights = [gm.get_buffer(x.args[1].target) for x in nodes]
        if len({w.shape[1] for w in weights}) != 1:
            continue
        name = f"_st_shared_proj_{n}"
        with torch.no_grad():
            gm.register_buffer(name, torch.cat(weights, dim=0).contiguous(), persistent=False)
        parts, row = [], 0
        for x, wt in zip(nodes, weights):  # the per-layer buffers' sources move into the shared buffer
            for qualname, attr, off, rows in getattr(gm, "_st_fused_sources", {}).get(x.args[1].target, []):
                parts.append((qualname, attr, row + off, rows))
            row += wt.shape[0]
        _record_fused(gm, name, parts)
        first = nodes[0]
        with gm.graph.inserting_before(first):
            w = gm.graph.get_attr(name)
            fused = gm.graph.call_function(W.linear_wrapper_functional, (src, w, None, False))
        off = 0
        for node, wt in zip(nodes, weights):
            with gm.graph.inserting_before(node):
                piece = gm.graph.call_function(operator.getitem, (fused, (Ellipsis, slice(off, off + wt.shape[0]))))
            off += wt.shape[0]
            node.replace_all_uses_with(piece)
        for x in nodes:  # the per-layer fused weights are no longer referenced
            gm.graph.erase_node(x)
        n += 1
    _finish(gm)
    for name in [k for k, _ in gm.named_buffers() if k.startswith("_st_fused_proj_")]:
        if not any(nd.op == "get_attr" and nd.target == name for nd in gm.graph.nodes):
            delattr(gm, name)
            getattr(gm, "_st_fused_sources", {}).pop(name, None)
    return n


def fuse_time_embedding_projections(gm: fx.GraphModule) -> int:
    """Linear(SiLU(emb)) for every resnet (17 in SDXL, unet_pt.py:81-82) reads the same `emb`: one tiny-M GEMM
    over the row-concatenated weights (13 760 x 1280), each resnet taking its column slice."""
    groups: Dict[Node, List[Node]] = {}
    for node in gm.graph.nodes:
        m = _module_of(gm, node, nn.Linear)
        if m is None or m.bias is None:
            continue
        src = node.args[0]
        if _is_silu(gm, src) and isinstance(src.args[0], Node):
            groups.setdefault(src.args[0], []).append(node)
    n = 0
    for emb, nodes in groups.items():
        mods = [gm.get_submodule(x.target) for x in nodes]
        if len(nodes) < 2 or len({m.in_features for m in mods}) != 1:
            continue
        wname, bname = f"_st_temb_proj_w_{n}", f"_st_temb_proj_b_{n}"
        with torch.no_grad():
            gm.register_buffer(wname, torch.cat([m.weight.detach() for m in mods], dim=0).contiguous(), persistent=False)
            gm.register_buffer(bname, torch.cat([m.bias.detach() for m in mods], dim=0).contiguous(), persistent=False)
        wparts, bparts, row = [], [], 0
        for x, m in zip(nodes, mods):
            wparts.append((x.target, "weight", row, m.out_features))
            bparts.append((x.target, "bias", row, m.out_features))
            row += m.out_features
        _record_fused(gm, wname, wparts)
        _record_fused(gm, bname, bparts)
        first = min(nodes, key=lambda x: list(gm.graph.nodes).index(x))
        with gm.graph.inserting_before(first):
            w, b = gm.graph.get_attr(wname), gm.graph.get_attr(bname)
            fused = gm.graph.call_function(W.linear_wrapper_functional, (emb, w, b, False), {"silu_input": True})
        off = 0
        for node, m in zip(nodes, mods):
            with gm.graph.inserting_before(node):
                piece = gm.graph.call_function(operator.getitem, (fused, (Ellipsis, slice(off, off + m.out_features))))
            off += m.out_features
            node.replace_all_uses_with(piece)
        n += 1
    _finish(gm)
    return n


def _geglu_parts(gm, mul: Node):
    """mul == state * gelu(gate) -> (state, gate) or None."""
    if not (_is_function(mul, *_MUL) and len(mul.args) == 2):
        return None
    for state, act in (mul.args, mul.args[::-1]):
        if isinstance(act, Node) and (_is_function(act, F.gelu, torch._C._nn.gelu) or _module_of(gm, act, nn.GELU)):
            if act.kwargs.get("approximate", "none") != "none":
                continue
            return state, act.args[0]
    return None


def fuse_linear_geglu(gm: fx.GraphModule) -> int:
    """proj -> chunk(2, -1) -> state * gelu(gate)  ==>  one GEMM with a GEGLU epilogue
    (unet_pt.py:155-158); the (M, 8C) projection never reaches HBM."""
    n = 0
    for mul in list(gm.graph.nodes):
        parts = _geglu_parts(gm, mul)
        if parts is None:
            continue
        state, gate = parts
        if not (_is_function(state, operator.getitem) and _is_function(gate, operator.getitem)):
            continue
        chunk = state.args[0]
        if chunk is not gate.args[0] or state.args[1] != 0 or gate.args[1] != 1:
            continue
        if not (_is_method(chunk, "chunk") and chunk.args[1] == 2 and len(chunk.users) == 2):
            continue
        dim = chunk.args[2] if len(chunk.args) > 2 else chunk.kwargs.get("dim", 0)
        proj = chunk.args[0]
        if dim != -1 or _module_of(gm, proj, nn.Linear) is None or len(proj.users) != 1:
            continue
        if len(state.users) != 1 or len(gate.users) != 1:
            continue
        mod = _module_attr(gm, mul, proj.target)
        fused = _call(gm, mul, W.linear_geglu_wrapper, (proj.args[0], mod))
        mul.replace_all_uses_with(fused)
        n += 1
    _finish(gm)
    return n


def fuse_geglu(gm: fx.GraphModule) -> int:
    """Any remaining `state * gelu(gate)`  ==>  geglu_wrapper (reference: replace_geglu.py:33-41)."""
    n = 0
    for mul in list(gm.graph.nodes):
        parts = _geglu_parts(gm, mul)
        if parts is None:
            continue
        fused = _call(gm, mul, W.geglu_wrapper, parts)
        mul.replace_all_uses_with(fused)
        n += 1
    _finish(gm)
    return n


def fuse_proj_out_residual(gm: fx.GraphModule) -> int:
    """Transformer2DModel tail (unet_pt.py:236-243): proj_out -> reshape(B,H,W,C) -> permute(0,3,1,2)
    -> contiguous -> + res.  With channels-last activations both layout ops are views, so the image
    residual is added inside the proj_out GEMM epilogue."""
    n = 0
    for add in list(gm.graph.nodes):
        if not (_is_function(add, *_ADD) and len(add.args) == 2):
            continue
        for img, res in (add.args, add.args[::-1]):
            if not (_is_method(img, "contiguous") and len(img.users) == 1):
                continue
            perm = img.args[0]
            if not (_is_method(perm, "permute") and tuple(perm.args[1:]) == (0, 3, 1, 2) and len(perm.users) == 1):
                continue
            resh = perm.args[0]
            if not (_is_method(resh, "reshape", "view") and len(resh.args) == 5 and len(resh.users) == 1):
                continue
            lin = resh.args[0]
            if _module_of(gm, lin, nn.Linear) is None or len(lin.users) != 1 or not isinstance(res, Node):
                continue
            b, h, w, c = resh.args[1:]
            with gm.graph.inserting_before(add):
                res_tok = gm.graph.call_method("permute", (res, 0, 2, 3, 1))
                res_tok = gm.graph.call_method("reshape", (res_tok, b, -1, c))
                mod = gm.graph.get_attr(lin.target)
                y = gm.graph.call_function(W.linear_wrapper, (lin.args[0], mod, False), {"residual": res_tok})
                y = gm.graph.call_method("reshape", (y, b, h, w, c))
                y = gm.graph.call_method("permute", (y, 0, 3, 1, 2))
            add.replace_all_uses_with(y)
            n += 1
            break
    _finish(gm)
    return n


def fuse_linear_residual(gm: fx.GraphModule) -> int:
    """Linear(x) + r  ==>  linear_wrapper(x, linear, False, residual=r): the attention / feed-forward
    output projections and their skip connections (unet_pt.py:194,203,209)."""
    n = 0
    for add in list(gm.graph.nodes):
        if not (_is_function(add, *_ADD) and len(add.args) == 2 and not add.kwargs):
            continue
        for lin, res in (add.args, add.args[::-1]):
            if _module_of(gm, lin, nn.Linear) is None or len(lin.users) != 1 or not isinstance(res, Node):
                continue
            mod = _module_attr(gm, add, lin.target)
            fused = _call(gm, add, W.linear_wrapper, (lin.args[0], mod, False), {"residual": res})
            add.replace_all_uses_with(fused)
            n += 1
            break
    _finish(gm)
    return n


def replace_linear_activ(gm: fx.GraphModule, activation: Optional[nn.Module] = None) -> int:
    """SiLU(Linear(x))  ==>  linear_wrapper(x, linear, True)   (reference: replace_linear.py:59-77), and
    Linear(SiLU(x))     ==>  linear_wrapper(x, linear, False, silu_input=True)  (unet_pt.py:81-82)."""
    if activation is not None and not isinstance(activation, nn.SiLU):
        raise ValueError("replace_linear_activ: only SiLU is supported")
    n = 0
    for lin in list(gm.graph.nodes):
        if _module_of(gm, lin, nn.Linear) is None:
            continue
        user = _only_user(lin)
        src = lin.args[0]
        if user is not None and _is_silu(gm, user):
            mod = _module_attr(gm, user, lin.target)
            fused = _call(gm, user, W.linear_wrapper, (src, mod, True))
            user.replace_all_uses_with(fused)
            n += 1
        elif _is_silu(gm, src):
            mod = _module_attr(gm, lin, lin.target)
            fused = _call(gm, lin, W.linear_wrapper, (src.args[0], mod, False), {"silu_input": True})
            lin.replace_all_uses_with(fused)
            n += 1
    _finish(gm)
    return n


def _replace_module_calls(gm: fx.GraphModule, cls, build: Callable[[Node, Node, nn.Module], Node]) -> int:
    n = 0
    for node in list(gm.graph.nodes):
        m = _module_of(gm, node, cls)
        if m is None:
            continue
        mod = _module_attr(gm, node, node.target)
        node.replace_all_uses_with(build(node, mod, m))
        n += 1
    _finish(gm)
    return n


def replace_linear(gm: fx.GraphModule) -> int:
    """Every remaining Linear  ==>  linear_wrapper(x, linear, False) (reference: replace_linear.py:40-57;
    disabled upstream because its Triton GEMM was slower than cuBLAS, optimization.py:18-20)."""
    return _replace_module_calls(
        gm, nn.Linear, lambda node, mod, m: _call(gm, node, W.linear_wrapper, (node.args[0], mod, False)))


def replace_group_norm_activation(gm: fx.GraphModule, activation: Optional[nn.Module] = None) -> int:
    """SiLU(GroupNorm(x))  ==>  group_norm_wrapper(x, gn, True) (reference: replace_groupnorm.py:42-61)."""
    if activation is not None and not isinstance(activation, nn.SiLU):
        raise ValueError("replace_group_norm_activation: only SiLU is supported")
    n = 0
    for gn in list(gm.graph.nodes):
        if _module_of(gm, gn, nn.GroupNorm) is None:
            continue
        user = _only_user(gn)
        if user is None or not _is_silu(gm, user):
            continue
        mod = _module_attr(gm, user, gn.target)
        fused = _call(gm, user, W.group_norm_wrapper, (gn.args[0], mod, True))
        user.replace_all_uses_with(fused)
        n += 1
    _finish(gm)
    return n


def replace_group_norm(gm: fx.GraphModule) -> int:
    """GroupNorm(x)  ==>  group_norm_wrapper(x, gn, False) (reference: replace_groupnorm.py:23-40)."""
    return _replace_module_calls(
        gm, nn.GroupNorm, lambda node, mod, m: _call(gm, node, W.group_norm_wrapper, (node.args[0], mod, False)))


def replace_layer_norm(gm: fx.GraphModule) -> int:
    """LayerNorm(x)  ==>  layer_norm_wrapper(x, ln) (reference: replace_layernorm.py:30-47)."""
    return _replace_module_calls(
        gm, nn.LayerNorm, lambda node, mod, m: _call(gm, node, W.layer_norm_wrapper, (node.args[0], mod)))


def _conv_input(gm, conv: Node) -> Tuple[Node, bool]:
    """(x, upsample) where upsample says x was followed by a nearest-2x F.interpolate feeding only conv."""
    src = conv.args[0]
    if _is_function(src, F.interpolate) and len(src.users) == 1:
        sf = src.kwargs.get("scale_factor", src.args[2] if len(src.args) > 2 else None)
        mode = src.kwargs.get("mode", src.args[3] if len(src.args) > 3 else "nearest")
        size = src.kwargs.get("size", src.args[1] if len(src.args) > 1 else None)
        if size is None and sf in (2, 2.0) and mode == "nearest":
            return src.args[0], True
    return src, False


def fuse_conv_epilogues(gm: fx.GraphModule) -> int:
    """Conv2d(x) + temb[:, :, None, None]  and  shortcut + Conv2d(x)  ==>  conv2d_wrapper with the add
    fused into the implicit-GEMM epilogue (unet_pt.py:82-83, :93)."""
    n = 0
    for add in list(gm.graph.nodes):
        if not (_is_function(add, *_ADD) and len(add.args) == 2 and not add.kwargs):
            continue
        a, b = add.args
        # prefer fusing into the second operand (input + conv2(h)), then the first (conv1(h) + temb)
        for conv, other in ((b, a), (a, b)):
            m = _module_of(gm, conv, nn.Conv2d)
            if m is None or len(conv.users) != 1 or not isinstance(other, Node):
                continue
            if m.kernel_size != (3, 3) or m.stride != (1, 1) or m.padding != (1, 1) or m.in_channels % 64 != 0:
                continue
            if _is_function(other, operator.getitem) and other.args[1] == (slice(None), slice(None), None, None):
                kwargs = {"temb": other.args[0]}
            else:
                kwargs = {"residual": other}
            x, up = _conv_input(gm, conv)
            if up:
                kwargs["upsample"] = True
            mod = _module_attr(gm, add, conv.target)
            fused = _call(gm, add, W.conv2d_wrapper, (x, mod), kwargs)
            add.replace_all_uses_with(fused)
            n += 1
            break
    _finish(gm)
    return n


def replace_conv(gm: fx.GraphModule) -> int:
    """Every remaining Conv2d (conv_in/out, shortcuts, down/up-samplers)  ==>  conv2d_wrapper; a preceding
    nearest-2x interpolate is folded in (unet_pt.py:265-266)."""
    n = 0
    for node in list(gm.graph.nodes):
        if _module_of(gm, node, nn.Conv2d) is None:
            continue
        x, up = _conv_input(gm, node)
        mod = _module_attr(gm, node, node.target)
        fused = _call(gm, node, W.conv2d_wrapper, (x, mod), {"upsample": True} if up else None)
        node.replace_all_uses_with(fused)
        n += 1
    _finish(gm)
    return n


def replace_cat(gm: fx.GraphModule) -> int:
    """torch.cat([a, b], dim=1) of feature maps (skip connections, unet_pt.py:356,385)  ==>  concat_wrapper."""
    n = 0
    for node in list(gm.graph.nodes):
        if not _is_function(node, torch.cat, torch.concat):
            continue
        tensors = node.args[0]
        dim = node.kwargs.get("dim", node.args[1] if len(node.args) > 1 else 0)
        if dim != 1 or not isinstance(tensors, (list, tuple)) or len(tensors) != 2:
            continue
        fused = _call(gm, node, W.concat_wrapper, (tensors[0], tensors[1]))
        node.replace_all_uses_with(fused)
        n += 1
    _finish(gm)
    return n


def fuse_group_norm_statistics(gm: fx.GraphModule) -> int:
    """GroupNorm statistics come from the producer's epilogue.  Every GroupNorm of the UNet reads a tensor written by
    a conv / GEMM launch (resnet conv1 + temb -> norm2; resnet conv2 + shortcut, a transformer's proj_out + residual,
    conv_in, a down- / up-sampler -> the next norm1 / transformer norm / conv_norm_out) or the channel concatenation of
    two such tensors (up blocks, unet_pt.py:356,385).  Those producers become `*_stats_wrapper` calls that return
    `(y, partials)`, and the GroupNorm receives the partials: it no longer makes a statistics pass over the activation
    (46 launches and one full read of every normalised tensor per step).  Runs after every other pass."""
    stats_of: Dict[Node, Node] = {}  # tensor node (the value other nodes consume) -> its partials node

    def producer_stats(t: Node) -> Optional[Node]:
        if t in stats_of:
            return stats_of[t]
        if _is_function(t, operator.getitem) and t.args[1] == 0 \
                and _is_function(t.args[0], W.conv2d_stats_wrapper, W.linear_stats_wrapper):
            for u in t.args[0].users:  # a producer converted for an earlier GroupNorm (skip connections are read twice)
                if _is_function(u, operator.getitem) and u.args[1] == 1:
                    return u
        # y = linear_wrapper(...).reshape(b, h, w, c).permute(0, 3, 1, 2): the transformer tail (fuse_proj_out_residual)
        lin, chain = t, []
        if _is_method(t, "permute") and tuple(t.args[1:]) == (0, 3, 1, 2) and _is_method(t.args[0], "reshape", "view") \
                and len(t.args[0].args) == 5 and len(t.args[0].users) == 1:
            lin, chain = t.args[0].args[0], [t.args[0]]
        if _is_function(lin, W.conv2d_wrapper) and lin is t:
            target, extra = W.conv2d_stats_wrapper, {}
        elif _is_function(lin, W.linear_wrapper) and chain and not lin.kwargs.get("silu_input", False) \
                and len(lin.users) == 1:
            target, extra = W.linear_stats_wrapper, {}
        else:
            return None
        with gm.graph.inserting_before(lin):
            both = gm.graph.call_function(target, lin.args, {**lin.kwargs, **extra})
            y = gm.graph.call_function(operator.getitem, (both, 0))
            part = gm.graph.call_function(operator.getitem, (both, 1))
        lin.replace_all_uses_with(y)
        gm.graph.erase_node(lin)
        stats_of[t] = part
        return part

    n = 0
    for gn in list(gm.graph.nodes):
        if not _is_function(gn, W.group_norm_wrapper) or "partials" in gn.kwargs or len(gn.args) > 3:
            continue
        src = gn.args[0]
        sources = list(src.args[:2]) if _is_function(src, W.concat_wrapper) else [src]
        if not all(isinstance(x, Node) for x in sources):
            continue
        parts = [producer_stats(x) for x in sources]
        if any(p is None for p in parts):
            continue
        gn.kwargs = {**gn.kwargs, "partials": tuple(parts)}
        n += 1
    _finish(gm)
    return n


def replace_timesteps(gm: fx.GraphModule) -> int:
    """Timesteps(t)  ==>  timestep_wrapper(t, num_channels).  (The reference's `fuse_timesteps` pattern
    matches nothing on its own model; here the module is kept a leaf while tracing.)"""
    n = 0
    for node in list(gm.graph.nodes):
        if node.op != "call_module":
            continue
        m = gm.get_submodule(node.target)
        if type(m).__name__ != "Timesteps":
            continue
        fused = _call(gm, node, W.timestep_wrapper, (node.args[0], int(m.num_channels)))
        node.replace_all_uses_with(fused)
        n += 1
    _finish(gm)
    return n


def keep_channels_last(gm: fx.GraphModule) -> int:
    """`x.permute(0, 3, 1, 2).contiguous()` would copy an NHWC tensor back to NCHW; ask for the
    channels-last format instead, which makes it a no-op for the tensors the kernels produce."""
    n = 0
    for node in gm.graph.nodes:
        if _is_method(node, "contiguous") and not node.kwargs and len(node.args) == 1:
            src = node.args[0]
            if _is_method(src, "permute") and tuple(src.args[1:]) == (0, 3, 1, 2):
                node.kwargs = {"memory_format": torch.channels_last}
                n += 1
    _finish(gm)
    return n


def split_prompt_constants(gm: fx.GraphModule, dynamic_inputs=("sample", "timesteps")):
    """Split the rewritten graph into a prompt-constant prologue and the per-step body.

    Everything that depends only on `encoder_hidden_states` / `added_cond_kwargs` -- the batched cross-attention K/V
    projection (one 154 x 166400 x 2048 GEMM for SDXL) and the text/time-ids embedding MLP -- is the same for all steps
    of a prompt (reference: unet_pt.py:123-132, 478-486 recompute it every step).  Returns

        prologue : GraphModule(encoder_hidden_states, added_cond_kwargs) -> tuple of boundary tensors
        body     : GraphModule(sample, timesteps, *boundary) -> [eps]
        n        : number of boundary tensors

    Both share parameters / buffers / submodules with `gm`; `gm` itself is left untouched, so the plain forward keeps
    the reference's four-argument signature.
    """
    g = gm.graph
    nodes = list(g.nodes)
    placeholders = [n for n in nodes if n.op == "placeholder"]
    dyn_roots = {n for n in placeholders if n.target in dynamic_inputs}
    static_roots = [n for n in placeholders if n not in dyn_roots]
    dynamic, prompt_dep = set(dyn_roots), set(static_roots)
    # `x.to(emb.dtype)` (unet_pt.py:486) makes the text/time-ids embedding depend on a step tensor through its dtype
    # only.  Every activation of the engine has the parameter dtype (compile() checks it), so such a dtype is a
    # constant of the model, not of the step.
    param_dtype = next(gm.parameters()).dtype
    dtype_consts = set()
    for n in nodes:  # topological order
        if n.op in ("placeholder", "output"):
            continue
        if n.op == "call_function" and n.target is getattr and len(n.args) == 2 and n.args[1] == "dtype":
            dtype_consts.add(n)
            continue
        ins = [a for a in n.all_input_nodes if a not in dtype_consts]
        if any(a in dynamic for a in ins):
            dynamic.add(n)
        elif any(a in prompt_dep for a in ins):
            prompt_dep.add(n)  # depends on the prompt only
        # else: "free" (get_attr and pure functions of constants) -- copied wherever it is needed
    output = next(n for n in nodes if n.op == "output")
    boundary = [n for n in nodes if n in prompt_dep and n.op != "placeholder"
                and any((u in dynamic) or (u is output) for u in n.users)]

    def build(is_body: bool):
        new = fx.Graph()
        env: Dict[Node, Node] = {}
        if is_body:
            for ph in placeholders:
                if ph in dyn_roots or any(u in dynamic for u in ph.users):
                    env[ph] = new.placeholder(ph.target, default_value=ph.args[0] if ph.args else inspect.Signature.empty)
            for i, b in enumerate(boundary):
                env[b] = new.placeholder(f"prompt_const_{i}")
        else:
            for ph in static_roots:
                env[ph] = new.placeholder(ph.target)

        def copy(n):
            if n in dtype_consts:
                return param_dtype
            if n in env:
                return env[n]
            assert n.op != "placeholder", f"{n.target} is not an input of this half"
            if is_body:
                assert n not in prompt_dep, f"prompt-constant node {n.name} leaked into the step body"
            else:
                assert n not in dynamic, f"step-dependent node {n.name} leaked into the prologue"
            for a in n.all_input_nodes:
                copy(a)
            env[n] = new.node_copy(n, lambda x: param_dtype if x in dtype_consts else env[x])
            return env[n]

        if is_body:
            for n in nodes:
                if n in dynamic and n.op != "placeholder" and n not in dtype_consts:
                    copy(n)
            new.output(fx.node.map_arg(output.args[0], lambda x: copy(x)))
        else:
            new.output(tuple(copy(b) for b in boundary))
        new.lint()
        return fx.GraphModule(gm, new, type(gm).__name__ + ("Step" if is_body else "PromptConstants"))

    return build(False), build(True), len(boundary)


def census(gm: fx.GraphModule) -> Dict[str, int]:
    """Count what is left in the graph: module types still called, and wrapper call sites."""
    out: Dict[str, int] = {}
    for node in gm.graph.nodes:
        if node.op == "call_module":
            key = "module:" + type(gm.get_submodule(node.target)).__name__
        elif node.op == "call_function" and getattr(node.target, "__module__", "") == W.__name__:
            key = node.target.__name__
        else:
            continue
        out[key] = out.get(key, 0) + 1
    return out
