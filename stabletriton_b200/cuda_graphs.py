"""CUDA-graph replay of a callable with static input/output buffers.

Same contract as the reference's stable-fast derived wrapper (`optimizers/cuda/graphs.py`):

  * `make_dynamic_graphed_callable(fn)` keeps one captured graph per argument *signature* -- device,
    dtype and shape of every tensor, value of every Python scalar and of CPU scalar tensors
    (graphs.py:13-35, hash_arg :193-209) -- created lazily under a lock with double-checked lookup;
  * capture happens after 3 warm-up runs on a side stream (graphs.py:71-76), on a per-device
    execution environment that owns a dedicated stream, a shared private memory pool and a lock
    (graphs.py:156-190), so captures/replays from several threads serialise per device and eight
    single-GPU processes never interfere;
  * a call copies the arguments into the static input buffers, replays, and returns *copies* of the
    static outputs (graphs.py:128-138), so results stay valid across later calls.

Every kernel behind `compile()` is capture-safe by construction (no allocation, sync or host read in
the C ABI; TMA descriptors are encoded on the host while capturing and baked into the kernel nodes,
which is valid because replay reuses the same pool addresses).
"""
from __future__ import annotations

import functools
import logging
import threading
from typing import Any, Callable, Dict

import torch

logger = logging.getLogger(__name__)

_envs: Dict[int, "GraphExecutionEnv"] = {}
_envs_lock = threading.Lock()


class GraphExecutionEnv:
    """Per-device capture/replay context: stream + private mempool + lock."""

    def __init__(self, device: int):
        self.device = device
        with torch.cuda.device(device):
            self.stream = torch.cuda.Stream()
            self.mempool = torch.cuda.graphs.graph_pool_handle()
        self.lock = threading.RLock()


def get_per_device_graph_execution_env(device=None) -> GraphExecutionEnv:
    if isinstance(device, torch.device):
        device = device.index
    if device is None:
        device = torch.cuda.current_device()
    with _envs_lock:
        env = _envs.get(device)
        if env is None:
            env = _envs[device] = GraphExecutionEnv(device)
        return env


# ---- pytree helpers (tensors inside tuples / lists / dicts) ----------------------------------------
def hash_arg(arg: Any):
    if isinstance(arg, torch.Tensor):
        value = arg.item() if (arg.device.type == "cpu" and arg.numel() == 1) else None
        return ("T", arg.device.type, arg.device.index, arg.dtype, tuple(arg.shape), value)
    if isinstance(arg, (str, int, float, bytes, bool, type(None))):
        return arg
    if isinstance(arg, (tuple, list)):
        return tuple(hash_arg(a) for a in arg)
    if isinstance(arg, dict):
        return tuple(sorted(((hash_arg(k), hash_arg(v)) for k, v in arg.items()), key=lambda kv: repr(kv[0])))
    return ("O", type(arg).__name__)


def tree_map(fn: Callable[[torch.Tensor], Any], obj: Any):
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, (tuple, list)):
        return type(obj)(tree_map(fn, o) for o in obj)
    if isinstance(obj, dict):
        return {k: tree_map(fn, v) for k, v in obj.items()}
    return obj


def tree_copy_(dst: Any, src: Any) -> None:
    if isinstance(dst, torch.Tensor):
        if dst.device.type == "cuda":
            dst.copy_(src, non_blocking=True)
        return
    if isinstance(dst, (tuple, list)):
        if len(dst) != len(src):
            raise ValueError("graphed callable: argument structure changed")
        for d, s in zip(dst, src):
            tree_copy_(d, s)
    elif isinstance(dst, dict):
        if dst.keys() != src.keys():
            raise ValueError("graphed callable: argument keys changed")
        for k in dst:
            tree_copy_(dst[k], src[k])


def _first_cuda_device(obj: Any):
    found = []
    tree_map(lambda t: found.append(t.device.index) if t.device.type == "cuda" else None, obj)
    return found[0] if found else None


class GraphedCallable:
    """One captured graph + its static buffers."""

    def __init__(self, fn: Callable, args: tuple, kwargs: dict, env: GraphExecutionEnv, copy_outputs: bool = True):
        self.env = env
        self.copy_outputs = copy_outputs
        clone = lambda t: t.detach().clone() if t.device.type == "cuda" else t  # noqa: E731
        with env.lock, torch.cuda.device(env.device):
            torch.cuda.synchronize()
            with torch.cuda.stream(env.stream):
                self.static_args = tree_map(clone, args)
                self.static_kwargs = tree_map(clone, kwargs)
                for _ in range(3):  # lazy initialisation must not end up in the capture
                    fn(*self.static_args, **self.static_kwargs)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            # thread_local: another host thread driving ANOTHER device (its own env, its own capture or plain cudaMalloc)
            # must not invalidate this capture -- the default "global" mode does (first 2-GPU run of
            # test_two_devices_from_two_threads_of_one_process: "operation failed due to a previous error during capture")
            with torch.cuda.graph(self.graph, pool=env.mempool, stream=env.stream, capture_error_mode="thread_local"):
                self.static_outputs = fn(*self.static_args, **self.static_kwargs)
            torch.cuda.synchronize()

    def __call__(self, *args, **kwargs):
        env = self.env
        with env.lock:
            # order the input copies after the caller's stream, the replay after the copies, and the
            # caller's subsequent work after the replay
            caller = torch.cuda.current_stream(env.device)
            env.stream.wait_stream(caller)
            with torch.cuda.stream(env.stream):
                tree_copy_(self.static_args, args)
                tree_copy_(self.static_kwargs, kwargs)
                self.graph.replay()
                outputs = tree_map(lambda t: t.clone(), self.static_outputs) if self.copy_outputs \
                    else self.static_outputs
            caller.wait_stream(env.stream)
            return outputs


def make_dynamic_graphed_callable(fn: Callable, copy_outputs: bool = True) -> Callable:
    """Wrap `fn` so that every distinct argument signature is captured once and replayed afterwards."""
    lock = threading.Lock()
    cache: Dict[Any, GraphedCallable] = {}

    @functools.wraps(fn)
    def dynamic_graphed_callable(*args, **kwargs):
        key = (hash_arg(args), hash_arg(kwargs))
        graphed = cache.get(key)
        if graphed is None:
            with lock:
                graphed = cache.get(key)
                if graphed is None:
                    device = _first_cuda_device((args, kwargs))
                    if device is None:
                        raise ValueError("graphed callable: no CUDA tensor among the arguments")
                    logger.info("Dynamically graphing %s", getattr(fn, "__name__", type(fn).__name__))
                    graphed = GraphedCallable(fn, args, kwargs, get_per_device_graph_execution_env(device),
                                              copy_outputs=copy_outputs)
                    cache[key] = graphed
        return graphed(*args, **kwargs)

    dynamic_graphed_callable._cached = cache
    return dynamic_graphed_callable
