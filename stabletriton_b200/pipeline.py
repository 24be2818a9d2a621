"""The denoise loop around the compiled UNet, and its data-parallel launcher.

What the reference delegates to Diffusers' `StableDiffusionXLPipeline.__call__`
(`implementations/Diffusers/load_sdxl_pipeline.py:39,46`): per step
`scale_model_input` -> UNet on the [uncond ; cond] pair -> classifier-free-guidance mix -> Euler step.
The reference notes that this eager pipeline code ate most of its UNet speed-up
(implementations/Diffusers/README.md:2); here the whole step -- scheduler included -- is ONE CUDA-graph
replay: the loop state (fp32 latents, step counter, current timestep, sigma table) lives on the device
and is advanced by tiny kernels inside the graph, so the same graph serves every step and the host
only issues `num_steps` replays.

Scheduler = Diffusers 0.21.2 `EulerDiscreteScheduler` (scaled-linear betas, "leading" spacing,
steps_offset 1, epsilon prediction), restated; third-party, so parity is pinned only against this
repo's oracle restatement (oracle/unet_oracle.py: euler_sigmas / denoise_loop).

Multi-GPU: one process per GPU, full weight replica, prompts sharded across ranks,
no collective inside the loop; one all-gather of the final latents at the end.  Special case
`cfg_split` (2 ranks, one prompt): rank 0 runs the uncond row, rank 1 the cond row, and the two eps rows
(131 KB each at 1024^2) meet inside the step -- either through an NCCL all-gather captured in the step graph
(`exchange="nccl"`) or through ONE kernel that stores the row into the peer GPU's memory over NVLink and applies
the Euler update as soon as the peer's row has landed (`exchange="peer"`, csrc/peer.cu).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import _cabi


def euler_schedule(num_inference_steps: int, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                   beta_end: float = 0.012, steps_offset: int = 1):
    """(timesteps fp32 [n], sigmas fp32 [n + 1], init_noise_sigma) -- float64 host arithmetic, no numpy."""
    n = num_train_timesteps
    s0, s1 = beta_start ** 0.5, beta_end ** 0.5
    alphas_cumprod, acc = [], 1.0
    for i in range(n):
        beta = (s0 + (s1 - s0) * i / (n - 1)) ** 2
        acc *= 1.0 - beta
        alphas_cumprod.append(acc)
    all_sigmas = [math.sqrt((1 - a) / a) for a in alphas_cumprod]
    ratio = n // num_inference_steps
    timesteps = [float(round(i * ratio)) + steps_offset for i in range(num_inference_steps)][::-1]
    sigmas = []
    for t in timesteps:  # linear interpolation on the integer grid (np.interp semantics, clamped)
        lo = min(max(int(math.floor(t)), 0), n - 1)
        hi = min(lo + 1, n - 1)
        frac = min(max(t - lo, 0.0), 1.0) if hi > lo else 0.0
        sigmas.append(all_sigmas[lo] * (1 - frac) + all_sigmas[hi] * frac)
    sigmas.append(0.0)
    init_noise_sigma = math.sqrt(max(sigmas) ** 2 + 1.0)
    return (torch.tensor(timesteps, dtype=torch.float32), torch.tensor(sigmas, dtype=torch.float32), init_noise_sigma)


class PeerExchange:
    """The pair of NVLink-mapped exchange slabs behind `st_cfg_exchange_euler_update` (csrc/peer.cu): this rank's slab
    (cudaMalloc, exported as a CUDA IPC handle) and the peer's, mapped into this process.  torch.distributed only carries
    the two 64-byte handles, once; the per-step data never touches NCCL or the host."""

    def __init__(self, n_elems: int, device, group=None):
        import ctypes

        import torch.distributed as dist

        assert dist.is_initialized() and dist.get_world_size(group) == 2, "the CFG split runs on exactly two ranks"
        self.device = torch.device(device)
        self.rank = dist.get_rank(group)
        self.group = group
        L = _cabi._load()
        with torch.cuda.device(self.device):
            # Every rank walks through the same collectives whatever fails locally (a rank that raised early would leave
            # its peer waiting in the next collective); the outcome is agreed on at the end.
            self.local = self.peer = None
            failure = None
            handle = ctypes.create_string_buffer(64)
            try:
                nbytes = int(L.st_peer_slab_bytes(n_elems))
                ptr = ctypes.c_void_p()
                _cabi.check(L.st_peer_alloc(nbytes, ctypes.byref(ptr)), "peer_alloc")
                self.local = ptr.value
                _cabi.check(L.st_peer_export(self.local, handle), "peer_export")
            except Exception as exc:  # noqa: BLE001
                failure = exc
            handles = [None, None]
            dist.all_gather_object(handles, None if failure else handle.raw, group=group)
            if failure is None and handles[1 - self.rank] is not None:
                try:
                    peer = ctypes.c_void_p()
                    _cabi.check(L.st_peer_import(handles[1 - self.rank], ctypes.byref(peer)), "peer_import")
                    self.peer = peer.value
                except Exception as exc:  # noqa: BLE001
                    failure = exc
            ok = torch.tensor([1 if (failure is None and self.peer is not None) else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also: both slabs are mapped before anybody stores
            if int(ok.item()) == 0:
                self.close()
                raise _cabi.StableTritonError(
                    f"peer-memory exchange is not available between the two ranks ({failure!r}); "
                    f"use DenoiseLoop(..., exchange='nccl')")
        self.n = n_elems

    def error(self) -> int:
        """Sequence number of an exchange that timed out waiting for the peer (0: none).  Host read: not inside capture."""
        import ctypes
        out = ctypes.c_uint(0)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            _cabi.check(_cabi._load().st_peer_error(self.local, ctypes.byref(out)), "peer_error")
        return int(out.value)

    def close(self) -> None:
        """Collective (both ranks): unmap the peer's slab, then free the local one."""
        L = _cabi._load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if self.peer is not None:
                _cabi.check(L.st_peer_close(self.peer), "peer_close")
            import torch.distributed as dist
            if dist.is_initialized():  # the exporter must not free a slab the importer still has mapped
                dist.barrier(self.group)
            if self.local is not None:
                _cabi.check(L.st_peer_free(self.local), "peer_free")
        self.local = self.peer = None


class DenoiseLoop:
    """Euler + CFG loop for `prompts` prompts at a fixed latent size, one graph replay per step.

    unet: the module returned by `compile()` (its un-graphed `eager_forward` is captured together with
    the scheduler kernels) or any callable `(sample, t, ctx, added_cond_kwargs) -> [eps]`.
    """

    def __init__(self, unet, prompts: int, latent_hw: int, num_steps: int = 30, guidance: float = 5.0,
                 in_channels: int = 4, device="cuda", cfg_row: Optional[int] = None, group=None,
                 hoist_prompt_constants: bool = True, exchange: str = "nccl"):
        self.unet_fn = getattr(unet, "eager_forward", unet)
        # compile() also exposes the graph split at the prompt / step boundary: K/V projections of the text context and
        # the text / time-ids embedding are computed in set_conditioning(), not in every step
        self.prepare_fn = getattr(unet, "prepare", None) if hoist_prompt_constants else None
        self.step_fn = getattr(unet, "step_forward", None) if self.prepare_fn is not None else None
        self.consts = None
        self.P, self.hw, self.steps, self.guidance = prompts, latent_hw, num_steps, float(guidance)
        self.C = in_channels
        self.device = torch.device(device)
        self.cfg_row = cfg_row  # None: both CFG rows locally; 0 / 1: this rank computes only uncond / cond
        self.group = group
        timesteps, sigmas, self.init_noise_sigma = euler_schedule(num_steps)
        dev = self.device
        # one trailing entry so that advance_step after the last step reads inside the table
        self.timesteps = torch.cat([timesteps, timesteps.new_zeros(1)]).to(dev)
        self.sigmas = sigmas.to(dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.t_cur = torch.empty((), dtype=torch.float32, device=dev)
        self.x = torch.empty((prompts, in_channels, latent_hw, latent_hw), dtype=torch.float32, device=dev)
        rows = prompts if cfg_row is not None else 2 * prompts
        self.model_in = torch.empty((rows, in_channels, latent_hw, latent_hw), dtype=torch.bfloat16, device=dev)
        self.ctx = None
        self.added: Dict[str, torch.Tensor] = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._eps_pair = None
        # CFG split: how the two eps rows meet every step -- "nccl": all_gather_into_tensor + the Euler kernel (two graph
        # nodes through NCCL); "peer": one kernel that stores the row into the peer GPU's memory over NVLink, waits for
        # the peer's row and applies the Euler update (csrc/peer.cu)
        assert exchange in ("nccl", "peer"), exchange
        self.exchange = exchange if cfg_row is not None else None
        self.peer: Optional[PeerExchange] = None
        if cfg_row is not None and exchange == "nccl":
            self._eps_pair = torch.empty((2,) + tuple(self.model_in.shape), dtype=torch.bfloat16, device=dev)
        elif cfg_row is not None:
            self.peer = PeerExchange(self.x.numel(), dev, group)

    # -- one step, as launched into the current stream ---------------------------------------------
    def _step_body(self):
        L = _cabi.lib()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        n = self.x.numel()
        copies = 1 if self.cfg_row is not None else 2
        _cabi.check(L.st_scale_model_input(self.x.data_ptr(), self.model_in.data_ptr(), n, copies,
                                           self.sigmas.data_ptr(), self.step.data_ptr(), stream), "scale_model_input")
        if self.step_fn is not None:
            eps = self.step_fn(self.model_in, self.t_cur, *self.consts)[0]
        else:
            eps = self.unet_fn(self.model_in, self.t_cur, self.ctx, self.added)[0]
        if not eps.is_contiguous():
            eps = eps.contiguous()
        if self.peer is not None:
            _cabi.check(L.st_cfg_exchange_euler_update(eps.data_ptr(), int(self.cfg_row), self.peer.local, self.peer.peer,
                                                       self.x.data_ptr(), n, self.guidance, self.sigmas.data_ptr(),
                                                       self.step.data_ptr(), stream), "cfg_exchange_euler_update")
            _cabi.check(L.st_advance_step(self.step.data_ptr(), self.t_cur.data_ptr(), self.timesteps.data_ptr(), stream),
                        "advance_step")
            return eps
        if self.cfg_row is None:
            eps_u, eps_c = eps[: self.P], eps[self.P:]
        else:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self._eps_pair, eps, group=self.group)
            eps_u, eps_c = self._eps_pair[0], self._eps_pair[1]
        _cabi.check(L.st_euler_cfg_update(eps_u.data_ptr(), eps_c.data_ptr(), self.x.data_ptr(), n, self.guidance,
                                          self.sigmas.data_ptr(), self.step.data_ptr(), stream), "euler_cfg_update")
        _cabi.check(L.st_advance_step(self.step.data_ptr(), self.t_cur.data_ptr(), self.timesteps.data_ptr(), stream),
                    "advance_step")
        return eps

    def set_conditioning(self, cond: Dict[str, torch.Tensor], uncond: Dict[str, torch.Tensor]) -> None:
        """cond / uncond: encoder_hidden_states (P,77,D), text_embeds (P,E), time_ids (P,6), bf16 on device.
        Buffers are allocated once; later calls copy into them (the captured graph reads these addresses)."""
        def pick(key):
            if self.cfg_row is None:
                return torch.cat([uncond[key], cond[key]], dim=0)
            return (uncond if self.cfg_row == 0 else cond)[key]

        ctx = pick("encoder_hidden_states").to(self.device, torch.bfloat16)
        text = pick("text_embeds").to(self.device, torch.bfloat16)
        ids = pick("time_ids").to(self.device, torch.bfloat16)
        if self.ctx is None:
            self.ctx = ctx.clone()
            self.added = {"text_embeds": text.clone(), "time_ids": ids.clone()}
        else:
            self.ctx.copy_(ctx)
            self.added["text_embeds"].copy_(text)
            self.added["time_ids"].copy_(ids)
        if self.prepare_fn is not None:
            with torch.no_grad():
                fresh = self.prepare_fn(self.ctx, self.added)
            if self.consts is None:
                self.consts = list(fresh)  # these addresses are what the captured step graph reads
            else:
                for dst, src in zip(self.consts, fresh):
                    dst.copy_(src)

    def reset(self, latents: torch.Tensor) -> None:
        """latents: (P, C, H, W) unit-variance noise.  x0 = latents * init_noise_sigma; step = 0."""
        self.x.copy_(latents.to(self.device, torch.float32) * self.init_noise_sigma)
        self.step.zero_()
        self.t_cur.copy_(self.timesteps[0])

    def capture(self) -> None:
        """Warm up and capture one step into a CUDA graph (state is restored afterwards)."""
        assert self.ctx is not None, "call set_conditioning() first"
        saved = (self.x.clone(), self.step.clone(), self.t_cur.clone())
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(2):
                self.step.zero_()
                self._step_body()
            torch.cuda.synchronize(self.device)
            self.step.zero_()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
                self._last_eps = self._step_body()
        torch.cuda.synchronize(self.device)
        self.x.copy_(saved[0])
        self.step.copy_(saved[1])
        self.t_cur.copy_(saved[2])

    def run_step(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self._last_eps = self._step_body()

    @torch.no_grad()
    def run(self, latents: torch.Tensor, cond: Dict[str, torch.Tensor], uncond: Dict[str, torch.Tensor],
            use_graph: bool = True) -> torch.Tensor:
        """Full loop; returns the final fp32 latents (P, C, H, W)."""
        self.set_conditioning(cond, uncond)
        self.reset(latents)
        if use_graph and self.graph is None:
            self.capture()
        for _ in range(self.steps):
            self.run_step()
        return self.x.clone()


def shard_prompts(total_prompts: int, world_size: int, rank: int):
    """Contiguous prompt shard [lo, hi) of this rank; the remainder goes to the lowest ranks."""
    base, rem = divmod(total_prompts, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_latents(local: torch.Tensor, total_prompts: int, group=None) -> torch.Tensor:
    """All-gather the final latents of every rank (the only data-path collective of the prompt-sharded
    mode: 131 072 B per 1024^2 image in bf16, latency-bound).  Works with NCCL (GPU) and gloo (CPU)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [shard_prompts(total_prompts, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in counts)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = [out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(counts)]
    return torch.cat(pieces, dim=0)
