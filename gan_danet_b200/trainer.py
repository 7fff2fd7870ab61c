"""The GAN-DANet generator/discriminator training step on the B200 kernels.

Restates the hot loop of the reference's ``ModelTrainer.train`` (``/root/reference/GAN_DANet_train.ipynb:205-308``; the
same class is duplicated in ``deep_ensemble.ipynb:52-295``): input preparation (:226-232), G forward (:243), the D step
(:246-256), the G step (:259-269) with ``loss_G = (1-w)*MSE + w*BCE + TV + Perceptual`` and ``w = epoch/epochs``
(:266-267), AdamW for both nets (:182-183) and per-epoch ``CosineAnnealingWarmRestarts`` (:186-187,294-295).

Exact, work-saving deviations from the notebook (SURVEY appendix A identities):
  * the real and the detached fake batch go through D as one 2B batch in the D step (fc1's weight is streamed once);
    ``(BCE(real,1)+BCE(fake,0))/2`` equals the mean BCE over the 2B logits with targets [1..1,0..0];
  * D's parameters do not require grad during the G step: the notebook computes those gradients at :268 and discards
    them at :246 of the next iteration;
  * SSIM (:263) is evaluated only when ``eval_ssim=True`` -- it never enters the objective (:267).
Data-parallel training (not in the reference) all-reduces gradients over NCCL, see ``GradientAllReduce``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Iterable, List, Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from . import engine as E
from .models import Discriminator1, FlexibleUpsamplingModule, PerceptualLoss, SSIM, TVLoss
from .models.generator import BuildCtx
from .models.losses import BCEWithLogitsLoss, MSELoss


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, eps outside the sqrt) with one fused
    sm_100a kernel per parameter tensor (``gdn_adamw``).  ``grad_scale`` folds the 1/world of data parallelism.

    ``capturable = True`` (set by ``GraphedTrainStep``): ``step()`` neither counts steps nor passes step-dependent scalars by value -- the
    kernels read lr, lr / (1 - beta1^t) and sqrt(1 - beta2^t) from a 3-float device array per parameter group, which ``advance()`` refreshes
    (one tiny host -> device copy on the current stream) before every replay of the captured step.  The arithmetic is bit-identical to the eager
    path: the three scalars are computed in double precision and rounded to float32 in both."""

    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = 1.0
        self.capturable = False
        # data parallel, sharded big tensors (GradientAllReduce(shard_big=True)): parameter -> (lo, hi) element range this rank updates; its
        # exp_avg / exp_avg_sq exist for that range only (ZeRO-1 for Discriminator1.fc1: 1/world of the 7.5 GB AdamW pass and of its 2 GB state)
        self.shard_of = None
        self._t = 0                     # steps taken (capturable mode: one common counter, as every parameter steps every iteration)
        self._dyn: List[torch.Tensor] = []
        self._dyn_host: List[torch.Tensor] = []

    SMALL = 1 << 20

    def _range(self, p) -> Optional[Tuple[int, int]]:
        return self.shard_of(p) if self.shard_of is not None else None

    def _init_state(self, p) -> None:
        st = self.state[p]
        st["step"] = 0
        r = self._range(p)
        if r is None:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        else:
            st["exp_avg"] = torch.zeros(r[1] - r[0], dtype=p.dtype, device=p.device)
            st["exp_avg_sq"] = torch.zeros(r[1] - r[0], dtype=p.dtype, device=p.device)
            st["shard"] = r

    def _ensure_state(self) -> None:
        for group in self.param_groups:
            for p in group["params"]:
                if not self.state[p]:
                    self._init_state(p)

    def make_capturable(self) -> None:
        """Switch to device-resident step scalars (call after at least one eager step, before the capture)."""
        self._ensure_state()
        steps = {int(self.state[p]["step"]) for g in self.param_groups for p in g["params"]}
        assert len(steps) == 1, "capturable FusedAdamW needs all parameters at the same step"
        self._t = steps.pop()
        dev = self.param_groups[0]["params"][0].device
        self._dyn = [torch.zeros(4, dtype=torch.float32, device=dev) for _ in self.param_groups]
        self._dyn_host = [torch.zeros(4, dtype=torch.float32).pin_memory() for _ in self.param_groups]
        self.capturable = True

    def advance(self) -> None:
        """Capturable mode: count one step and upload its scalars (outside the graph, on the current stream, before the replay)."""
        assert self.capturable
        self._t += 1
        for group, dyn, host in zip(self.param_groups, self._dyn, self._dyn_host):
            # gdn_adamw receives lr and the betas as C floats and widens them to double for the bias corrections: mirror that exactly
            # (float32-rounded inputs, double arithmetic, float32-rounded results) so that the captured step is bit-identical to the eager one
            f32 = lambda v: float(torch.tensor(float(v), dtype=torch.float32))   # noqa: E731
            b1, b2 = (f32(b) for b in group["betas"])
            lr = f32(group["lr"])
            host[0] = lr
            host[1] = lr / (1.0 - math.pow(b1, float(self._t)))
            host[2] = math.sqrt(1.0 - math.pow(b2, float(self._t)))
            dyn.copy_(host, non_blocking=True)
            for p in group["params"]:
                self.state[p]["step"] = self._t

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        import ctypes as C
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            small = {}          # step -> list of (p, g, m, v, n): tensors below SMALL share one launch per 64 (gdn_adamw_multi)
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    assert not self.capturable
                    self._init_state(p)
                if not self.capturable:
                    st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                assert p.is_contiguous()
                r = self._range(p)
                if r is not None:          # this rank's slice of a sharded tensor (the gradient slice holds the reduce-scattered sum)
                    assert st.get("shard") == r, "the shard of a parameter changed after its optimizer state was created"
                    pw, gw, n = p.data.view(-1)[r[0]:r[1]], g.view(-1)[r[0]:r[1]], r[1] - r[0]
                    if self.capturable:
                        L.check(E._lib(p).gdn_adamw_dyn(pw.data_ptr(), gw.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), n, self._dyn[gi].data_ptr(),
                                                        float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), float(self.grad_scale), E._stream()),
                                "gdn_adamw_dyn")
                    else:
                        L.check(E._lib(p).gdn_adamw(pw.data_ptr(), gw.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), n,
                                                    float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                                    int(st["step"]), float(self.grad_scale), E._stream()), "gdn_adamw")
                    continue
                if p.numel() < self.SMALL:
                    small.setdefault(0 if self.capturable else int(st["step"]), []).append(
                        (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), p, g))
                    continue
                if self.capturable:
                    L.check(E._lib(p).gdn_adamw_dyn(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), self._dyn[gi].data_ptr(),
                                                    float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), float(self.grad_scale), E._stream()),
                            "gdn_adamw_dyn")
                else:
                    L.check(E._lib(p).gdn_adamw(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                                                float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                                int(st["step"]), float(self.grad_scale), E._stream()), "gdn_adamw")
            for step_no, items in small.items():
                k = len(items)
                arr = lambda j: (C.c_void_p * k)(*[it[j] for it in items])  # noqa: E731
                ns = (C.c_longlong * k)(*[it[4] for it in items])
                if self.capturable:
                    L.check(E._lib(items[0][5]).gdn_adamw_multi_dyn(k, arr(0), arr(1), arr(2), arr(3), ns, self._dyn[gi].data_ptr(), float(b1), float(b2),
                                                                    float(group["eps"]), float(group["weight_decay"]), float(self.grad_scale), E._stream()),
                            "gdn_adamw_multi_dyn")
                else:
                    L.check(E._lib(items[0][5]).gdn_adamw_multi(k, arr(0), arr(1), arr(2), arr(3), ns, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                                float(group["weight_decay"]), step_no, float(self.grad_scale), E._stream()), "gdn_adamw_multi")
        return loss


class GradientAllReduce:
    """Sums gradients over the data-parallel group (NCCL over NVLink 5 / NVSwitch on B200; gloo in CPU tests).

    Tensors above ``big`` elements are reduced in place one by one (D's fc1 gradient is ~1 GB and is its own bucket);
    the rest are packed into one flat bucket per call.  The 1/world factor is applied by the optimizer (``grad_scale``)."""

    def __init__(self, group=None, big: int = 1 << 20, own_group: bool = False, shard_big: bool = False):
        """``own_group``: create a process group (a communicator of its own) for the gradient reductions.  Needed when the step is captured in a CUDA
        graph: on this stack an EAGER collective issued on a communicator after collectives of that communicator were captured never completes
        (tools/probe_nccl_graph.py), so the captured reductions get a communicator nobody else uses -- barriers, broadcasts and the bench's timing
        reductions stay on the default group.  Collective call: every rank must construct the object."""
        import torch.distributed as dist
        if own_group and group is None and dist.is_initialized() and dist.get_world_size() > 1:
            group = dist.new_group()
        self.dist, self.group, self.big = dist, group, big
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # ``shard_big`` (ZeRO-1 for the big tensors): a tensor above ``big`` elements is REDUCE-SCATTERED in place -- rank r ends up with
        # the summed gradient of elements [r n / W, (r + 1) n / W) only -- the optimizer updates that slice (``shard_of``, FusedAdamW.shard_of) and
        # ``gather_params`` all-gathers the updated slices in place.  Same bytes on the wire as the all-reduce it replaces (a ring all-reduce IS a
        # reduce-scatter followed by an all-gather), but AdamW touches 1 / W of Discriminator1.fc1 (7.5 GB per step) and its state shrinks W-fold.
        self.shard_big = bool(shard_big) and self.world > 1
        self._shards: Dict[int, Tuple[int, int]] = {}

    def shard_of(self, p) -> Optional[Tuple[int, int]]:
        """(lo, hi) element range of ``p`` this rank owns, or None when ``p`` is replicated."""
        if not self.shard_big or p.numel() <= self.big or p.numel() % self.world != 0:
            return None
        n = p.numel() // self.world
        return (self.rank * n, (self.rank + 1) * n)

    def gather_params(self, params: Iterable[torch.nn.Parameter]) -> None:
        """All-gathers (in place) the parameter slices the ranks have just updated; call after the optimizer step."""
        if not self.shard_big:
            return
        handles = []
        for p in params:
            r = self.shard_of(p)
            if r is not None:
                flat = p.data.view(-1)
                handles.append(self.dist.all_gather_into_tensor(flat, flat[r[0]:r[1]], group=self.group, async_op=True))
        for h in handles:
            h.wait()

    def start(self, params: Iterable[torch.nn.Parameter]):
        """Launches the reductions (ordered after the work already queued on the current stream) and returns a ``finish`` callable.
        Kernels enqueued between ``start`` and ``finish`` overlap the collective (NCCL runs on its own stream)."""
        if self.world == 1:
            return lambda: None
        params = [p for p in params if p.grad is not None]
        grads = [p.grad for p in params]
        small = [g for g in grads if g.numel() <= self.big]
        handles = []
        for p, g in zip(params, grads):
            if g.numel() <= self.big:
                continue
            r = self.shard_of(p)
            if r is not None and g.is_contiguous():
                flat = g.view(-1)
                handles.append(self.dist.reduce_scatter_tensor(flat[r[0]:r[1]], flat, group=self.group, async_op=True))
            else:
                handles.append(self.dist.all_reduce(g, group=self.group, async_op=True))
        flat = None
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            handles.append(self.dist.all_reduce(flat, group=self.group, async_op=True))

        def finish() -> None:
            for h in handles:
                h.wait()
            if flat is not None:
                off = 0
                for g in small:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()

        return finish

    def __call__(self, params: Iterable[torch.nn.Parameter]) -> None:
        self.start(params)()


class HostBatchPipeline:
    """Double-buffered host -> device input pipeline (the DataLoader -> ``.to(device)`` hand-off of
    GAN_DANet_train.ipynb:225-228): the pinned host batch of step i+1 is copied on a side stream while step i computes, so
    the 12-24 MB/sample aux stack never stalls the kernels.  The device side is two PREALLOCATED staging sets (no per-step allocation, no
    allocator bookkeeping across streams): ``next()`` returns the staged tensors of the batch submitted last, safe to use on the current
    stream; ``release()`` -- called once the consumer's reads are enqueued -- lets the copy after next overwrite that set."""

    def __init__(self, device: torch.device, depth: int = 2):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.depth = depth
        self.slots: List[Optional[List[torch.Tensor]]] = [None] * depth
        self.free_ev: List[Optional[torch.cuda.Event]] = [None] * depth
        self._n = 0
        self._pending: Optional[Tuple[int, torch.cuda.Event]] = None
        self._held: Optional[int] = None

    def submit(self, host_tensors) -> None:
        """Start the asynchronous copy of one (pinned) host batch."""
        k = self._n % self.depth
        self._n += 1
        if self.slots[k] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(self.slots[k], host_tensors)):
            self.slots[k] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors]
            self.free_ev[k] = None
        with torch.cuda.stream(self.copy_stream):
            if self.free_ev[k] is not None:
                self.copy_stream.wait_event(self.free_ev[k])          # the consumer of this set's previous batch has finished reading it
            else:
                self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            for d, h in zip(self.slots[k], host_tensors):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._pending = (k, ev)

    def next(self):
        """Device tensors of the batch submitted last (waits for its copy on the current stream, not on the host)."""
        k, ev = self._pending
        self._pending = None
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._held = k
        return self.slots[k]

    def release(self) -> None:
        """The reads of the batch handed out by the last ``next()`` are enqueued on the current stream: its staging set may be refilled after them."""
        if self._held is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.free_ev[self._held] = ev
            self._held = None


def prepare_input_nhwc(lr_grace_05: torch.Tensor, hr_aux: torch.Tensor) -> torch.Tensor:
    """GAN_DANet_train.ipynb:226-232 fused: bicubic x0.5 of the 0.5-degree field and bicubic x0.25 of the aux stack,
    written straight into the channel slices of one NHWC generator input [B, h, w, 1 + C_aux]."""
    B, _, H2, W2 = lr_grace_05.shape
    _, Ca, H4, W4 = hr_aux.shape
    h, w = H2 // 2, W2 // 2
    assert (H4 // 4, W4 // 4) == (h, w), "lr_grace_05 and hr_aux grids are inconsistent"
    lib = E._lib(hr_aux)
    x = torch.empty((B, h, w, 1 + Ca), dtype=torch.float32, device=hr_aux.device)
    a, g = hr_aux.contiguous(), lr_grace_05.contiguous()
    L.check(lib.gdn_bicubic_down_nchw_to_nhwc(g.data_ptr(), x.data_ptr(), 1 + Ca, 0, B, 1, H2, W2, 2, E._stream()), "bicubic_down(grace)")
    if a.dtype == torch.bfloat16:       # the aux stack transported as bf16 (half the host -> device bytes); taps widened to fp32 on the device
        L.check(lib.gdn_bicubic_down_nchw_to_nhwc_bf16(a.data_ptr(), x.data_ptr(), 1 + Ca, 1, B, Ca, H4, W4, 4, E._stream()), "bicubic_down(aux, bf16)")
    else:
        L.check(lib.gdn_bicubic_down_nchw_to_nhwc(a.data_ptr(), x.data_ptr(), 1 + Ca, 1, B, Ca, H4, W4, 4, E._stream()), "bicubic_down(aux)")
    return x


def generator_forward_nhwc(G: FlexibleUpsamplingModule, x_nhwc: torch.Tensor) -> torch.Tensor:
    """G on an already NHWC input (skips the NCHW->NHWC boundary conversion); returns NCHW [B,1,4h,4w]."""
    params = list(G.parameters())

    def build(tape, xt, x_needs_grad, pvars):
        ctx = BuildCtx(tape, {id(p): v for p, v in zip(params, pvars)})
        xin = E.Var(xt, x_needs_grad)
        return G._build_nhwc(ctx, xin), xin

    return E.TapeFunction.apply(build, x_nhwc, *params)


def cosine_warm_restarts_lr(epoch: int, base_lr: float, t0: int = 10, t_mult: int = 2, eta_min: float = 1e-6) -> float:
    """CosineAnnealingWarmRestarts(T_0=10, T_mult=2, eta_min=1e-6) stepped per epoch (GAN_DANet_train.ipynb:186-187)."""
    t_i, t_cur = t0, epoch
    while t_cur >= t_i:
        t_cur -= t_i
        t_i *= t_mult
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2


class GANTrainer:
    """One object = the notebook's ``ModelTrainer`` state: G, D, both AdamW optimizers, schedulers and losses."""

    def __init__(self, G: FlexibleUpsamplingModule, D: Discriminator1, perceptual: Optional[PerceptualLoss], *, epochs: int = 150,
                 lr_g: float = 2e-4, lr_d: float = 4e-4, betas: Tuple[float, float] = (0.5, 0.999), weight_decay: float = 1e-4,
                 tv_weight: float = 1e-5, fused_adamw: bool = True, eval_ssim: bool = False, allreduce: Optional[GradientAllReduce] = None):
        self.G, self.D, self.perceptual = G, D, perceptual
        self.epochs, self.epoch = epochs, 0
        self.lr_g, self.lr_d, self.betas, self.weight_decay = lr_g, lr_d, betas, weight_decay
        self.fused_adamw = fused_adamw
        self.opt_G = self._make_opt(G.parameters(), lr_g)
        self.opt_D: Optional[torch.optim.Optimizer] = None      # created after fc1 is materialised by the first D forward
        self.bce, self.mse, self.tv, self.ssim = BCEWithLogitsLoss(), MSELoss(), TVLoss(tv_weight), SSIM()
        self.eval_ssim = eval_ssim
        self.allreduce = allreduce
        if allreduce is not None and allreduce.shard_big and not fused_adamw:
            raise L.GdnError("GradientAllReduce(shard_big=True) needs the fused AdamW (torch.optim.AdamW keeps full-size state)")
        if allreduce is not None and fused_adamw:
            self.opt_G.grad_scale = 1.0 / allreduce.world
            if allreduce.shard_big:
                self.opt_G.shard_of = allreduce.shard_of
        self._w_dev: Optional[torch.Tensor] = None       # [w, 1 - w] on the device: set by GraphedTrainStep (the captured step must not bake epoch / epochs)
        # Optional side-stream overlaps (off by default: measured on B200, graph replay, batch 32: D's AdamW -- a 7.5 GB HBM-bound pass -- beside the
        # tensor-bound VGG19 forward passes: 52.66 -> 52.51 ms; additionally the perceptual target branch beside G's forward: 53.06 ms, i.e. SLOWER --
        # the step's kernels already fill the SMs, concurrency only adds L2 / scheduling contention)
        self.overlap_opt_D = os.environ.get("GDN_OVERLAP_OPT_D", "0") == "1"
        self.overlap_target = os.environ.get("GDN_OVERLAP_TARGET", "0") == "1"
        self._side: Optional[torch.cuda.Stream] = None

    def _make_opt(self, params, lr):
        cls = FusedAdamW if self.fused_adamw else torch.optim.AdamW
        return cls(params, lr=lr, betas=self.betas, weight_decay=self.weight_decay)

    def _ensure_opt_D(self, sample: torch.Tensor) -> None:
        if self.opt_D is None:
            self.D._materialise_fc1(sample)
            self.opt_D = self._make_opt(self.D.parameters(), self.lr_d)
            if self.allreduce is not None and self.fused_adamw:
                self.opt_D.grad_scale = 1.0 / self.allreduce.world
                if self.allreduce.shard_big:
                    self.opt_D.shard_of = self.allreduce.shard_of

    def end_epoch(self) -> None:
        """scheduler_D.step(); scheduler_U.step() (GAN_DANet_train.ipynb:294-295)."""
        self.epoch += 1
        for opt, base in ((self.opt_D, self.lr_d), (self.opt_G, self.lr_g)):
            if opt is not None:
                for g in opt.param_groups:
                    g["lr"] = cosine_warm_restarts_lr(self.epoch, base)

    def _reduce(self, params) -> None:
        if self.allreduce is not None:
            self.allreduce(params)
            if not self.fused_adamw and self.allreduce.world > 1:
                for p in params:
                    if p.grad is not None:
                        p.grad.mul_(1.0 / self.allreduce.world)

    def _step_D(self, d_params, finish_reduce) -> None:
        """optimizer_D.step() (GAN_DANet_train.ipynb:256) on the current stream, after the gradient reduction (data parallel)."""
        if finish_reduce is not None:
            finish_reduce()
            if not self.fused_adamw and self.allreduce.world > 1:
                for p in d_params:
                    if p.grad is not None:
                        p.grad.mul_(1.0 / self.allreduce.world)
        self.opt_D.step()
        if self.allreduce is not None:
            self.allreduce.gather_params(d_params)

    def train_step(self, lr_grace_05: torch.Tensor, lr_grace_025: torch.Tensor, hr_aux: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One iteration of the hot loop.  Inputs are device tensors (NCHW, as ``CustomDataset`` yields them).
        Returns the scalar losses as 0-dim device tensors (no host sync)."""
        G, D = self.G, self.D
        real = lr_grace_025
        B = real.shape[0]
        self._ensure_opt_D(real)
        if real.is_cuda and (self.overlap_opt_D or self.overlap_target) and (self._side is None or self._side.device != real.device):
            self._side = torch.cuda.Stream(device=real.device)
        if self.overlap_target and real.is_cuda and self.perceptual is not None:
            self.perceptual.prefetch_target(real, self._side)
        x = prepare_input_nhwc(lr_grace_05, hr_aux)                       # :226-232
        hr = generator_forward_nhwc(G, x)                                 # :243

        # ---- discriminator step (:246-256)
        self.opt_D.zero_grad(set_to_none=True)
        both = torch.cat([real, hr.detach()], dim=0)
        logits = D(both)
        target = torch.cat([torch.ones(B, 1, device=real.device), torch.zeros(B, 1, device=real.device)], dim=0)
        loss_D = self.bce(logits, target)
        loss_D.backward()
        d_params = list(D.parameters())
        # data parallel: D's gradient all-reduce (~1 GB: fc1) is launched here and awaited only before D's optimiser step; the terms of the
        # generator objective that do not involve D (pixel, TV, perceptual: two VGG19 forward passes) are evaluated while it is in flight
        side = None
        if self.overlap_opt_D and real.is_cuda:
            side = self._side
            side.wait_stream(torch.cuda.current_stream(real.device))
            with torch.cuda.stream(side):
                self._step_D(d_params, self.allreduce.start(d_params) if self.allreduce is not None else None)
            finish_reduce = None
        else:
            finish_reduce = self.allreduce.start(d_params) if self.allreduce is not None else None

        # ---- generator step (:259-269), D-independent terms
        self.opt_G.zero_grad(set_to_none=True)
        loss_pix = self.mse(hr, real)
        out: Dict[str, torch.Tensor] = {}
        if self.eval_ssim:
            out["ssim"] = 1 - self.ssim(hr, real)
        loss_tv = self.tv(hr)
        loss_perc = self.perceptual(hr, real) if self.perceptual is not None else None

        if side is not None:
            torch.cuda.current_stream(real.device).wait_stream(side)      # D's updated parameters are read from here on
        else:
            self._step_D(d_params, finish_reduce)

        # ---- adversarial term: D already updated, its parameter gradients are not needed here
        for p in d_params:
            p.requires_grad_(False)
        try:
            fake_out = D(hr)
        finally:
            for p in d_params:
                p.requires_grad_(True)
        loss_adv = self.bce(fake_out, torch.ones_like(fake_out))
        if self._w_dev is not None:
            loss_G = self._w_dev[1] * loss_pix + self._w_dev[0] * loss_adv + loss_tv
        else:
            w = self.epoch / self.epochs
            loss_G = (1 - w) * loss_pix + w * loss_adv + loss_tv
        if loss_perc is not None:
            loss_G = loss_G + loss_perc
        loss_G.backward()
        g_params = list(G.parameters())
        self._reduce(g_params)
        self.opt_G.step()
        if self.allreduce is not None:
            self.allreduce.gather_params(g_params)
        out.update({"loss_D": loss_D.detach(), "loss_G": loss_G.detach(), "adv": loss_adv.detach(), "pixel": loss_pix.detach(),
                    "tv": loss_tv.detach()})
        if loss_perc is not None:
            out["perceptual"] = loss_perc.detach()
        out["hr"] = hr.detach()
        return out


class GraphedTrainStep:
    """``GANTrainer.train_step`` captured in ONE CUDA graph (input preparation, G forward, D step, G step with all loss terms, both backward passes,
    both AdamW updates, BatchNorm running statistics): ~800 kernel launches replayed by a single ``cudaGraphLaunch`` instead of ~48 ms of host-side
    enqueue per step.  What changes between iterations lives in device memory, not in kernel arguments:

    * the batch: three static input buffers (``__call__`` copies the caller's device tensors into them);
    * AdamW's step-dependent scalars and the learning rate: ``FusedAdamW.advance()`` (3 floats per parameter group);
    * the adversarial weight ``w = epoch / epochs`` (GAN_DANet_train.ipynb:266): a 2-float device tensor refreshed when the epoch changes.

    Losses are bit-identical to the eager step (same kernels, same order, same scalars: tests/test_gpu_graph.py).  Data parallel: the NCCL
    all-reduces are captured with the step.  Everything the captured step allocates lives in the graph's private memory pool."""

    def __init__(self, trainer: "GANTrainer", lr_grace_05: torch.Tensor, lr_grace_025: torch.Tensor, hr_aux: torch.Tensor, warmup: int = 2):
        if not trainer.fused_adamw:
            raise L.GdnError("GraphedTrainStep needs the fused AdamW (torch.optim.AdamW's step counter is not capturable here)")
        self.tr = trainer
        dev = lr_grace_025.device
        self.static_in = [torch.empty_like(t) for t in (lr_grace_05, lr_grace_025, hr_aux)]
        for dst, src in zip(self.static_in, (lr_grace_05, lr_grace_025, hr_aux)):
            dst.copy_(src)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):            # real training steps: optimizer state, lazy fc1, persistent operand buffers now exist
                trainer.train_step(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        trainer.opt_G.make_capturable()
        trainer.opt_D.make_capturable()
        self._w_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        trainer._w_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        self._epoch = None
        self._refresh_scalars()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count
        # thread_local: ProcessGroupNCCL's watchdog thread may query events while this thread captures (data parallel)
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local" if trainer.allreduce is not None else "global"):
            self.static_out = trainer.train_step(*self.static_in)
        self.launches_per_step = L.launch_count - n0       # C-ABI calls recorded into the graph (each one or more kernels)

    def _refresh_scalars(self) -> None:
        tr = self.tr
        if self._epoch != tr.epoch:
            w = tr.epoch / tr.epochs
            self._w_host[0], self._w_host[1] = w, 1 - w        # float32(1 - w) computed in double, as the eager expression (1 - w) * loss does
            tr._w_dev.copy_(self._w_host, non_blocking=True)
            self._epoch = tr.epoch

    def __call__(self, lr_grace_05: torch.Tensor, lr_grace_025: torch.Tensor, hr_aux: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One training iteration.  Returns the step's losses / generated field as STATIC device tensors (overwritten by the next call)."""
        for dst, src in zip(self.static_in, (lr_grace_05, lr_grace_025, hr_aux)):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._refresh_scalars()
        self.tr.opt_D.advance()
        self.tr.opt_G.advance()
        self.graph.replay()
        return self.static_out


def init_like_reference(G: FlexibleUpsamplingModule, D: Discriminator1, sample_real: torch.Tensor, seed: Optional[int] = None) -> None:
    """The notebook's initialisation (GAN_DANet_train.ipynb:164-165) with the authors' effective behaviour:
    ``weights_init_normal`` on every conv / BN / materialised Linear; the lazy ``fc1`` keeps nn.Linear's default init."""
    from .models import weights_init_normal
    if seed is not None:
        torch.manual_seed(seed)
    G.apply(weights_init_normal)
    D.apply(weights_init_normal)
    D._materialise_fc1(sample_real)
