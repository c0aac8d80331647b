"""Training engine: one data-parallel step = H2D tiles -> forward -> weighted CE (+grad) -> backward -> gradient
all-reduce (NCCL over NVLink, N>1 only) -> optimizer -> bf16 weight re-staging; the whole device part is captured in
ONE CUDA graph (≈700 kernel launches per step would otherwise be bounded by host launch latency).

Mirrors what fastai's Learner does per batch for the reference (train.py:246-250 fit_one_cycle -> _do_one_batch:
pred = model(xb); loss = loss_func(pred, yb); loss.backward(); opt.step(); opt.zero_grad()).
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib, ops
from .network import UNetB200


def one_cycle(pct: float, lr_max: float, div: float = 25.0, div_final: float = 1e5, pct_start: float = 0.25,
              moms: Sequence[float] = (0.95, 0.85, 0.95)) -> Tuple[float, float]:
    """fastai fit_one_cycle schedule (combined cosine for lr and momentum)."""
    def cos(a, b, p):
        return a + (1 + math.cos(math.pi * (1 - p))) * (b - a) / 2
    if pct < pct_start:
        q = pct / pct_start
        return cos(lr_max / div, lr_max, q), cos(moms[0], moms[1], q)
    q = (pct - pct_start) / (1 - pct_start)
    return cos(lr_max, lr_max / div_final, q), cos(moms[1], moms[2], q)


# DIAGNOSTIC ONLY (bench A/B at N > 1): B2U_DIAG_SKIP_ALLREDUCE=1 leaves the gradients un-reduced, i.e. N independent
# replicas that still meet at the step barrier - what the slowest of N ranks costs without any collective.  Training
# with it is wrong by construction; Learner / train_func refuse to run with it set.
_SKIP_ALLREDUCE = os.environ.get("B2U_DIAG_SKIP_ALLREDUCE") is not None

class Trainer:
    def __init__(self, net: UNetB200, optimizer: str = "sgd", lr: float = 1e-3, wd: float = 0.01,
                 encoder_factor: float = 10.0, use_graph: bool = True, bucket_mb: float = 32.0,
                 input_dtype: torch.dtype = torch.uint8, grad_bf16: Optional[bool] = None):
        assert net.training
        self.net, self.lr, self.optimizer = net, lr, optimizer.lower()
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.use_graph = use_graph
        dev = net.device
        N, C, H, W = net.N, net.n_in, net.H, net.W
        self.x_static = torch.zeros((N, C, H, W), dtype=input_dtype, device=dev)   # raw band values (A0 input contract)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._prefetched = None
        self._stage_free: Optional[torch.cuda.Event] = None
        self.step_count = 0
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.overlap_allreduce = os.environ.get("B2U_NO_AR_OVERLAP") is None   # A/B switches
        # gradients travel as bf16 (half the all-reduce bytes; the cast kernels sit inside the captured graphs): the
        # default for N > 1, B2U_GRAD_FP32=1 (or grad_bf16=False) keeps the fp32 exchange - with 2 ranks that one is
        # bit-exact (a + b == b + a), which the multi-GPU parity test uses
        self.grad_bf16 = (self.world > 1 and os.environ.get("B2U_GRAD_FP32") is None) if grad_bf16 is None else bool(grad_bf16)
        self.gbf = torch.zeros(net.layout.total, dtype=torch.bfloat16, device=dev) if self.grad_bf16 else None
        self.min_seg_mb = float(os.environ.get("B2U_AR_MIN_SEG_MB", "24"))
        if os.environ.get("B2U_BUCKET_MB"):
            self.bucket_elems = int(float(os.environ["B2U_BUCKET_MB"]) * (1 << 20) / 4)
        if self.optimizer == "adam":
            L = net.layout
            self.m = torch.zeros_like(net.params)
            self.v = torch.zeros_like(net.params)
            ends, self._seg_group, wds = [], [], []
            for e in L.entries:
                ends.append(e.offset + e.numel)
                self._seg_group.append(e.group)
                wds.append(wd if e.decay else 0.0)
            ends[-1] = L.total
            self.seg_end = torch.tensor(ends, dtype=torch.int64, device=dev)
            self.seg_wd = torch.tensor(wds, dtype=torch.float32, device=dev)
            self.seg_lr = torch.zeros(len(ends), dtype=torch.float32, device=dev)
            self.hyper = torch.zeros(6, dtype=torch.float32, device=dev)
            self.encoder_factor = encoder_factor
            self.set_adam_hyper(lr, 0.9)
        elif self.optimizer != "sgd":
            raise ValueError("optimizer must be 'sgd' or 'adam'")

    # fastai: lr_max = slice(lr/encoder_factor, lr) -> geometric spacing over the 3 parameter groups (train.py:246-250)
    def set_adam_hyper(self, lr: float, mom: float, sqr_mom: float = 0.99, eps: float = 1e-5) -> None:
        lo = lr / self.encoder_factor
        group_lr = [lo, math.sqrt(lo * lr), lr]
        self.seg_lr.copy_(torch.tensor([group_lr[g] for g in self._seg_group], dtype=torch.float32), non_blocking=True)
        step = self.step_count + 1
        self.hyper.copy_(torch.tensor([mom, sqr_mom, eps, 1 - mom ** step, 1 - sqr_mom ** step, 1.0 / self.world],
                                      dtype=torch.float32), non_blocking=True)

    # ------------------------------------------------------------------------------------------------ device step
    def _allreduce(self, lo: int = 0, hi: Optional[int] = None, async_op: bool = False):
        if self.world == 1 or _SKIP_ALLREDUCE:
            return []
        g = self.gbf if self.grad_bf16 else self.net.grads
        hi = g.numel() if hi is None else hi
        # bucketed so that NCCL pipelines over NVLink; summed here, divided by world inside the optimizer kernel
        works = []
        for a, b in bucket_ranges(lo, hi, self.bucket_elems):
            w = dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, async_op=async_op)
            if async_op:
                works.append(w)
        return works

    def _device_step(self) -> None:
        self._fwd_bwd()
        self._allreduce()
        self._update()

    def _to_wire(self, lo: int, hi: int) -> None:
        """gradient range [lo, hi) -> its bf16 wire copy (no-op for the fp32 exchange); enqueued behind the kernels that
        produced it, i.e. captured at the end of its backward segment"""
        if self.grad_bf16 and hi > lo:
            lo4 = lo // 4 * 4         # keep the 16-byte alignment of the fp32 range (neighbours are rewritten identically)
            _lib.check(self.net.lib.b2u_cast_f32_bf16(self.net.grads[lo4:].data_ptr(), self.gbf[lo4:].data_ptr(), hi - lo4,
                                                      ops.stream_ptr()), "b2u_cast_f32_bf16")

    def _fwd_bwd(self, part: Optional[int] = None) -> None:
        """part None: everything; part i: segment i of self.segments (segment 0 also holds input cast + forward + loss)."""
        net = self.net
        s = ops.stream_ptr()
        if part in (None, 0):
            net.set_input(self.x_static, s)
            net.forward(s)
            net.loss_and_grad(s)
        if part is None:
            net.backward(s)
            if self.world > 1:
                self._to_wire(0, net.layout.total)
        else:
            b, e, lo, hi = self.segments[part]
            net.backward(s, b, e)
            self._to_wire(lo, hi)

    def _plan_segments(self) -> None:
        net = self.net
        self.segments = plan_segments(net.bwd_marks, net.layout.total, len(net.bwd_ops),
                                      int(self.min_seg_mb * (1 << 20) / 4))

    def _update(self) -> None:
        net = self.net
        s = ops.stream_ptr()
        if self.grad_bf16 and self.world > 1:
            _lib.check(net.lib.b2u_cast_bf16_f32(self.gbf.data_ptr(), net.grads.data_ptr(), net.layout.total, s),
                       "b2u_cast_bf16_f32")
        if self.optimizer == "sgd":
            _lib.check(net.lib.b2u_sgd_step(net.params.data_ptr(), net.grads.data_ptr(), net.layout.total, self.lr,
                                            1.0 / self.world, s), "b2u_sgd_step")
        else:
            _lib.check(net.lib.b2u_adam_step(net.params.data_ptr(), net.grads.data_ptr(), self.m.data_ptr(),
                                             self.v.data_ptr(), net.layout.total, self.seg_end.data_ptr(),
                                             self.seg_lr.data_ptr(), self.seg_wd.data_ptr(), self.seg_end.numel(),
                                             self.hyper.data_ptr(), s), "b2u_adam_step")
        net._stage_weights(s)

    def capture(self) -> None:
        """Warm up once eagerly on a side stream, then capture the device step into a CUDA graph."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            saved_p = self.net.params.clone()
            saved_b = {k: v.clone() for k, v in self.net.buffers.items()}
            if self.optimizer == "adam":
                sm, sv = self.m.clone(), self.v.clone()
            self._device_step()
            # the warm-up step must not count: restore parameters, BN buffers and optimizer state
            self.net.params.copy_(saved_p)
            for k, v in saved_b.items():
                self.net.buffers[k].copy_(v)
            if self.optimizer == "adam":
                self.m.copy_(sm)
                self.v.copy_(sv)
            self.net._stage_weights(ops.stream_ptr())
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # NCCL stays OUT of the graphs (an eager collective after a captured one dead-locked on this stack): with
        # N > 1 the step is graph(fwd+loss+bwd) -> eager bucketed all-reduce -> graph(optimizer + weight staging)
        self.graph = torch.cuda.CUDAGraph()
        # (measured: capturing the chain on a high-priority stream starves the weight-gradient stream until the chain is
        # done - 24.0 instead of 20.4 ms/step; both streams stay at the default priority)
        cap = torch.cuda.Stream()
        if self.world == 1:
            with torch.cuda.graph(self.graph, stream=cap):
                self._device_step()
        elif self.overlap_allreduce:
            # one graph per backward segment: [fwd + loss + decoder backward] -> async all-reduce of the decoder gradients
            # behind [encoder stage 7 backward] -> async all-reduce of stage 7 behind [stages 6..stem] -> all-reduce of
            # the rest -> [optimizer + weight staging]
            self._plan_segments()
            self.graphs = []
            for i in range(len(self.segments)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self.graphs[0].pool() if self.graphs else None, stream=cap):
                    self._fwd_bwd(i)
                self.graphs.append(g)
            self.graph = self.graphs[0]
            self.graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_update, pool=self.graph.pool(), stream=cap):
                self._update()
            return
        else:
            with torch.cuda.graph(self.graph, stream=cap):
                self._fwd_bwd()
            self.graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_update, pool=self.graph.pool(), stream=cap):
                self._update()

    def _load_batch(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """Inputs of this step into the static buffers the plan reads: from the prefetch staging buffers when this very
        batch was announced by the previous step (a device-to-device copy behind the copy stream's event), else
        directly (H2D from pinned memory or D2D)."""
        pf = self._prefetched
        if pf is not None and pf[0] is x and pf[1] is y:
            st = torch.cuda.current_stream()
            st.wait_event(pf[2])
            self.x_static.copy_(self._x_stage, non_blocking=True)
            self.net.labels.copy_(self._y_stage, non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(st)          # the staging buffers may be overwritten once these copies are done
        else:
            self.x_static.copy_(x, non_blocking=True)
            ld = self.net.labels.dtype
            self.net.labels.copy_(y if y.dtype == ld else y.to(ld), non_blocking=True)
        self._prefetched = None

    def _prefetch(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """H2D of the NEXT batch on a copy stream while this step computes (its 21 MB take ~0.4 ms of PCIe time)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.net.device)
            self._x_stage = torch.empty_like(self.x_static)
            self._y_stage = torch.empty_like(self.net.labels)
        cs = self._copy_stream
        if self._stage_free is not None:
            cs.wait_event(self._stage_free)      # NOT the compute stream as a whole: the copy must overlap this step
        with torch.cuda.stream(cs):
            self._x_stage.copy_(x, non_blocking=True)
            self._y_stage.copy_(y if y.dtype == self._y_stage.dtype else y.to(self._y_stage.dtype), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._prefetched = (x, y, ev)

    def step(self, x: torch.Tensor, y: torch.Tensor,
             prefetch: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
        """x: raw band values [N,C,H,W] of the trainer's `input_dtype` (uint8; uint16 / int16 for 16-bit imagery), host
        pinned or device; y: uint8/int64 class ids [N,H,W] (float32 targets for the regression variant). Returns the device loss scalar of this
        step (mean over the local batch, before the parameter update).  `prefetch=(x_next, y_next)` starts the
        host-to-device copy of the next batch behind this step's kernels; pass the same tensor objects to the next call."""
        if self.use_graph and self.graph is None:
            self._load_batch(x, y)
            self.capture()
            self._prefetched = None
        self._load_batch(x, y)
        if self.use_graph:
            if self.world > 1 and self.overlap_allreduce:
                works = []
                for i, (b, e, lo, hi) in enumerate(self.segments):
                    self.graphs[i].replay()
                    last = i == len(self.segments) - 1
                    # gradients [lo, hi) are final: reduce them while the next segment's backward runs
                    works += self._allreduce(lo, hi, async_op=not last)
                for w in works:
                    w.wait()                                            # stream-level wait, the host does not block
                self.graph_update.replay()
            else:
                self.graph.replay()
                if self.world > 1:
                    self._allreduce()
                    self.graph_update.replay()
        else:
            self._device_step()
        if prefetch is not None:
            self._prefetch(*prefetch)
        self.step_count += 1
        return self.net.loss


def bucket_ranges(lo: int, hi: int, bucket_elems: int) -> List[Tuple[int, int]]:
    """[lo, hi) of the flat gradient buffer cut into all-reduce buckets of at most `bucket_elems` elements"""
    return [(a, min(hi, a + bucket_elems)) for a in range(lo, hi, max(1, bucket_elems))]


def plan_segments(marks: Sequence[Tuple[int, int]], total: int, n_ops: int,
                  min_seg_elems: int) -> List[Tuple[int, int, int, int]]:
    """Backward segments [(op begin, op end, grad lo, grad hi)] from the network's all-reduce marks (op index k, offset
    off: after op k every gradient at offsets >= off is final).  After segment i the gradients [lo, hi) are final and
    their all-reduce can run behind segment i+1.  Marks that would leave fewer than `min_seg_elems` gradients in a
    segment are merged into the next one (launch latency, not bandwidth, is what a small bucket costs over NVSwitch).
    The segments tile the op list [0, n_ops) and the gradient buffer [0, total) exactly."""
    segs, prev_k, prev_off = [], 0, total
    for k, off in marks:
        if off <= 0 or prev_off - off < min_seg_elems:
            continue
        segs.append((prev_k, k, off, prev_off))
        prev_k, prev_off = k, off
    segs.append((prev_k, n_ops, 0, prev_off))
    return segs


def init_distributed() -> Tuple[int, int, int]:
    """One process per GPU (torchrun env). Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        # the persistent conv kernels occupy every SM: NCCL gets a small, fixed number of CTAs (NVSwitch bandwidth does not
        # need many channels; more CTAs only take SMs away from the backward pass the all-reduce hides behind)
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("B2U_NCCL_MAX_CTAS", "8"))
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return rank, local, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of n_items over ranks (first n_items % world ranks take one more)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
