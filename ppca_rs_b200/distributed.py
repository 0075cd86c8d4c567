"""Sample-sharded EM: one process per GPU, one all-reduce of the additive statistics per iteration.

Every statistic of the EM iteration is a sum over independent samples (ppca_model.rs:290-293,303-306,350-358;
mix.rs:173,312-325), so the dataset is split into contiguous row blocks, one per rank, the model is replicated,
and per iteration each rank
    1. accumulates its local statistics buffer on the device       (ppca_b200_em_stats)
    2. joins ONE all-reduce(SUM) over that buffer                   (torch.distributed, NCCL over NVLink)
    3. finishes the M-step from the reduced buffer                  (ppca_b200_em_finish; replicated, O(d k^3))
Mixtures add one all-reduce(MAX) of m doubles for the per-component responsibility maxima (mix.rs:312-318).

The arithmetic is behind a small engine protocol so the host logic (sharding, buffer layout, reduction,
replicated finish) can be exercised with world_size-2 gloo tests on CPU tensors; the product engine is
`CudaEngine`, which has no CPU fallback.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat
from .model import Dataset, HostDataset, PPCAMix, PPCAModel, Prior


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block [lo, hi) of rank `rank` out of `world`: sizes differ by at most one row."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stats_len(d: int, k: int) -> int:
    """[A: d x kkp | B: d x kp | tdev: d | totals: d | 8 scalars] — ppca_b200_em_stats_len, restated."""
    kk = k * (k + 1) // 2
    kkp = (max(kk, 1) + 7) // 8 * 8
    kp = (max(k, 1) + 7) // 8 * 8
    return d * kkp + d * kp + 2 * d + 8


class CudaEngine:
    """The product engine: statistics on the device through the C ABI."""

    def __init__(self, ctx: nat.Context):
        import torch  # plumbing only: device buffers that NCCL can reduce
        self.torch = torch
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)
        # Everything torch does for this engine (buffer initialisation, the NCCL all-reduce) is enqueued on the CONTEXT's
        # stream, so it is ordered with the kernels of em_stats / em_finish without any host synchronisation
        # (ppca_b200_em_stats returns while its kernels are still running).
        self.stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=self.device)

    def stream_scope(self):
        return self.torch.cuda.stream(self.stream)

    def new_stats(self, d: int, k: int):
        n = nat.lib().ppca_b200_em_stats_len(d, k)
        assert n == stats_len(d, k)
        with self.stream_scope():
            return self.torch.zeros(n, dtype=self.torch.float64, device=self.device)

    def em_stats(self, ds, model: PPCAModel, stats) -> None:
        if isinstance(ds, HostDataset):  # this rank's rows stay in host memory and are streamed every step
            nat.check(nat.lib().ppca_b200_em_stats_host(self.ctx.handle, nat.dptr(ds._x), len(ds), ds._output_size(),
                                                        nat.dptr(ds._w), model.state_size, nat.dptr(model._C),
                                                        nat.dptr(model._mu), model._sigma,
                                                        C.c_void_p(stats.data_ptr())))
            return
        nat.check(nat.lib().ppca_b200_em_stats(self.ctx.handle, ds._h, model.state_size, nat.dptr(model._C),
                                               nat.dptr(model._mu), model._sigma, C.c_void_p(stats.data_ptr())))

    def em_finish(self, model: PPCAModel, prior: Optional[Prior], stats) -> Tuple[PPCAModel, float]:
        d, k = model.output_size, model.state_size
        C_out, mu_out = np.empty((d, k)), np.empty(d)
        s_out, llk = C.c_double(0.0), C.c_double(0.0)
        pr_ref, keep = None, None
        if prior is not None:
            pr, keep = prior._c(d)
            pr_ref = C.byref(pr)
        nat.check(nat.lib().ppca_b200_em_finish(self.ctx.handle, d, k, nat.dptr(model._C), nat.dptr(model._mu),
                                                model._sigma, pr_ref, C.c_void_p(stats.data_ptr()), nat.dptr(C_out),
                                                nat.dptr(mu_out), C.byref(s_out), C.byref(llk)))
        return PPCAModel(s_out.value, C_out, mu_out), llk.value

    # mixtures
    def new_logpost(self, n: int, m: int):
        with self.stream_scope():
            return self.torch.empty(max(n, 1) * m, dtype=self.torch.float64, device=self.device)

    def mix_posteriors(self, ds: Dataset, mix: PPCAMix, logpost) -> Tuple[np.ndarray, float]:
        ks, Cs, mus, sig, lw = mix._pack()
        cmax = np.empty(len(ks))
        llk = C.c_double(0.0)
        nat.check(nat.lib().ppca_b200_mix_posteriors(self.ctx.handle, ds._h, len(ks), ks.ctypes.data_as(nat.c_ip),
                                                     nat.dptr(Cs), nat.dptr(mus), nat.dptr(sig), nat.dptr(lw),
                                                     C.c_void_p(logpost.data_ptr()), nat.dptr(cmax), C.byref(llk)))
        return cmax, llk.value

    def mix_em_stats(self, ds: Dataset, mix: PPCAMix, j: int, logpost, cmax_j: float, stats) -> None:
        mj = mix._models[j]
        nat.check(nat.lib().ppca_b200_mix_em_stats(self.ctx.handle, ds._h, len(mix._models), j, mj.state_size,
                                                   nat.dptr(mj._C), nat.dptr(mj._mu), mj._sigma,
                                                   C.c_void_p(logpost.data_ptr()), float(cmax_j),
                                                   C.c_void_p(stats.data_ptr())))

    def to_tensor(self, a: np.ndarray):
        with self.stream_scope():
            return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def scalar_sumw(self, stats, d: int, k: int) -> float:
        return float(stats[stats_len(d, k) - 8 + 3].item())


def _scope(engine):
    """Stream scope of the engine's collectives (the product engine: the context's stream; CPU test engines: none)."""
    fn = getattr(engine, "stream_scope", None)
    return fn() if fn is not None else contextlib.nullcontext()


def native_comm_init(ctx: nat.Context, group) -> None:
    """Gives `ctx` the library's own NCCL communicator (ppca_b200_comm_init) over the ranks of a torch.distributed
    group: rank 0 creates the unique id, the group only ferries its 128 bytes."""
    if getattr(ctx, "comm_world", 1) > 1:
        return
    rank, world = group.get_rank(), group.get_world_size()
    box = [nat.Context.comm_unique_id() if rank == 0 else None]
    group.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)


class ShardedPPCA:
    """EM for one PPCAModel over a dataset sharded across the ranks of `group` (None = single process).

    collective="native" (the product default on GPUs): one call per step into ppca_b200_iterate_sharded /
    ppca_b200_iterate_host_sharded; statistics, NCCL all-reduce and finish all run inside the library on the context's
    stream.  collective="torch": the same protocol spelled out here (em_stats -> group.all_reduce -> em_finish), which is
    what the CPU (gloo) tests exercise with a numpy engine."""

    def __init__(self, ctx, dataset, model: PPCAModel, group=None, prior: Optional[Prior] = None, engine=None,
                 collective: Optional[str] = None):
        if collective is None:
            collective = "native" if engine is None else "torch"
        self.native = collective == "native"
        self.dataset, self.model, self.prior, self.group = dataset, model, prior, group
        self.last_llk = float("nan")
        if self.native:
            if group is not None:
                native_comm_init(ctx, group)
            return
        self.engine = engine or CudaEngine(ctx)
        self.stats = self.engine.new_stats(model.output_size, model.state_size)

    def step(self) -> float:
        """One EM iteration; returns the (global) log-likelihood of the model the step started from."""
        if self.native:  # one call into the library: iterate / iterate_host / iterate_packed_host (+ _sharded with a group)
            self.model, self.last_llk = self.model._iterate(self.dataset, self.prior, sharded=self.group is not None)
            return self.last_llk
        for _ in range(3):  # at most two climbs of the precision ladder (include/ppca_b200.h, PPCA_ERR_PRECISION)
            self.engine.em_stats(self.dataset, self.model, self.stats)
            if self.group is not None:
                with _scope(self.engine):
                    self.group.all_reduce(self.stats)  # SUM, enqueued behind the statistics kernels
            try:
                self.model, self.last_llk = self.engine.em_finish(self.model, self.prior, self.stats)
                break
            except nat.NativeError as e:
                # the guard counters ride in the reduced buffer, so every rank takes this branch together
                if e.code != nat.ERR_PRECISION:
                    raise
        return self.last_llk


class ShardedPPCAMix:
    """EM for a PPCAMix over a sharded dataset (mix.rs:281-337)."""

    def __init__(self, ctx, dataset, mix: PPCAMix, group=None, prior: Optional[Prior] = None, engine=None,
                 collective: Optional[str] = None):
        if collective is None:
            collective = "native" if engine is None else "torch"
        # "native": single-pass mixture EM inside the library (ppca_b200_mix_iterate[_sharded]); "torch": the
        # two-pass protocol below (posteriors, then one weighted pass per component), kept for the CPU (gloo) tests
        self.native = collective == "native"
        self.dataset, self.mix, self.prior, self.group = dataset, mix, prior, group
        self.last_llk = float("nan")
        if self.native:
            if group is not None:
                native_comm_init(ctx, group)
            return
        self.engine = engine or CudaEngine(ctx)
        d = mix.output_size
        self.stats = [self.engine.new_stats(d, k) for k in mix.state_sizes]
        self.logpost = self.engine.new_logpost(len(dataset), len(mix._models))

    def step(self) -> float:
        if self.native:
            self.mix, self.last_llk = self.mix._iterate(self.dataset, self.prior, sharded=self.group is not None)
            return self.last_llk
        eng, mix = self.engine, self.mix
        m, d = len(mix._models), mix.output_size
        cmax, llk = eng.mix_posteriors(self.dataset, mix, self.logpost)
        if self.group is not None:
            with _scope(eng):
                t = eng.to_tensor(np.concatenate([cmax, [llk]]))
                tmax = t[:m].clone()
                self.group.all_reduce(tmax, op=self.group.ReduceOp.MAX)
                tsum = t[m:].clone()
                self.group.all_reduce(tsum)
                cmax, llk = tmax.cpu().numpy(), float(tsum.item())
        models, logsum = [], np.empty(m)
        for j, mj in enumerate(mix._models):
            for _ in range(3):  # precision ladder, as in ShardedPPCA.step
                eng.mix_em_stats(self.dataset, mix, j, self.logpost, float(cmax[j]), self.stats[j])
                if self.group is not None:
                    with _scope(eng):
                        self.group.all_reduce(self.stats[j])
                try:
                    new, _ = eng.em_finish(mj, self.prior, self.stats[j])
                    break
                except nat.NativeError as e:
                    if e.code != nat.ERR_PRECISION:
                        raise
            models.append(new)
            logsum[j] = np.log(eng.scalar_sumw(self.stats[j], d, mj.state_size)) + cmax[j]  # mix.rs:323-324
        new_mix = PPCAMix.__new__(PPCAMix)
        new_mix._models = models
        mx = logsum.max()
        new_mix._logw = logsum - mx - np.log(np.sum(np.exp(logsum - mx)))  # mix.rs:335
        self.mix, self.last_llk = new_mix, llk
        return llk
