"""Multi-GPU host logic of the hot path (one process per GPU, torch.distributed).

GAF records are independent (filter-alignments.py:124 keeps no cross-record state but the
appends), so the file is cut into one contiguous byte range per rank, snapped to line ends;
the graph tables are replicated.  The only exchange is ONE all-reduce (sum) of the per-SV
REF/ALT counters before genotyping; hit tuples stay with their rank and are concatenated in
rank order = file order when informative_aln.json is written."""
from __future__ import annotations

import numpy as np


def shard_cuts(gaf, world):
    """world+1 byte offsets: rank r owns gaf[cuts[r]:cuts[r+1]].  Every cut but the last sits right
    behind a newline, so no record is split and the ranges tile the buffer."""
    a = gaf if isinstance(gaf, np.ndarray) else np.frombuffer(gaf, dtype=np.uint8)
    n = int(a.size)
    cuts = [0]
    for r in range(1, world):
        pos = max(cuts[-1], (n * r) // world)
        if pos >= n:
            cuts.append(n)
            continue
        if pos > 0 and a[pos - 1] == 10:          # already at a line start
            cuts.append(pos)
            continue
        # first newline at or after pos (searched in blocks so huge files are not scanned whole)
        nxt = n
        step = 1 << 20
        q = pos
        while q < n:
            blk = a[q:min(n, q + step)]
            hit = np.flatnonzero(blk == 10)
            if hit.size:
                nxt = q + int(hit[0]) + 1
                break
            q += step
        cuts.append(nxt)
    cuts.append(n)
    return cuts


def allreduce_counts(counts, group=None):
    """Sum of the [num_sv, 2] counters over all ranks, in place (NCCL for CUDA tensors, gloo for
    CPU ones).  Integer sums: bit-exact in any order."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return counts
    t = counts if isinstance(counts, torch.Tensor) else torch.from_numpy(counts)
    if t.dtype == torch.uint32:
        t = t.view(torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return counts


def gather_hits(hit_sv2, hit_off, hit_len, base_offset, group=None):
    """All ranks' hit tuples on rank 0 (None elsewhere), offsets made absolute in the whole file;
    concatenated in rank order."""
    import torch.distributed as dist
    mine = (np.asarray(hit_sv2, np.uint32), np.asarray(hit_off, np.uint64) + np.uint64(base_offset),
            np.asarray(hit_len, np.uint32))
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mine
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0, group=group)
    if rank != 0:
        return None
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))


class CounterExchange:
    """Per-SV counters of all ranks of one node, summed where they lie (include/svjg.h, svjg_xchg_*):
    every rank filters into ``counts_ptr(step)`` (three buffers taken in turn) and then calls ``genotype(step, ...)``, which
    announces the rank's counters, waits on the device for the other ranks and reads their counters
    over NVLink peer access -- no all-reduce.  ``exchange`` is a callable that all-gathers a bytes
    object across the ranks (e.g. a wrapper of torch.distributed.all_gather_object)."""

    def __init__(self, num_sv, rank, world, exchange):
        import ctypes as C
        from . import capi
        self._C, self._capi = C, capi
        self.num_sv, self.rank, self.world = int(num_sv), int(rank), int(world)
        base, handle = C.c_void_p(), C.create_string_buffer(64)
        capi.check(capi.lib.svjg_xchg_create(self.num_sv, C.byref(base), handle))
        self._base = base
        handles = exchange(handle.raw)
        self._regions = (C.c_void_p * self.world)()
        for q, h in enumerate(handles):
            if q == self.rank:
                self._regions[q] = base.value
            else:
                p = C.c_void_p()
                capi.check(capi.lib.svjg_xchg_open(h, C.byref(p)))
                self._regions[q] = p.value

    def counts_ptr(self, step):
        """Device address of this rank's counter buffer for ``step`` (1, 2, ...)."""
        return self._capi.lib.svjg_xchg_counts(self._base, self.num_sv, step % 3)

    def signal(self, step, stream):
        self._capi.check(self._capi.lib.svjg_xchg_signal(self._regions, self.world, self.rank, step, stream))

    def genotype(self, step, d_idx, d_ty, n, min_support, la, lb, lh, d_lut, lut_nmax, d_pl, d_gt, d_ad, d_fl, stream,
                 signal=True, k_override=None):
        """Asynchronous launch.  ``signal``: the kernel announces this rank's counters itself (no separate signal()
        launch).  The caller either uses :meth:`genotype_checked` or looks at the outcome itself: ``timed_out()``
        after a synchronise, and the SVJG_GT_NEED_K flag of the SVs whose counts lie beyond the log10 C(n,k) table."""
        self._capi.check(self._capi.lib.svjg_genotype_xchg(self._regions, self.world, self.rank, self.num_sv, step % 3, step,
                                                           1 if signal else 0, d_idx, d_ty, n, min_support, la, lb, lh, d_lut, lut_nmax,
                                                           k_override, d_pl, d_gt, d_ad, d_fl, stream))

    def genotype_checked(self, step, d_idx, d_ty, n, min_support, la, lb, lh, d_lut, lut_nmax, d_pl, d_gt, d_ad, d_fl, stream, sync):
        """:meth:`genotype`, then the checks the plain launch leaves to its caller: waits (``sync()``), raises if a
        peer never announced its counters (the kernel gives up after ~2 s and its sums are partial), and re-runs the
        SVs that need log10 C(n,k) beyond the table with CPython's own values, as the one-GPU path does
        (genotype.genotype_device).  ``d_fl`` / ``d_ad``: torch tensors (the flags and counts are read back)."""
        import numpy as np
        from . import genotype as G
        self.genotype(step, d_idx.data_ptr(), d_ty.data_ptr(), n, min_support, la, lb, lh, d_lut.data_ptr(), lut_nmax,
                      d_pl.data_ptr(), d_gt.data_ptr(), d_ad.data_ptr(), d_fl.data_ptr(), stream)
        sync()
        if self.timed_out():
            raise RuntimeError(f"rank {self.rank}: a peer did not announce its counters for step {step} (fused counter exchange)")
        flags = d_fl.cpu().numpy()[:n]
        if (flags & self._capi.GT_NEED_K).any():
            import torch
            kov = G._k_overrides(flags, d_ad.cpu().numpy().view(np.uint32)[:n], n)
            d_kov = torch.from_numpy(kov).to(d_fl.device)
            # the counters of this step are announced and summed where they lie: no second announcement
            self.genotype(step, d_idx.data_ptr(), d_ty.data_ptr(), n, min_support, la, lb, lh, d_lut.data_ptr(), lut_nmax,
                          d_pl.data_ptr(), d_gt.data_ptr(), d_ad.data_ptr(), d_fl.data_ptr(), stream, signal=False,
                          k_override=d_kov.data_ptr())
            sync()
            if (d_fl.cpu().numpy()[:n] & self._capi.GT_NEED_K).any():
                raise RuntimeError("genotype kernel could not represent a likelihood exactly")

    def timed_out(self):
        v = self._C.c_uint32()
        self._capi.check(self._capi.lib.svjg_xchg_timed_out(self._base, self._C.byref(v)))
        return bool(v.value)

    def close(self):
        if self._base:
            for q in range(self.world):
                if q != self.rank and self._regions[q]:
                    self._capi.lib.svjg_xchg_close(self._regions[q])
            self._capi.lib.svjg_xchg_free(self._base)
            self._base = None
