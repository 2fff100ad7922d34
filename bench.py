#!/usr/bin/env python3
"""bench.py — throughput of the B200 hot path (GAF filter + allele counts +
genotype) on the BASELINE.json workload, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--scale 1.0]
    python bench.py --impl reference ...      # CPU arm: the oracle port on host cores

A "step" is one pass of the hot path over one batch of synthetic GAF: counters
reset -> filter chain (probe, scan_parse, link, exact kernels) -> (NCCL all-reduce of
the counters when N > 1) -> genotype kernel.  `value` = alignments/s with the batch resident in HBM;
`e2e` = the same through svjg_filter_host (pinned HOST bytes in, counts + hits +
genotypes back on the host, copies inside the timed region).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "svjedi-graph_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "gaf_alignments_filtered_assigned_per_sec"
# dram__bytes_read.sum + dram__bytes_write.sum of the filter chain per launch, from the ncu --set full
# capture committed under profiles/ (None where no capture exists for the workload)
FILTER_TRAFFIC = {"C2": 557_700_000}   # profiles/r1/ncu_chain_v15_summary.txt: scan_parse 517.4+18.7, link 21.4, probe+exact 0.2 MB
UNIT = "alignments/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--scale", type=float, default=None, help="shrink SV and record counts (default: full; C5: per-GPU shard)")
    ap.add_argument("--catalogue", default="scaled", choices=["scaled", "full"],
                    help="full: the SV catalogue (and genome) at the config's stated size whatever --scale says; "
                         "C5 then probes its 1 M-SV tables (459 MB, beyond L2) with a --scale shard of the records")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--kernel-only", action="store_true", help="developer mode: print the kernel times and stop")
    ap.add_argument("--cpu-sample", type=int, default=150_000, help="records the CPU baseline is timed on")
    ap.add_argument("--scan-blocks", type=int, default=0, help="developer: blocks per SM the scan kernel is sized for (6 or 8)")
    ap.add_argument("--tile-lines", type=int, default=0, help="developer: lines a tile of the scan kernel should hold")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: counters summed inside the genotype kernel over NVLink peer memory (p2p), or ncclAllReduce")
    return ap.parse_args()


def workload(args, rank):
    """Synthetic batch of the named config (per-GPU batch: weak scaling).
    Tables are identical on every rank; reads differ by rank."""
    from svjg import synth
    scale = args.scale
    if scale is None:
        # C5 (1M SVs / 200M records over 8 GPUs) is sized per GPU: 1/8 of the records
        scale = 1.0 if args.workload != "C5" else 0.02
    t0 = time.time()
    # synthetic inputs are deterministic; keep them for later runs on the same box (untimed either way)
    import pickle
    cscale = 1.0 if args.catalogue == "full" else None
    cache = os.path.join(os.environ.get("SVJG_CACHE", "/tmp"), f"svjg_wl_{args.workload}_{scale:g}_{args.catalogue}_{rank}.pkl")
    if os.path.exists(cache):
        with open(cache, "rb") as fh:
            g, vcf, gaf = pickle.load(fh)
    else:
        g, vcf, gaf = synth.make_workload(args.workload, scale=scale, stream0=rank, catalogue_scale=cscale)
        try:
            with open(cache + f".{os.getpid()}", "wb") as fh:
                pickle.dump((g, vcf, gaf), fh, protocol=pickle.HIGHEST_PROTOCOL)
            os.replace(cache + f".{os.getpid()}", cache)
        except Exception:
            pass
    import io
    buf = io.StringIO()
    g.write_gfa(buf)
    return g, vcf, gaf, buf.getvalue(), scale, time.time() - t0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.06] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(gaf_text, edges_text, gfa_text, vcf_text, n_sample):
    """The oracle port (oracle/svjg_oracle.py) timed on the first n_sample
    records of the same batch, single-threaded like the reference."""
    from oracle import svjg_oracle as O
    lines = []
    pos = 0
    for _ in range(n_sample):
        j = gaf_text.find("\n", pos)
        if j < 0:
            break
        lines.append(gaf_text[pos:j + 1])
        pos = j + 1
    edges = json.loads(edges_text)
    alt = {}
    for line in gfa_text.splitlines(True):
        if line.startswith("S"):
            c = line.split("\t")
            if "." in c[1].split(":")[-1]:
                alt[c[1]] = len(line.rstrip().split("\t")[2])
    t0 = time.perf_counter()
    d = O.filter_alignments(lines, edges, alt)
    O.dumps_informative(d)
    t1 = time.perf_counter()
    _, n_gt = O.genotype_vcf(O.hit_counts(d), vcf_text.splitlines(True))
    t2 = time.perf_counter()
    return len(lines), t1 - t0, t2 - t1, n_gt


def cpu_baseline_c(gaf_text, edges_text, gfa_text, n_sample=None, check=None):
    """The C restatement of the reference filter (oracle/svjg_oracle.c) on the host cores: filter only
    (no JSON text, no genotypes), one thread and all threads.  An extra line of context beside
    cpu_baseline — the reference itself is single-threaded Python — never the thing measured.
    ``check`` = (sv ids, counters [num_sv, 2], number of hits) of the GPU path for the same batch: the
    whole batch is then compared counter by counter (key "parity")."""
    try:
        from oracle import c_oracle as CO
        CO.ensure_built()
        alt = {}
        for line in gfa_text.splitlines(True):
            if line.startswith("S"):
                c = line.split("\t")
                if "." in c[1].split(":")[-1]:
                    alt[c[1]] = len(line.rstrip().split("\t")[2])
        t = CO.Tables(json.loads(edges_text), alt)
        if n_sample:
            pos = 0
            for _ in range(n_sample):
                j = gaf_text.find("\n", pos)
                if j < 0:
                    break
                pos = j + 1
            gaf_text = gaf_text[:pos]
        gaf = gaf_text.encode()
        n_rec = gaf.count(b"\n")
        cores = min(32, os.cpu_count() or 1)
        out = {"unit": UNIT, "kind": "port (C)", "records": n_rec, "what": "oracle/svjg_oracle.c, filter + counters + hit tuples only"}
        for label, th in (("one_thread", 1), ("all_threads", cores)):
            t0 = time.perf_counter()
            counts, st = CO.filter_counts(t, gaf, threads=th)
            out[label] = {"value": st["n_records"] / (time.perf_counter() - t0), "threads": th}
        if check is not None:
            ids, got, n_hits = check
            same = list(ids) == list(t.sv_ids) and got.shape == counts.shape and bool((got == counts).all()) \
                and int(n_hits) == st["n_hits"]
            out["parity"] = {"ok": same, "checked": f"all {n_rec} records of rank 0's batch: {counts.shape[0]} x 2 counters and the "
                                                     f"number of hits ({st['n_hits']}) against the C oracle, bit for bit"}
        return out
    except Exception as exc:                      # context only: never fails the bench
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}


_SHARD = {}


def _ref_shard(k):
    """One line shard through the oracle port (filter + json.dumps), in a forked worker."""
    from oracle import svjg_oracle as O
    t0 = time.perf_counter()
    d = O.filter_alignments(_SHARD["lines"][k], _SHARD["edges"], _SHARD["alt"])
    O.dumps_informative(d)
    return O.hit_counts(d), time.perf_counter() - t0


def run_reference(args):
    """CPU arm: the oracle port on ALL host cores — one process per core over a
    line-sharded sample (appends are in file order, so per-shard outputs
    concatenate to the single-process result)."""
    import multiprocessing as mp
    from oracle import svjg_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, vcf, gaf, gfa_text, scale, gen_s = workload(args, 0)
    edges_text = g.edges_json()
    vcf_lines = vcf.splitlines(True)
    n_vcf = sum(1 for l in vcf_lines if not l.startswith("#"))
    n_rec = gaf.count("\n")
    cores = max(1, os.cpu_count() or 1)
    per_step = max(1000 * cores, (args.cpu_sample * cores) // max(1, args.steps))
    per_step = min(per_step, n_rec)
    lines, pos = [], 0
    for _ in range(per_step):
        j = gaf.find("\n", pos)
        lines.append(gaf[pos:j + 1])
        pos = j + 1
    shard = (len(lines) + cores - 1) // cores
    _SHARD["lines"] = [lines[i:i + shard] for i in range(0, len(lines), shard)]
    _SHARD["edges"] = json.loads(edges_text)
    _SHARD["alt"] = {k: len(v) for k, v in g.alt_nodes.items()}
    n_shards = len(_SHARD["lines"])
    tfs, tgs = [], []
    with mp.get_context("fork").Pool(n_shards) as pool:
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            parts = pool.map(_ref_shard, range(n_shards), chunksize=1)
            counts = {}
            for c, _ in parts:
                for k, (a, b) in c.items():
                    x = counts.get(k, (0, 0))
                    counts[k] = (x[0] + a, x[1] + b)
            t1 = time.perf_counter()
            O.genotype_vcf(counts, vcf_lines)
            t2 = time.perf_counter()
            if i >= args.warmup:
                tfs.append(t1 - t0)
                tgs.append(t2 - t1)
    n = len(lines)
    # whole-batch projection: the filter scales with records, the genotyper with SVs
    proj_s = (sum(tfs) / len(tfs)) * n_rec / n + sum(tgs) / len(tgs)
    val = n_rec / proj_s
    line = {
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000 * (sum(tfs) + sum(tgs)) / len(tfs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int64", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"{args.workload} x{scale:g}: {n_rec} GAF records, {n_vcf} VCF SVs ({n} records timed per step)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": n_shards, "kind": "port",
                         "sample": f"{n} records per step over {n_shards} processes (filter + json.dumps), projected to "
                                   f"{n_rec} records, plus genotyping all {n_vcf} SVs once per step; oracle/svjg_oracle.py "
                                   "(the reference itself is single-threaded)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_baseline_c": cpu_baseline_c("".join(lines), edges_text, gfa_text),
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from svjg import alnfilter, capi, genotype

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.scan_blocks:
        capi.check(capi.lib.svjg_filter_tune(capi.TUNE_SCAN_BLOCKS, args.scan_blocks))
    if args.tile_lines:
        capi.check(capi.lib.svjg_filter_tune(capi.TUNE_TILE_LINES, args.tile_lines))
    g, vcf, gaf, gfa_text, scale, gen_s = workload(args, rank)
    edges_text = g.edges_json()
    tables = alnfilter.Tables.from_memory(edges_text, gfa_text).to_device(local)
    gaf_bytes = gaf.encode()
    n_bytes = len(gaf_bytes)
    n_rec = gaf_bytes.count(b"\n")
    h_gaf = torch.frombuffer(bytearray(gaf_bytes), dtype=torch.uint8).pin_memory()
    d_gaf = h_gaf.to(dev)

    # genotype inputs (host string work done once, untimed: it is per-catalogue, not per-read)
    header, recs = genotype.parse_vcf(vcf.splitlines(True))
    n_sv = len(recs)
    sv_idx = np.array([capi.NO_SV if (r[2] is None or tables.find_sv(r[2]) is None) else tables.find_sv(r[2]) for r in recs],
                      dtype=np.uint32)
    sv_ty = np.array([r[1] for r in recs], dtype=np.uint8)
    lo, hi = (n_sv * rank) // world, (n_sv * (rank + 1)) // world      # SV shard of this rank
    n_loc = hi - lo
    d_idx = torch.from_numpy(sv_idx[lo:hi].view(np.int32).copy()).to(dev)
    d_ty = torch.from_numpy(sv_ty[lo:hi].copy()).to(dev)
    lut = torch.from_numpy(genotype.log10comb_lut()).to(dev)
    la, lb, lh = math.log10(1 - genotype.ERR), math.log10(genotype.ERR), math.log10(1 / 2)
    d_pl = torch.empty((max(1, n_loc), 3), dtype=torch.int64, device=dev)
    d_gt = torch.empty(max(1, n_loc), dtype=torch.uint8, device=dev)
    d_ad = torch.empty((max(1, n_loc), 2), dtype=torch.int32, device=dev)
    d_fl = torch.empty(max(1, n_loc), dtype=torch.uint8, device=dev)
    pl_all = torch.empty((world, max(1, (n_sv + world - 1) // world), 3), dtype=torch.int64, device=dev) if world > 1 else None

    # size the hit buffers from one untimed pass
    filt = alnfilter.DeviceFilter(tables, hit_cap=1024, device=local)
    filt.reset()
    filt.run(d_gaf)                                   # the cursor counts every hit, stored or not
    st = filt.read_stats()
    if st["status"]:
        raise SystemExit(f"synthetic GAF rejected: {st}")
    n_hits = st["n_hits"]
    filt = alnfilter.DeviceFilter(tables, hit_cap=n_hits + 1024, device=local)
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)
    lib = capi.lib

    xchg = None
    if world > 1 and args.collective == "p2p":
        from svjg.shard import CounterExchange

        def gather_bytes(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        xchg = CounterExchange(tables.num_sv, rank, world, gather_bytes)
    step_no = [0]

    def step(ev=None):
        step_no[0] += 1
        k = step_no[0]
        # p2p: this rank's counters live in its exchange region (two buffers, alternating by step)
        counts_ptr = xchg.counts_ptr(k) if xchg else filt.counts.data_ptr()
        capi.check(lib.svjg_filter_reset(counts_ptr, tables.num_sv, filt.stats.data_ptr(), sp))
        if ev:
            ev[0].record(stream)
        capi.check(lib.svjg_filter_device(tables._h, d_gaf.data_ptr(), n_bytes, 0, 100, counts_ptr,
                                          filt.hit_sv2.data_ptr(), filt.hit_off.data_ptr(), filt.hit_len.data_ptr(),
                                          filt.hit_cap, filt.stats.data_ptr(), sp))
        if ev:
            ev[1].record(stream)
        if world > 1 and not xchg:
            dist.all_reduce(filt.counts)                   # per-SV REF/ALT counters, NCCL sum over NVLink
        if ev:
            ev[2].record(stream)
        if xchg:
            # announces this rank's counters, waits on the device for all ranks, then sums their counters
            # where they lie (NVLink peer reads): all-reduce and genotype step in one kernel
            xchg.genotype(k, d_idx.data_ptr(), d_ty.data_ptr(), n_loc, 3, la, lb, lh, lut.data_ptr(), genotype.LUT_NMAX,
                          d_pl.data_ptr(), d_gt.data_ptr(), d_ad.data_ptr(), d_fl.data_ptr(), sp)
        else:
            capi.check(lib.svjg_genotype_device(filt.counts.data_ptr(), d_idx.data_ptr(), d_ty.data_ptr(), n_loc, 3, la, lb, lh,
                                                lut.data_ptr(), genotype.LUT_NMAX, None, d_pl.data_ptr(), d_gt.data_ptr(),
                                                d_ad.data_ptr(), d_fl.data_ptr(), sp))
        if ev:
            ev[3].record(stream)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    if xchg:
        # untimed self-check of the fused exchange: NCCL all-reduce + plain genotype kernel must give the same
        capi.check(lib.svjg_filter_reset(filt.counts.data_ptr(), tables.num_sv, filt.stats.data_ptr(), sp))
        capi.check(lib.svjg_filter_device(tables._h, d_gaf.data_ptr(), n_bytes, 0, 100, filt.counts.data_ptr(),
                                          filt.hit_sv2.data_ptr(), filt.hit_off.data_ptr(), filt.hit_len.data_ptr(),
                                          filt.hit_cap, filt.stats.data_ptr(), sp))
        dist.all_reduce(filt.counts)
        capi.check(lib.svjg_genotype_device(filt.counts.data_ptr(), d_idx.data_ptr(), d_ty.data_ptr(), n_loc, 3, la, lb, lh,
                                            lut.data_ptr(), genotype.LUT_NMAX, None, d_pl.data_ptr(), d_gt.data_ptr(),
                                            d_ad.data_ptr(), d_fl.data_ptr(), sp))
        want = (d_pl.clone(), d_gt.clone(), d_ad.clone(), d_fl.clone())
        d_pl.zero_(), d_gt.zero_(), d_ad.zero_(), d_fl.zero_()
        step()
        torch.cuda.synchronize(dev)
        got = (d_pl, d_gt, d_ad, d_fl)
        if xchg.timed_out() or not all(torch.equal(a, b) for a, b in zip(want, got)):
            raise SystemExit(f"rank {rank}: fused counter exchange disagrees with NCCL all-reduce + genotype")
    for _ in range(max(3, args.warmup)):
        step()
    sync_all()
    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.15)
    t_wall0 = time.time()
    e0.record(stream)
    for k in range(K):
        step(evs[k])
    e1.record(stream)
    sync_all()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    total_ms = e0.elapsed_time(e1)
    filt_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    comm_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / K
    geno_ms = sum(e[2].elapsed_time(e[3]) for e in evs) / K
    if world > 1:
        t = torch.tensor([total_ms, filt_ms, geno_ms, comm_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, filt_ms, geno_ms, comm_ms = t.tolist()
        tot = torch.tensor([n_rec, n_bytes, n_hits], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        job_rec = int(tot[0])
    else:
        job_rec = n_rec
    st = filt.read_stats()
    if xchg and xchg.timed_out():
        raise SystemExit(f"rank {rank}: a wait in the fused counter exchange timed out")

    # the dominant kernel on its own: the library records CUDA events around its scan kernel on this stream
    def timed_scan(n_iter):
        capi.check(lib.svjg_filter_profile(1))
        tot, ms = 0.0, C.c_float()
        for _ in range(n_iter):
            capi.check(lib.svjg_filter_reset(filt.counts.data_ptr(), tables.num_sv, filt.stats.data_ptr(), sp))
            capi.check(lib.svjg_filter_device(tables._h, d_gaf.data_ptr(), n_bytes, 0, 100, filt.counts.data_ptr(),
                                              filt.hit_sv2.data_ptr(), filt.hit_off.data_ptr(), filt.hit_len.data_ptr(),
                                              filt.hit_cap, filt.stats.data_ptr(), sp))
            capi.check(lib.svjg_filter_scan_ms(C.byref(ms)))
            tot += ms.value
        capi.check(lib.svjg_filter_profile(0))
        return tot / n_iter
    timed_scan(2)
    scan_ms = timed_scan(max(3, min(K, 10)))
    if args.kernel_only:
        if rank == 0:
            print(json.dumps({"kernel_ms": {"filter": filt_ms, "allreduce": comm_ms, "genotype": geno_ms},
                              "scan_ms": scan_ms, "GBps": n_bytes / filt_ms / 1e6, "stats": st}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the C ABI with host buffers (copies inside the timed region)
    Ke = args.e2e_steps or max(3, min(K, 20))
    hit_cap = n_hits + 1024

    host_out = alnfilter.HostBuffers(tables, hit_cap)     # pinned result arrays, allocated once like h_gaf

    def e2e_step():
        res = alnfilter.filter_host(tables, h_gaf, out=host_out)
        dc = torch.from_numpy(res.counts.view(np.int32)).to(dev, non_blocking=True)
        if world > 1:
            dist.all_reduce(dc)
        out = genotype.genotype_device(dc, d_idx, d_ty)           # the catalogue's index / type arrays stay on the device
        return res, out

    for _ in range(2):
        res, out = e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(Ke):
        res, out = e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    h2d = n_bytes + tables.num_sv * 8
    d2h = tables.num_sv * 8 + 64 + res.n_hits * 16 + n_loc * (24 + 1 + 8 + 1)     # hits: u32 sv, u64 offset, u32 length

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    algo_bytes = n_bytes + 16 * n_hits            # line bytes once + 12 B hit tuple + 4 B counter RMW
    achieved = algo_bytes / (filt_ms * 1e-3) / 1e9
    geno_bytes = 42 * n_loc
    sample_n, tf, tg, n_gt = cpu_baseline(gaf, edges_text, gfa_text, vcf, args.cpu_sample)
    line = {
        "metric": METRIC, "value": job_rec * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(3, args.warmup), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int64", "data": "synthetic",
        "config": {
            "workload": f"{args.workload} x{scale:g} per GPU: {n_rec} GAF records ({n_bytes / 1e6:.1f} MB), "
                        f"{n_sv} VCF SVs, {tables.num_links} link keys, {tables.num_sv} sv keys",
            "records_per_gpu": n_rec, "gaf_bytes_per_gpu": n_bytes, "hits_per_gpu": n_hits, "svs": n_sv,
            "multi_node_records": st["n_multi"], "l2": "input larger than L2 (no flush needed)" if n_bytes > 200e6
            else "input smaller than L2: resident re-reads possible",
            "tables_device_bytes": tables.device_bytes, "gen_seconds": round(gen_s, 1),
        },
        "svs_genotyped_per_sec": n_sv / (geno_ms * 1e-3) if geno_ms > 0 else None,
        "kernel_ms": {"filter": filt_ms, "filter_scan_only": scan_ms, "allreduce": comm_ms, "genotype": geno_ms},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": FILTER_TRAFFIC.get(args.workload), "kernel": "filter chain: probe+scan+exact", "algorithmic_bytes_per_launch": algo_bytes,
                     "peak_source": peak_src,
                     "dominant_kernel": {"name": "scan_kernel", "ms": scan_ms, "share_of_chain": scan_ms / filt_ms,
                                         "achieved": n_bytes / (scan_ms * 1e-3) / 1e9, "frac": n_bytes / (scan_ms * 1e-3) / 1e9 / peak,
                                         "algorithmic_bytes_per_launch": n_bytes},
                     "genotype_kernel": {"achieved": geno_bytes / (geno_ms * 1e-3) / 1e9 if geno_ms > 0 else None,
                                         "bytes_per_launch": geno_bytes}},
        "cpu_baseline": {"value": n_rec / (tf * n_rec / sample_n + tg), "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"first {sample_n} records of the batch (filter + json.dumps {tf:.2f}s) projected to "
                                   f"{n_rec} records, plus genotyping all {n_sv} SVs ({tg:.2f}s); oracle/svjg_oracle.py, "
                                   "single thread like the reference"},
        "cpu_baseline_c": cpu_baseline_c(gaf, edges_text, gfa_text, check=(tables.sv_ids, res.counts, res.n_hits)),
        "e2e": {"value": job_rec * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "ms_per_step": 1000 * e2e_s / Ke},
        "collective": ("p2p-fused: counters summed inside the genotype kernel over NVLink peer memory" if xchg else
                       ("nccl all_reduce" if world > 1 else "none (one GPU)")),
        "gpu_launches": 5 * K,   # reset + probe + scan + exact + genotype
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
