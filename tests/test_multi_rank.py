"""world_size-2 run of the multi-GPU host logic on CPU (gloo): byte-range sharding snapped to
line ends, the single all-reduce of the per-SV counters and the rank-ordered merge of the hit
tuples.  The per-shard filter is the oracle here (the CUDA filter needs a GPU; its own parity
is tests/test_gpu_parity.py) — what is checked is that sharding + exchange reproduce the
single-process result exactly."""
import json
import os
import socket

import numpy as np
import pytest

from conftest import alt_len_from_gfa_text, read_golden
from oracle import svjg_oracle as O
from svjg import shard


def test_shard_cuts_tile_the_buffer_at_line_ends():
    gaf = read_golden("s3.gaf.gz").encode()
    for world in (1, 2, 3, 8, 64):
        cuts = shard.shard_cuts(gaf, world)
        assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == len(gaf)
        assert all(a <= b for a, b in zip(cuts, cuts[1:]))
        for c in cuts[1:-1]:
            assert c == len(gaf) or gaf[c - 1:c] == b"\n"
        assert b"".join(gaf[a:b] for a, b in zip(cuts, cuts[1:])) == gaf
    # ragged inputs: empty, no newline at all, fewer lines than ranks
    assert shard.shard_cuts(b"", 4) == [0, 0, 0, 0, 0]
    assert shard.shard_cuts(b"abc", 2) == [0, 3, 3]
    assert shard.shard_cuts(b"a\nb\n", 8)[-1] == 4


def _worker(rank, world, port, tag, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gaf = read_golden(f"{tag}.gaf.gz").encode()
    edges = json.loads(read_golden(f"{tag}_svs_edges.json.gz"))
    alt = alt_len_from_gfa_text(read_golden(f"{tag}.gfa.gz"))
    ids = sorted({sv for ents in edges.values() for sv, _ in ents})
    index = {s: i for i, s in enumerate(ids)}
    cuts = shard.shard_cuts(gaf, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    counts = np.zeros((len(ids), 2), dtype=np.int32)
    sv2, off, ln = [], [], []
    pos = 0
    for line in gaf[lo:hi].decode().splitlines(True):
        nb = len(line.encode())
        for sv, allele in O.record_hits(line, edges, alt):
            counts[index[sv], allele] += 1
            sv2.append(index[sv] * 2 + allele)
            off.append(pos)
            ln.append(nb)
        pos += nb
    t = torch.from_numpy(counts)
    shard.allreduce_counts(t)
    merged = shard.gather_hits(sv2, off, ln, lo)
    np.save(os.path.join(out_dir, f"counts_{rank}.npy"), counts)
    if rank == 0:
        np.savez(os.path.join(out_dir, "hits.npz"), sv2=merged[0], off=merged[1], ln=merged[2])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_gloo_run_equals_single_process(tmp_path, world):
    import torch.multiprocessing as mp
    tag = "s3"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, tag, str(tmp_path)), nprocs=world, join=True)
    gaf = read_golden(f"{tag}.gaf.gz")
    edges = json.loads(read_golden(f"{tag}_svs_edges.json.gz"))
    alt = alt_len_from_gfa_text(read_golden(f"{tag}.gfa.gz"))
    ids = sorted({sv for ents in edges.values() for sv, _ in ents})
    want = O.hit_counts(O.filter_alignments(gaf.splitlines(True), edges, alt))
    for r in range(world):
        c = np.load(tmp_path / f"counts_{r}.npy")
        got = {ids[i]: (int(c[i, 0]), int(c[i, 1])) for i in range(len(ids)) if c[i].any()}
        assert got == want, f"rank {r}"
    h = np.load(tmp_path / "hits.npz")
    # rank order is file order: offsets ascend, and the tuples are those of a single pass
    assert (np.diff(h["off"].astype(np.int64)) >= 0).all()
    single = []
    pos = 0
    index = {s: i for i, s in enumerate(ids)}
    for line in gaf.splitlines(True):
        for sv, allele in O.record_hits(line, edges, alt):
            single.append((index[sv] * 2 + allele, pos, len(line.encode())))
        pos += len(line.encode())
    assert list(zip(h["sv2"].tolist(), h["off"].tolist(), h["ln"].tolist())) == single


def _xchg_worker(rank, world, port, tag, out_dir):
    """One process per rank, both on cuda:0 (CUDA IPC works between processes on one device): filter the
    rank's shard into the exchange region, then the fused exchange + genotype kernel."""
    import ctypes as C
    import math

    import torch
    import torch.distributed as dist
    from svjg import alnfilter, capi, genotype
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()                  # one device per rank where the box has them
    torch.cuda.set_device(dev)
    gaf = read_golden(f"{tag}.gaf.gz").encode()
    t = alnfilter.Tables.from_memory(read_golden(f"{tag}_svs_edges.json.gz"), read_golden(f"{tag}.gfa.gz")).to_device(dev)
    cuts = shard.shard_cuts(gaf, world)
    d_gaf = torch.frombuffer(bytearray(gaf[cuts[rank]:cuts[rank + 1]]), dtype=torch.uint8).cuda()

    def gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    x = shard.CounterExchange(t.num_sv, rank, world, gather)
    header, recs = genotype.parse_vcf(read_golden(f"{tag}.vcf.gz").splitlines(True))
    idx = np.array([capi.NO_SV if (r[2] is None or t.find_sv(r[2]) is None) else t.find_sv(r[2]) for r in recs], dtype=np.uint32)
    ty = np.array([r[1] for r in recs], dtype=np.uint8)
    n = len(recs)
    d_idx, d_ty = torch.from_numpy(idx.view(np.int32)).cuda(), torch.from_numpy(ty).cuda()
    lut = torch.from_numpy(genotype.log10comb_lut()).cuda()
    d_pl = torch.zeros((n, 3), dtype=torch.int64, device="cuda")
    d_gt = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_ad = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    d_fl = torch.zeros(n, dtype=torch.uint8, device="cuda")
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    la, lb, lh = math.log10(1 - genotype.ERR), math.log10(genotype.ERR), math.log10(1 / 2)
    for step in (1, 2, 3):                                   # both counter buffers, and a reuse
        cp = x.counts_ptr(step)
        capi.check(capi.lib.svjg_filter_reset(cp, t.num_sv, stats.data_ptr(), sp))
        capi.check(capi.lib.svjg_filter_device(t._h, d_gaf.data_ptr(), d_gaf.numel(), 0, 100, cp, None, None, None, 0,
                                               stats.data_ptr(), sp))
        if step == 3:                                        # the launch with its checks (time-out, counts beyond the table)
            x.genotype_checked(step, d_idx, d_ty, n, 3, la, lb, lh, lut, genotype.LUT_NMAX, d_pl, d_gt, d_ad, d_fl, sp,
                               torch.cuda.synchronize)
        else:
            x.genotype(step, d_idx.data_ptr(), d_ty.data_ptr(), n, 3, la, lb, lh, lut.data_ptr(), genotype.LUT_NMAX,
                       d_pl.data_ptr(), d_gt.data_ptr(), d_ad.data_ptr(), d_fl.data_ptr(), sp)
            torch.cuda.synchronize()
        assert not x.timed_out()
        np.savez(os.path.join(out_dir, f"geno_{rank}_{step}.npz"), pl=d_pl.cpu().numpy(), gt=d_gt.cpu().numpy(),
                 ad=d_ad.cpu().numpy(), fl=d_fl.cpu().numpy())
        dist.barrier()
    x.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_fused_counter_exchange_two_ranks(tmp_path):
    """svjg_xchg_* / svjg_genotype_xchg: two ranks (on two devices where the box has them) sum their counters
    inside the genotype kernel over mapped peer memory; every rank must get what one process gets from the
    whole file, and that is what the reference's genotyper printed (tests/golden)."""
    import torch
    import torch.multiprocessing as mp
    from svjg import alnfilter, capi, genotype
    assert torch.cuda.is_available()
    tag, world = "s3", 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_xchg_worker, args=(world, port, tag, str(tmp_path)), nprocs=world, join=True)
    t = alnfilter.Tables.from_memory(read_golden(f"{tag}_svs_edges.json.gz"), read_golden(f"{tag}.gfa.gz")).to_device(0)
    res = alnfilter.filter_host(t, read_golden(f"{tag}.gaf.gz").encode(), want_hits=False)
    header, recs = genotype.parse_vcf(read_golden(f"{tag}.vcf.gz").splitlines(True))
    idx = np.array([capi.NO_SV if (r[2] is None or t.find_sv(r[2]) is None) else t.find_sv(r[2]) for r in recs], dtype=np.uint32)
    ty = np.array([r[1] for r in recs], dtype=np.uint8)
    gt, fl, ad, pl = genotype.genotype_device(torch.from_numpy(res.counts.view(np.int32)).cuda(), idx, ty)
    for r in range(world):
        for step in (1, 2, 3):
            z = np.load(tmp_path / f"geno_{r}_{step}.npz")
            assert (z["pl"] == pl).all() and (z["gt"] == gt).all() and (z["ad"].view(np.uint32) == ad).all() and (z["fl"] == fl).all()
    z = np.load(tmp_path / "geno_1_3.npz")
    assert genotype.format_vcf(header, recs, z["gt"], z["fl"], z["ad"].view(np.uint32), z["pl"])[0] == read_golden(f"{tag}_genotype.vcf.gz")
