"""world_size-2 gloo test of the hop sharding + per-interval gather (host logic
of rtlsdr_b200/sweep.py).  The per-rank spectra come from the oracle here; on
the GPU box the same code moves what rtlsdr_gpu_scan_collect_device() wrote."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtlsdr_b200.sweep import SpectrumGather, max_hops_per_rank, shard_hops


def test_shard_hops_partitions_exactly():
    for tc in (1, 2, 9, 10, 512, 623, 3000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                seen += list(shard_hops(tc, world, r))
            assert seen == list(range(tc))
            sizes = [len(shard_hops(tc, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
            assert max(sizes) == max_hops_per_rank(tc, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tune_count, bin_e, out_path):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from oracles import PortOracle, SYNTH_BIASED
    from scan_cases import expected, make_reads, plan_dict
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        port_o = PortOracle()
        n = 1 << bin_e
        plan = plan_dict(bin_e, tune_count=tune_count, crop=0.2, rate=2777777)
        w = port_o.window_coefs("hamming", n)
        reads, hops = make_reads(port_o.lib, plan, 2, SYNTH_BIASED, seed=11, param=12)
        mine = shard_hops(tune_count, world, rank)
        sel = np.isin(hops, list(mine))
        # every rank only ever sees the reads of its own hops
        sub = dict(plan)
        avg, smp, db = expected(port_o, sub, w, reads[sel], hops[sel])
        g = SpectrumGather(tune_count, n, db.shape[1], world, rank, "cpu")
        assert g.mode == "host"
        full = expected(port_o, plan, w, reads, hops)
        ok = True
        # three intervals through the two alternating report buffers; interval j scales the bins by j + 1 so that
        # a stale or swapped buffer cannot go unnoticed
        for j in range(3):
            k = j & 1
            a, d, s = g.views(k)
            a.copy_(torch.from_numpy(avg[mine.start: mine.stop] * (j + 1)))
            d.copy_(torch.from_numpy(db[mine.start: mine.stop] + j))
            s.copy_(torch.from_numpy(smp[mine.start: mine.stop].astype(np.int32) + j))
            g.before_collect(k)
            g.publish(k, to_host=True)
            rep = g.fetch(k)
            if rank == 0:
                ok = ok and (np.array_equal(rep.avg, full[0] * (j + 1)) and np.array_equal(rep.samples, full[1] + j)
                             and np.array_equal(rep.db, full[2] + j, equal_nan=True))
            else:
                assert rep is None
        if rank == 0:
            with open(out_path, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tune_count", [1, 2, 5])
def test_gather_world_size_2(tmp_path, tune_count):
    out = tmp_path / "result.txt"
    port = _free_port()
    mp.spawn(_worker, args=(2, port, tune_count, 8, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def _worker_reads(rank, world, port, tune_count, bin_e, peak, out_path):
    """read-sharded scan (single-hop configs): every rank reports ALL hops for ITS share of the sweeps; rank 0 folds
    the gathered partial accumulator sets like rtlsdr_gpu_scan_merge_device does (sum, or maximum under peak hold)"""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from oracles import PortOracle, SYNTH_BIASED
    from scan_cases import expected, make_reads, plan_dict
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        port_o = PortOracle()
        n, sweeps = 1 << bin_e, 5
        plan = plan_dict(bin_e, tune_count=tune_count, peak_hold=peak)
        w = port_o.window_coefs("hamming", n)
        g = SpectrumGather(tune_count, n, n + 1, world, rank, "cpu", replicated=True)
        assert g.mode == "host" and list(g.my_hops) == list(range(tune_count)) and g.hmax == tune_count
        ok = True
        for j in range(3):
            reads, hops = make_reads(port_o.lib, plan, sweeps, SYNTH_BIASED, seed=20 + j, param=9)
            full = expected(port_o, plan, w, reads, hops)
            mine = shard_hops(sweeps, world, rank)          # contiguous share of the sweeps
            lo, hi = mine.start * tune_count, mine.stop * tune_count
            avg, smp, _ = expected(port_o, plan, w, reads[lo:hi], hops[lo:hi])
            k = j & 1
            a, _, s = g.views(k)
            a.copy_(torch.from_numpy(avg))
            s.copy_(torch.from_numpy(smp.astype(np.int32)))
            g.before_collect(k)
            g.publish(k)
            if rank == 0:
                bufs = g.recv[k].numpy()
                sets = bufs[:, : tune_count * n].reshape(world, tune_count, n)
                o = tune_count * (n + n + 1)
                counts = bufs[:, o: o + g.smp_words].copy().view(np.int32).reshape(world, -1)[:, :tune_count]
                merged = sets.max(axis=0) if peak else sets.sum(axis=0)
                ok = ok and np.array_equal(merged, full[0]) and np.array_equal(counts.sum(axis=0), full[1])
        if rank == 0:
            with open(out_path, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tune_count,peak", [(1, 0), (1, 1), (3, 0)])
def test_read_sharded_partial_sets_world_size_2(tmp_path, tune_count, peak):
    out = tmp_path / "result.txt"
    port = _free_port()
    mp.spawn(_worker_reads, args=(2, port, tune_count, 7, peak, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
