"""Hop-sharded sweeps over the GPUs of one box.

Frequency hops are independent (reference src/rtl_power.c:650-719: tunes[i] are
disjoint), so hops are dealt to ranks in contiguous ranges and every rank runs
its own rtlsdr_gpu_scan handle; there is no collective on the data path.  Once
per integration interval the per-hop spectra (int64 bins, dB doubles, sample
counts) are gathered to rank 0 with ONE gather, and rank 0 prints the rows in
hop order like the reference's report loop (rtl_power.c:995-1000).

One process per GPU; torch.distributed is only the plumbing (NCCL over NVLink
on the GPU box, gloo in the CPU tests).

On a box with NVLink peer access the gather needs no collective at all: the report
epilogue (rtlsdr_gpu_scan_collect_device) writes wherever it is pointed, so every
rank points it at its slot of rank 0's buffer (torch symmetric memory peer mapping)
and one symmetric-memory barrier per interval tells rank 0 that the slots are
complete (RTLSDR_B200_NCCL_GATHER=1 keeps the NCCL gather).
"""
import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_hops(tune_count: int, world: int, rank: int) -> range:
    """Contiguous, balanced hop range of `rank`: sizes differ by at most one."""
    base, extra = divmod(tune_count, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def max_hops_per_rank(tune_count: int, world: int) -> int:
    return -(-tune_count // world)


@dataclass
class IntervalReport:
    avg: np.ndarray       # [tune_count, N] int64, natural FFT order
    db: np.ndarray        # [tune_count, db_count] float64
    samples: np.ndarray   # [tune_count] int32


class SpectrumGather:
    """Packs one rank's interval result into a single int64 buffer
    [avg | db (bit pattern) | samples] padded to the largest shard, and gathers all
    ranks' buffers to rank 0 with one collective call per interval."""

    def __init__(self, tune_count, n_bins, db_count, world, rank, device):
        self.tune_count, self.n, self.db_count = tune_count, n_bins, db_count
        self.world, self.rank = world, rank
        self.hmax = max_hops_per_rank(tune_count, world)
        self.words = self.hmax * (n_bins + db_count + 1)
        self.my_hops = shard_hops(tune_count, world, rank)
        self.peer = self._try_peer(torch.device(device)) if world > 1 else None
        if self.peer is not None:
            # this rank's slot of rank 0's buffer, mapped into this process over NVLink
            self.send = self.peer[2][rank * self.words:(rank + 1) * self.words]
            self.recv = None
        else:
            self.send = torch.zeros(self.words, dtype=torch.int64, device=device)
            self.recv = ([torch.zeros(self.words, dtype=torch.int64, device=device) for _ in range(world)]
                         if rank == 0 and world > 1 else None)

    def _try_peer(self, device):
        """(local buffer, symmetric-memory handle, rank 0's buffer as seen from here), or None; all ranks agree"""
        ok, peer = 0, None
        if device.type == "cuda" and not os.environ.get("RTLSDR_B200_NCCL_GATHER"):
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(self.world * self.words, dtype=torch.int64, device=device)
                buf.zero_()
                hdl = symm.rendezvous(buf, dist.group.WORLD)
                root = hdl.get_buffer(0, (self.world * self.words,), torch.int64)
                peer, ok = (buf, hdl, root), 1
            except Exception:  # noqa: BLE001 -- any failure means: use the collective
                peer = None
        if device.type != "cuda":
            return None  # (gloo tests: every rank takes this branch, no agreement round needed)
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return peer if int(flag.item()) == 1 else None

    # views into the send buffer, sized for THIS rank's hop count
    def views(self):
        h = len(self.my_hops)
        a = self.send[: self.hmax * self.n].view(self.hmax, self.n)[:h]
        o = self.hmax * self.n
        d = self.send[o: o + self.hmax * self.db_count].view(torch.float64).view(self.hmax, self.db_count)[:h]
        o += self.hmax * self.db_count
        s = self.send[o: o + self.hmax][:h]
        return a, d, s

    def pointers(self):
        """device addresses for rtlsdr_gpu_scan_collect_device(avg, samples(int32), db)"""
        base = self.send.data_ptr()
        p_db = base + self.hmax * self.n * 8
        p_smp = p_db + self.hmax * self.db_count * 8
        return base, p_smp, p_db

    def samples_are_int32(self):
        """collect_device writes int32 sample counts; widen them in place to int64 words."""
        h = len(self.my_hops)
        o = self.hmax * (self.n + self.db_count)
        raw = self.send[o: o + self.hmax].view(torch.int32)
        vals = raw[:h].clone().to(torch.int64)
        self.send[o: o + self.hmax].zero_()
        self.send[o: o + h] = vals

    def gather(self) -> Optional[IntervalReport]:
        if self.peer is not None:
            buf, hdl, _ = self.peer
            hdl.barrier(channel=0)          # every rank's epilogue has stored into rank 0's buffer
            if self.rank != 0:
                hdl.barrier(channel=1)      # rank 0 has read it: the slot may be rewritten
                return None
            host = buf.cpu()                # synchronises with the barrier above
            hdl.barrier(channel=1)
            bufs = [host[r * self.words:(r + 1) * self.words] for r in range(self.world)]
        elif self.world > 1:
            dist.gather(self.send, self.recv, dst=0)
            if self.rank != 0:
                return None
            bufs = self.recv
        else:
            bufs = [self.send]
        avg = np.zeros((self.tune_count, self.n), dtype=np.int64)
        db = np.zeros((self.tune_count, self.db_count), dtype=np.float64)
        smp = np.zeros(self.tune_count, dtype=np.int32)
        for r, buf in enumerate(bufs):
            hops = shard_hops(self.tune_count, self.world, r)
            if len(hops) == 0:
                continue
            b = buf.cpu().numpy()
            h = len(hops)
            avg[hops.start: hops.stop] = b[: self.hmax * self.n].reshape(self.hmax, self.n)[:h]
            o = self.hmax * self.n
            db[hops.start: hops.stop] = b[o: o + self.hmax * self.db_count].view(np.float64).reshape(
                self.hmax, self.db_count)[:h]
            o += self.hmax * self.db_count
            smp[hops.start: hops.stop] = b[o: o + h].astype(np.int32)
        return IntervalReport(avg, db, smp)


def format_rows(plan, report: IntervalReport, stamp: str) -> List[str]:
    """rank 0: 'date, time, low, high, step, samples, dB...' rows in hop order"""
    return [f"{stamp}, " + plan.csv_row(h, int(report.samples[h]), report.db[h])
            for h in range(plan.tune_count)]
