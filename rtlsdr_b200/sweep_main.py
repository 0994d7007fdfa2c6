"""Hop-sharded rtl_power sweep over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        -m rtlsdr_b200.sweep_main -f 24M:1766M:1k --sweeps 6 --intervals 2 -o scan.csv

Every rank plans the same scan (host/rtl_power_plan.c), takes a contiguous range of hops
(`shard_hops`), drives its own rtlsdr_gpu_scan handle from the synthetic source (bytes are a pure
function of (hop, sweep), so the result does not depend on N), and once per integration interval
bins + dB rows + sample counts reach rank 0 -- written by every rank's report epilogue straight
into rank 0's buffer over NVLink (symmetric-memory peer mapping, one barrier per interval), or by
ONE NCCL gather where peer mapping is unavailable / RTLSDR_B200_NCCL_GATHER=1 -- and rank 0 prints
the reference's CSV rows in hop order (rtl_power.c:995-1000).  With N = 1 no process group is created.
"""
import argparse
import ctypes
import os
import sys

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-f", dest="range", required=True, help="lower:upper:bin_size (like rtl_power)")
    ap.add_argument("-c", dest="crop", default="0", help="crop percent, e.g. 20%%")
    ap.add_argument("-w", dest="window", default="rectangle")
    ap.add_argument("-F", dest="fir", type=int, default=None)
    ap.add_argument("-P", dest="peak", action="store_true")
    ap.add_argument("--sweeps", type=int, default=8, help="sweeps over all hops per integration interval")
    ap.add_argument("--intervals", type=int, default=1)
    ap.add_argument("--synth", default="xorshift", choices=["xorshift", "counter", "const", "biased", "tone"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--param", type=int, default=0)
    ap.add_argument("--stamp", default="2026-01-01, 00:00:00", help="fixed 'date, time' prefix of the rows")
    ap.add_argument("-o", dest="out", default="-")
    args = ap.parse_args(argv)

    import torch
    import torch.distributed as dist

    from . import scan as rs
    from .planner import host_library, plan_scan
    from .sweep import SpectrumGather, format_rows, shard_hops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("sweep_main: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    host = host_library()
    crop = host.rp_atofp(args.crop.encode())
    plan = plan_scan(args.range, crop, args.fir)
    pd = plan.as_dict()
    pd["peak_hold"] = 1 if args.peak else 0
    tc, b, n = pd["tune_count"], pd["buf_len"], 1 << pd["bin_e"]
    mine = shard_hops(tc, world, rank)
    mode = ["xorshift", "counter", "const", "biased", "tone"].index(args.synth)

    g = None
    if len(mine):
        window = rs.window_coefs(args.window, n) if pd["bin_e"] else None
        g = rs.GpuScan.from_plan(pd, window_coefs=window, device=local, hops=list(mine))
    db_count = g.db_count if g else plan.db_count
    gather = SpectrumGather(tc, n, db_count, world, rank, torch.device("cuda", local))
    if rank == 0 and world > 1:
        print("sweep_main: interval reports " + ("are written by every rank straight into rank 0's buffer (NVLink peer "
              "mapping, one symmetric-memory barrier per interval)" if gather.peer is not None else
              "are gathered with one NCCL gather per interval"), file=sys.stderr)
    stream = torch.cuda.ExternalStream(g.get_stream()) if g else torch.cuda.current_stream()
    pinned = rs.PinnedBuffer(max(1, args.sweeps * len(mine) * b)) if g else None

    rows_out = []
    for interval in range(args.intervals):
        if g:
            cube = pinned.view(np.uint8, (args.sweeps, len(mine), b))
            for s in range(args.sweeps):
                sweep_index = interval * args.sweeps + s
                for k, hop in enumerate(mine):
                    host.synth_generate(mode, ctypes.c_uint64(args.seed), args.param, tc, hop,
                                        ctypes.c_uint64(sweep_index), cube[s, k].ctypes.data, ctypes.c_size_t(b))
            g.submit_batch(0, len(mine), args.sweeps, pinned.ptr, len(mine) * b, b)
            p_avg, p_smp, p_db = gather.pointers()
            g.collect_device(p_avg, p_smp, p_db)
            with torch.cuda.stream(stream):
                gather.samples_are_int32()
        # one collective per interval; the current stream must see the report first
        torch.cuda.current_stream().wait_stream(stream)
        report = gather.gather()
        if rank == 0:
            rows_out += format_rows(plan, report, args.stamp)
    if rank == 0:
        text = "".join(rows_out)
        if args.out == "-":
            sys.stdout.write(text)
        else:
            with open(args.out, "w") as f:
                f.write(text)
    if g:
        g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
