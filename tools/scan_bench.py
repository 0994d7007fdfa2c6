#!/usr/bin/env python
"""Device-resident throughput of any rtl_power configuration (not the driver's bench.py contract:
a helper for the BASELINE.json configs other than configs[1] and for the decimating companions).

    python tools/scan_bench.py --range 24M:1457.6M:700 --passes 64
    python tools/scan_bench.py --range 100M:100.1M:100 --passes 512            # boxcar ds=28
    python tools/scan_bench.py --range 100M:102.4M:19 -w blackman-harris -P --passes 64   # 2^17 bins
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--range", default=None)
    ap.add_argument("--boxcar", default=None, metavar="BIN_E:DS",
                    help="narrow boxcar scan given directly (one hop, buf_len = 2 * 2^BIN_E * DS) instead of --range")
    ap.add_argument("--hops", type=int, default=1, help="with --boxcar: hop count")
    ap.add_argument("-c", "--crop", type=float, default=0.0)
    ap.add_argument("-w", "--window", default="rectangle")
    ap.add_argument("-F", "--fir", type=int, default=None)
    ap.add_argument("-P", "--peak", action="store_true")
    ap.add_argument("--passes", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sets", type=int, default=0, help="distinct input sets (0 = enough to exceed 2x L2)")
    ap.add_argument("--no-kernel-time", action="store_true", help="do not bracket the transform kernels with events")
    args = ap.parse_args()

    import torch
    import rtlsdr_b200.scan as rs
    from rtlsdr_b200.planner import plan_scan

    if args.boxcar:
        be, ds = (int(v) for v in args.boxcar.split(":"))
        pd = dict(tune_count=args.hops, bin_e=be, buf_len=2 * (1 << be) * ds, downsample=ds, downsample_passes=0,
                  boxcar=1, comp_fir_size=0, rate=2800000, crop=args.crop)
        args.range = "boxcar " + args.boxcar
    else:
        pd = plan_scan(args.range, args.crop, args.fir).as_dict()
    pd["peak_hold"] = 1 if args.peak else 0
    tc, b, n = pd["tune_count"], pd["buf_len"], 1 << pd["bin_e"]
    g = rs.GpuScan.from_plan(pd, window_coefs=rs.window_coefs(args.window, n) if pd["bin_e"] else None)
    stream = torch.cuda.ExternalStream(g.get_stream())
    step_bytes = args.passes * tc * b
    n_sets = args.sets or max(1, -(-(256 << 20) // step_bytes))
    dev_in = torch.randint(0, 256, (n_sets, args.passes, tc, b), dtype=torch.uint8, device="cuda")
    out = torch.zeros(tc * (n + g.db_count + 1), dtype=torch.int64, device="cuda")
    p_avg = out.data_ptr()
    p_db = p_avg + tc * n * 8
    p_smp = p_db + tc * g.db_count * 8

    def step(i):
        g.submit_device(0, tc, args.passes, dev_in[i % n_sets].data_ptr(), tc * b, b)
        g.collect_device(p_avg, p_smp, p_db)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if not args.no_kernel_time:
        g.kernel_time()
    s0 = g.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    k_ms, k_n = (0.0, 0) if args.no_kernel_time else g.kernel_time()
    s1 = g.stats()
    samples = step_bytes // 2
    print(json.dumps({
        "range": args.range, "window": args.window, "fir": args.fir, "peak": args.peak,
        "plan": {k: pd[k] for k in ("tune_count", "bin_e", "buf_len", "downsample", "downsample_passes", "rate")},
        "passes": args.passes, "bytes_per_step": step_bytes, "input_sets": n_sets, "ms_per_step": ms,
        "transform_ms_per_step": k_ms / max(args.steps, 1), "timed_scopes": k_n,
        "Msamples_per_s": samples / (ms * 1e-3) / 1e6, "GBps_algorithmic": step_bytes / (ms * 1e-3) / 1e9,
        "frac_of_6542.7": step_bytes / (ms * 1e-3) / 1e9 / 6542.7,
        "launches_per_step": (s1["kernel_launches"] - s0["kernel_launches"]) / args.steps}))
    g.close()


if __name__ == "__main__":
    main()
