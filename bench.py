#!/usr/bin/env python
"""bench.py -- throughput of rtl_power's per-hop scan pipeline on B200.

Metric (BASELINE.json): input Msamples/s (1 sample = 1 complex IQ pair = 2 input
bytes), whole job over all GPUs, plus the fraction of the HBM roofline.

Workload = BASELINE.json configs[1]: FM band scan 88-108 MHz, 4096 bins,
hamming, -c 20%, one 10 s integration interval = 9 hops x 377 sweeps of 16384
bytes.  One "step" = one whole integration interval: every read of
the interval through the transform, then the report epilogue (DC nuke, half
swap, crop, dB) and, with N > 1 GPUs, ONE NCCL gather of the spectra to rank 0.
With N GPUs every rank owns its own 9 hops (9 N hop streams in total, weak
scaling: hops are independent, there is no data-path collective).

  value : inputs already resident in HBM (device-resident replay), CUDA-event
          timed on the launching stream, max over ranks.
  e2e   : the same interval through the public C ABI with HOST buffers:
          rtlsdr_gpu_scan_submit_batch() from pinned memory (H2D inside the timed
          region) and rtlsdr_gpu_scan_collect_all() (D2H of bins + dB).
  --impl reference : the reference's own CPU code (oracle/_ref, the unmodified
          rtl_power.c object) on all host cores, bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RANGE, CROP, WINDOW = "88M:108M:1k", 0.2, "hamming"
PASSES = 377                 # sweeps in one 10 s interval at 2 777 777 S/s
WORKLOAD = "fm_band_scan_88-108MHz_4096bins_hamming_crop20_10s"


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get("scan_small_kernel_dram_bytes_per_launch")
    except Exception:
        return None


def issue_roofline(kernel_ms, sm_mhz, n_sms):
    """The bound this kernel actually runs against: warp instructions per clock per SM.  The instruction count
    per launch comes from the committed ncu capture of the same workload (it does not depend on the input
    bytes), the peak from tools/ubench.cu's measured dual-issue rate of a perfectly mixed IMAD + ALU stream."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        inst = d["scan_small_kernel_warp_instructions_per_launch"]
        peak = d["issue_peak_warp_instr_per_clk_per_sm"]
        ipc = inst / (kernel_ms * 1e-3 * sm_mhz * 1e6 * n_sms)
        return {"bound": "issue", "achieved": ipc, "peak": peak, "unit": "warp-instr/clk/SM", "frac": ipc / peak,
                "warp_instructions_per_launch": inst, "sm_mhz": sm_mhz, "sms": n_sms, "peak_source": d["issue_peak_source"]}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- reference arm

_REF = {}


def _ref_init(seed_base):
    """pool initializer: one configured reference instance per worker process"""
    import multiprocessing as mp
    from oracles import RefOracle, PortOracle, SYNTH_XORSHIFT  # noqa
    ident = mp.current_process()._identity
    seed = seed_base + (ident[0] if ident else 0)
    try:
        r = RefOracle()
        r.configure(RANGE, CROP, WINDOW)
        r.source(SYNTH_XORSHIFT, seed, 0)
        r.scan(2)
        _REF.update(kind="reference", ref=r, per_pass=r.plan["tune_count"] * (r.plan["buf_len"] // 2))
    except Exception:
        # compiled reference missing: time the C restatement instead
        import numpy as np
        from rtlsdr_b200.planner import plan_scan
        p = PortOracle()
        plan = plan_scan(RANGE, CROP).as_dict()
        plan["peak_hold"] = 0
        w = p.window_coefs(WINDOW, 1 << plan["bin_e"])
        rng = np.random.default_rng(seed)
        reads = rng.integers(0, 256, (plan["tune_count"] * 4, plan["buf_len"]), dtype=np.uint8)
        hops = [i % plan["tune_count"] for i in range(len(reads))]
        _REF.update(kind="port", port=p, plan=plan, w=w, reads=reads, hops=hops,
                    per_pass=plan["tune_count"] * (plan["buf_len"] // 2))


def _ref_step(passes):
    """(seconds, samples, kind) for `passes` sweeps of the 9-hop plan in this worker"""
    if _REF["kind"] == "reference":
        return _REF["ref"].scan_timed(passes), passes * _REF["per_pass"], "reference"
    t0 = time.perf_counter()
    n = 0
    while n < passes:
        _REF["port"].scan(_REF["plan"], _REF["w"], _REF["reads"], _REF["hops"], _REF["plan"]["tune_count"])
        n += 4
    return time.perf_counter() - t0, n * _REF["per_pass"], "port"


class CpuReference:
    """`procs` persistent worker processes, each an independent instance of the reference's
    single-threaded scanner() (the reference has no threads: rtl_power.c:29-36, 844-846)."""

    def __init__(self, procs):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_ref_init, initargs=(17,))
        self.pool.map(_ref_step, [1] * procs)  # all workers up and configured

    def step(self, passes):
        res = self.pool.map(_ref_step, [passes] * self.procs, chunksize=1)
        return max(r[0] for r in res), sum(r[1] for r in res), res[0][2]

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    ref = CpuReference(cores)
    # bounded sample per step: the whole run stays within ~2 minutes whatever K is
    t_probe, s_probe, kind = ref.step(8)
    per_pass_s = max(t_probe / 8, 1e-5)
    budget_s = 100.0
    passes = int(max(2, min(3000, budget_s / (args.steps + args.warmup) / per_pass_s)))
    for _ in range(args.warmup):
        ref.step(passes)
    total_t, total_s = 0.0, 0
    for _ in range(args.steps):
        t, smp, kind = ref.step(passes)
        total_t += t
        total_s += smp
    ref.close()
    value = total_s / total_t / 1e6
    sample = f"{cores} independent processes x {passes} sweeps x 9 hops x 8192 samples per step"
    line = {
        "metric": "input Msamples/s", "value": value, "unit": "Msamples/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16/int64 fixed point", "data": "synthetic",
        "config": {"workload": WORKLOAD, "hops": 9, "bins": 4096, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- GPU arm

def companion(rs, plan_scan, torch, name, freq, crop, window, fir, peak, passes, steps, peak_gbs):
    """Device-resident throughput of another rtl_power configuration (same method as `value`):
    the HBM-bound regimes of the path that BASELINE.json's bench workload (ds = 1) never enters."""
    plan = plan_scan(freq, crop, fir)
    pd = plan.as_dict()
    pd["peak_hold"] = peak
    tc, b, n = pd["tune_count"], pd["buf_len"], 1 << pd["bin_e"]
    g = rs.GpuScan.from_plan(pd, window_coefs=rs.window_coefs(window, n) if pd["bin_e"] else None)
    stream = torch.cuda.ExternalStream(g.get_stream())
    step_bytes = passes * tc * b
    n_sets = max(1, -(-(300 << 20) // step_bytes))
    dev_in = torch.randint(0, 256, (n_sets, passes, tc, b), dtype=torch.uint8, device="cuda")
    out = torch.zeros(tc * (n + g.db_count + 1), dtype=torch.int64, device="cuda")
    p_avg = out.data_ptr()
    p_db = p_avg + tc * n * 8
    p_smp = p_db + tc * g.db_count * 8

    def step(i):
        g.submit_device(0, tc, passes, dev_in[i % n_sets].data_ptr(), tc * b, b)
        g.collect_device(p_avg, p_smp, p_db)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(3 + i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    g.close()
    del dev_in
    gbs = step_bytes / (ms * 1e-3) / 1e9
    return {"workload": name, "cli": f"-f {freq}" + (f" -F {fir}" if fir is not None else "") + (" -P" if peak else ""),
            "plan": {k: pd[k] for k in ("tune_count", "bin_e", "buf_len", "downsample", "downsample_passes")},
            "value": step_bytes / 2 / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms,
            "bytes_per_step": step_bytes, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peak_gbs}


# diagnostic only (never set by the driver): time the multi-GPU step without its per-interval gather
NO_GATHER = bool(os.environ.get("BENCH_DIAG_NO_GATHER"))


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import rtlsdr_b200.scan as rs
    from rtlsdr_b200.planner import plan_scan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rs.load_library()

    plan = plan_scan(RANGE, CROP)
    pd = plan.as_dict()
    tc, b, n = pd["tune_count"], pd["buf_len"], 1 << pd["bin_e"]
    window = rs.window_coefs(WINDOW, n)
    g = rs.GpuScan.from_plan(pd, window_coefs=window, device=local)
    # time on the handle's own stream (the launching stream), wrapped for torch events / NCCL ordering
    stream = torch.cuda.ExternalStream(g.get_stream())
    db_count = g.db_count

    step_bytes = PASSES * tc * b
    samples_per_step = step_bytes // 2
    # inputs larger than L2 (126 MB): rotate through distinct interval-sized sets
    n_sets = max(2, -(-(300 << 20) // step_bytes))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    dev_in = torch.randint(0, 256, (n_sets, PASSES, tc, b), dtype=torch.uint8, device="cuda", generator=gen)
    out_words = tc * n + tc * db_count + tc
    # two report buffers: interval k's spectra are gathered (comm stream) while interval k+1 is transformed
    sends = [torch.zeros(out_words, dtype=torch.int64, device="cuda") for _ in range(2)]
    gathers = [[torch.zeros(out_words, dtype=torch.int64, device="cuda") for _ in range(world)]
               if (world > 1 and rank == 0) else None for _ in range(2)]
    comm = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    gathered = [torch.cuda.Event() for _ in range(2)]

    # Per-interval exchange, preferred form: no copy and no collective kernel at all.  Every rank's report
    # epilogue stores its bins / dB / sample counts straight into rank 0's buffer through an NVLink peer mapping
    # (torch symmetric memory), and one symmetric-memory barrier per interval (a one-CTA signalling kernel on a
    # second stream) tells rank 0 that every slot is complete.  An NCCL send/receive kernel needs SM resources
    # that two resident transform CTAs per SM do not leave: measured 4 % (2 GPUs) to 10 % (8 GPUs) of the step.
    # BENCH_NCCL_GATHER=1, or a box without peer access, selects the grouped NCCL gather instead.
    peer = None
    if world > 1 and not NO_GATHER:
        ok = 0
        if not os.environ.get("BENCH_NCCL_GATHER"):
            try:
                import torch.distributed._symmetric_memory as symm
                pbuf = symm.empty(2 * world * out_words, dtype=torch.int64, device=torch.device("cuda", local))
                pbuf.zero_()
                hdl = symm.rendezvous(pbuf, dist.group.WORLD)
                peer = (pbuf, hdl, int(hdl.buffer_ptrs[0]))
                ok = 1
            except Exception as exc:  # noqa: BLE001 -- any failure means: use NCCL
                print(f"bench.py: symmetric memory unavailable ({exc!r}); using the NCCL gather", file=sys.stderr)
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            peer = None

    def gather_interval(k):
        """report buffer k is complete once everything enqueued so far on `stream` has run"""
        ready[k].record(stream)
        with torch.cuda.stream(comm):
            comm.wait_event(ready[k])
            if peer is not None:
                peer[1].barrier(channel=k)
            else:
                dist.gather(sends[k], gathers[k], dst=0)
            gathered[k].record(comm)

    def step_device(i):
        # Stream order: scan(i) | [event + gather of interval i-1] | wait(buffer free) | epilogue(i).
        # Nothing sits between epilogue(i-1) and scan(i), so the transform can be launched
        # programmatically dependent on the previous report (it only waits before its flush).
        k = i & 1
        if peer is not None:
            p_avg = peer[2] + (k * world + rank) * out_words * 8   # this rank's slot in rank 0's memory
        else:
            p_avg = sends[k].data_ptr()
        p_db = p_avg + tc * n * 8
        p_smp = p_db + tc * db_count * 8
        g.submit_device(0, tc, PASSES, dev_in[i % n_sets].data_ptr(), tc * b, b)
        if world > 1 and not NO_GATHER:
            if step_device.pending is not None:
                gather_interval(step_device.pending)
            stream.wait_event(gathered[k])          # buffer k's previous gather has finished
            step_device.pending = k
        g.collect_device(p_avg, p_smp, p_db)

    step_device.pending = None

    def drain_gathers():
        if world > 1 and step_device.pending is not None:
            gather_interval(step_device.pending)
            step_device.pending = None
        stream.wait_stream(comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ----
    for i in range(args.warmup):
        step_device(i)
    drain_gathers()
    barrier()
    g.kernel_time()  # arm / reset the per-kernel timers
    g.set_timing(16)  # every 16th transform is bracketed with events (the others can launch dependently)
    s0 = g.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
    drain_gathers()                                 # the last gathers are inside the timed region
    with torch.cuda.stream(stream):
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if peer is not None and rank == 0:
        # every rank's last two reports must have landed in rank 0's buffer: sample counts = 2 * PASSES per hop
        tail = peer[0].view(2, world, out_words)[:, :, tc * n + tc * db_count:].contiguous()
        got = tail.view(torch.int32).view(2, world, 2 * tc)[:, :, :tc]   # int32 sample counts of the tc hops
        assert bool((got == 2 * PASSES).all()), "peer-written reports are incomplete on rank 0"
    s1 = g.stats()
    k_ms, k_n = g.kernel_time()
    g.set_timing(0)
    if rank == 0 and world == 1:
        # a short timed region (small --steps) may end before nvidia-smi's first 50 ms tick: keep the
        # same load running, untimed, until a few clock samples exist
        t_end = time.perf_counter() + 1.0
        i = args.warmup + args.steps
        while len(sampler.rows) < 4 and time.perf_counter() < t_end:
            for _ in range(50):
                step_device(i)
                i += 1
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- end to end through the host-buffer ABI ("e2e") ----
    # A stream of integration intervals: every interval's bytes start in pinned host memory and
    # its report (int64 bins, sample counts, dB rows) ends in host memory.  Two handles alternate,
    # so interval k+1 crosses PCIe while interval k is transformed and collected -- what a
    # continuously running rtl_power does.  Each step pays its own H2D and D2H.
    handles = [g, rs.GpuScan.from_plan(pd, window_coefs=window, device=local)]
    hosts, outs = [], []
    for k in range(2):
        hb = rs.PinnedBuffer(step_bytes)
        hb.array[:] = np.frombuffer(dev_in[k % n_sets].cpu().numpy().tobytes(), dtype=np.uint8)
        ob = rs.PinnedBuffer(tc * n * 8 + tc * db_count * 8 + tc * 4)
        hosts.append(hb)
        outs.append((ob, (ob.view(np.int64, (tc, n)), ob.view(np.int32, (tc,), tc * n * 8 + tc * db_count * 8),
                          ob.view(np.float64, (tc, db_count), tc * n * 8))))
    e2e_steps = max(4, min(args.steps, 400))

    def run_host(steps):
        handles[0].submit_batch(0, tc, PASSES, hosts[0].ptr, tc * b, b)
        res = None
        for i in range(steps):
            if i + 1 < steps:
                handles[(i + 1) & 1].submit_batch(0, tc, PASSES, hosts[(i + 1) & 1].ptr, tc * b, b)
            res = handles[i & 1].collect_all(out=outs[i & 1][1])
        return res

    run_host(max(args.warmup, 4))
    barrier()
    t0 = time.perf_counter()
    res = run_host(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert int(res[1][0]) == 2 * PASSES, "e2e report does not cover one whole interval"
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    handles[1].close()
    d2h = tc * n * 8 + tc * db_count * 8

    if rank == 0:
        peak, peak_src = read_peaks()
        value = world * samples_per_step * args.steps / (ms_max * 1e-3) / 1e6
        e2e = world * samples_per_step * e2e_steps / e2e_s / 1e6
        k_avg_ms = k_ms / max(k_n, 1)
        achieved = 2.0 * samples_per_step / (k_avg_ms * 1e-3) / 1e9 if k_n else None
        line = {
            "metric": "input Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16/int64 fixed point", "data": "synthetic",
            "config": {"workload": WORKLOAD if world == 1 else f"{WORKLOAD} x {world} (9 hops per GPU)",
                       "hops_per_gpu": tc, "bins": n, "reads_per_step_per_gpu": PASSES * tc,
                       "bytes_per_step_per_gpu": step_bytes,
                       "cache": f"inputs larger than L2: {n_sets} distinct interval sets ({n_sets * step_bytes >> 20} MiB) rotated",
                       "gather": ("none: every rank's report epilogue stores its int64 bins + dB straight into rank 0's "
                                  "buffer over NVLink (symmetric-memory peer mapping); one symmetric-memory barrier per "
                                  "step on a second stream") if (world > 1 and peer is not None) else
                                 ("one NCCL gather of int64 bins + dB per step, on a second stream, overlapped with "
                                  "the next interval's transform") if world > 1 else "none"},
            "per_gpu_value": value / world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": read_traffic(),
                         "kernel": "scan_small_kernel<12>", "kernel_ms": k_avg_ms, "kernel_launches_timed": k_n, "kernel_timing": "CUDA events around every 16th launch of the timed region",
                         "algorithmic_bytes_per_launch": 2 * samples_per_step, "peak_source": peak_src,
                         "note": "integer-issue bound, not HBM bound: see DESIGN.md"},
            "e2e": {"value": e2e, "unit": "Msamples/s", "h2d_bytes_per_step": step_bytes,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "rtlsdr_gpu_scan_submit_batch (pinned host input) + rtlsdr_gpu_scan_collect_all "
                           "(pinned host output), two handles alternating so copies overlap the transform"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if k_n and clocks and clocks.get("sm_mhz"):
            line["roofline"]["issue"] = issue_roofline(
                k_avg_ms, clocks["sm_mhz"], torch.cuda.get_device_properties(0).multi_processor_count)
        if world == 1 and not args.no_companions:
            try:
                line["companions"] = [
                    companion(rs, plan_scan, torch, "narrow_scan_boxcar_ds28_1024bins", "100M:100.1M:100", 0.0,
                              "rectangle", None, 0, 8192, 10, peak),
                    companion(rs, plan_scan, torch, "rms_1MHz_bins_10hops", "100M:110M:1M", 0.0,
                              "rectangle", None, 0, 4096, 10, peak),
                    companion(rs, plan_scan, torch, "narrow_scan_fifth_order_x4_fir9_1024bins", "100M:100.1M:100", 0.0,
                              "blackman", 9, 0, 8192, 10, peak),
                    companion(rs, plan_scan, torch, "large_fft_2^17_blackman-harris_peak_hold", "100M:102.4M:19", 0.0,
                              "blackman-harris", None, 1, 256, 10, peak),
                ]
            except Exception as exc:  # companions are extra evidence, never fail the bench line
                line["companions"] = [{"error": repr(exc)}]
        if world == 1 and not args.no_cpu:
            ref = CpuReference(1)
            t, smp, kind = ref.step(4000)
            ref.close()
            line["cpu_baseline"] = {"value": smp / t / 1e6, "unit": "Msamples/s", "cores": 1, "kind": kind,
                                    "sample": "4000 sweeps x 9 hops x 8192 samples (295 M samples), 1 thread"}
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-companions", action="store_true", help="skip the other-configuration measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
