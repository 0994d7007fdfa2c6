"""Round-2 GPU parity tests (through the C ABI, `-m gpu`):

* the CUDA bins DIRECTLY against the compiled, unmodified reference (oracle/_ref) for every BASELINE config
  -- one hop in the chain instead of two (the reference .so travels to the GPU box),
* the equal-run schedule of scan_small_kernel (runs that cut reads in half and cross hop boundaries),
* the new entry points: submit_reads (any hop order), -s iir smoothing, the first release's cfg size,
* the hop-sharded drivers at more than one rank / worker on ONE GPU (gloo ranks, two C workers on device 0)
  with several intervals: rows byte-identical to the 1-rank / 1-worker run.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from oracles import SYNTH_BIASED, SYNTH_XORSHIFT, WINDOWS, RefOracle, fnv1a_int64
from scan_cases import KAT_ROWS, db_close, expected, make_reads, plan_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def scan_mod():
    import rtlsdr_b200.scan as s
    s.load_library()
    return s


def gpu_scan_device(scan_mod, plan, window, reads, **kw):
    """reads: uint8 [passes * tune_count, buf_len] in sweep order, device-resident submission"""
    import torch
    tc, b = plan["tune_count"], plan["buf_len"]
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=window, **kw)
    try:
        dev = torch.from_numpy(np.ascontiguousarray(reads)).cuda()
        torch.cuda.synchronize()
        g.submit_device(0, tc, len(reads) // tc, dev.data_ptr(), tc * b, b)
        return g.collect_all()
    finally:
        g.close()


# BASELINE.json configs 1..5: range, crop, window, -F, -P, sweeps, synthetic mode / param
BASELINE_CASES = [
    ("cfg1", "100M:102.4M:2400", 0.0, "rectangle", -1, 0, 12, SYNTH_XORSHIFT, 0),
    ("cfg1-dc", "100M:102.4M:2400", 0.0, "rectangle", -1, 0, 5, SYNTH_BIASED, 23),
    ("cfg2", "88M:108M:1k", 0.2, "hamming", -1, 0, 5, SYNTH_BIASED, -19),
    ("cfg3", "24M:1766M:1k", 0.0, "rectangle", 9, 0, 2, SYNTH_XORSHIFT, 0),
    ("cfg4", "100M:102.4M:19", 0.0, "blackman-harris", -1, 1, 3, SYNTH_BIASED, 31),
    ("cfg5", "24M:1457.6M:700", 0.0, "rectangle", -1, 0, 3, SYNTH_XORSHIFT, 0),
    ("narrow-boxcar", "100M:100.1M:100", 0.0, "rectangle", -1, 0, 3, SYNTH_BIASED, 40),
    ("narrow-F9", "100M:100.1M:100", 0.0, "blackman", 9, 0, 3, SYNTH_BIASED, 40),
    ("rms", "100M:110M:1M", 0.0, "rectangle", -1, 1, 4, SYNTH_BIASED, -33),
]


@pytest.mark.parametrize("case", BASELINE_CASES, ids=[c[0] for c in BASELINE_CASES])
def test_gpu_against_compiled_reference(scan_mod, case):
    """CUDA path == the unmodified rtl_power.c object (oracle/_ref/librtlpower_ref.so) on the same bytes:
    int64 bins and sample counts bit-exact, and the reference's own CSV text for the first and last hop."""
    if not RefOracle.available():
        pytest.skip("compiled reference (oracle/_ref) not present on this box")
    from rtlsdr_b200.planner import plan_scan
    name, freq, crop, window, fir, peak, sweeps, mode, param = case
    ref = RefOracle()
    plan = ref.configure(freq, crop, window, fir, peak)
    ref.source(mode, 77, param)
    ref.scan(sweeps)
    want_avg, want_smp = ref.avg(), ref.samples()
    w = ref.window_coefs() if plan["bin_e"] else None
    reads, hops = make_reads(ref.lib, plan, sweeps, mode, seed=77, param=param)
    avg, smp, db = gpu_scan_device(scan_mod, plan, w, reads)
    assert np.array_equal(avg, want_avg), name
    assert np.array_equal(smp, want_smp), name
    # rows: the reference prints them itself (csv_dbm, rtl_power.c:722-765); ours come from host/rtl_power_plan.c
    host_plan = plan_scan(freq, crop, None if fir < 0 else fir)
    for h in sorted({0, plan["tune_count"] - 1}):      # (csv_dbm zeroes the hop: one call per hop)
        assert host_plan.csv_row(h, int(smp[h]), db[h]) == ref.csv(h), (name, h)


@pytest.mark.parametrize("bin_e,tc,sweeps,peak", [(12, 7, 45, 0), (12, 600, 1, 0), (12, 3, 1, 1), (12, 1, 1, 0),
                                                  (11, 5, 130, 1), (9, 4, 150, 0), (5, 3, 200, 1), (1, 2, 160, 0)])
def test_equal_run_schedule(scan_mod, port_oracle, bin_e, tc, sweeps, peak):
    """scan_small_kernel's persistent CTAs share the working sets in equal runs: runs start / end in the middle
    of a read (half reads), cross hop boundaries (flush + restart of the register accumulators, through the
    shared bin array below 4096 bins) and may be shorter than one read."""
    n = 1 << bin_e
    plan = plan_dict(bin_e, peak_hold=peak, tune_count=tc, crop=0.0)
    w = port_oracle.window_coefs(WINDOWS[(bin_e + tc) % len(WINDOWS)], n)
    reads, hops = make_reads(port_oracle.lib, plan, sweeps, SYNTH_BIASED, seed=bin_e + tc, param=14)
    reads[1 % len(reads), :] = 255
    want = expected(port_oracle, plan, w, reads, hops)
    got = gpu_scan_device(scan_mod, plan, w, reads)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


def test_submit_reads_any_hop_order(scan_mod, port_oracle):
    """randomised hopping (reference TODO, rtl_power.c:29-36): bins do not depend on the visiting order"""
    plan = plan_dict(12, tune_count=11, crop=0.1)
    b = plan["buf_len"]
    w = port_oracle.window_coefs("hamming", 4096)
    reads, hops = make_reads(port_oracle.lib, plan, 24, SYNTH_BIASED, seed=8, param=-9)
    want = expected(port_oracle, plan, w, reads, hops)
    rng = np.random.default_rng(3)
    for trial in range(3):
        order = rng.permutation(len(reads))
        pinned = scan_mod.PinnedBuffer(len(reads) * b)
        pinned.view(np.uint8, (len(reads), b))[:] = reads[order]
        g = scan_mod.GpuScan.from_plan(plan, window_coefs=w)
        g.submit_reads(hops[order], pinned.ptr)
        got = g.collect_all()
        g.close()
        pinned.free()
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and db_close(got[2], want[2])
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w)
    pinned = scan_mod.PinnedBuffer(2 * b)
    with pytest.raises(scan_mod.ScanError) as e:
        g.submit_reads([0, 11], pinned.ptr)      # hop out of range
    assert e.value.code == -3
    g.close()


def test_iir_smoothing_across_reports(scan_mod, port_oracle):
    """-s iir: s = d at the first report, s += alpha (d - s) afterwards on the linear value csv_dbm logs;
    raw bins / counts untouched; hops without samples print d and keep their state"""
    alpha = 0.25
    plan = plan_dict(8, tune_count=3, crop=0.2, rate=2000000)
    w = port_oracle.window_coefs("blackman", 256)
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w, iir_alpha=alpha)
    state = None
    for rep in range(4):
        reads, hops = make_reads(port_oracle.lib, plan, 2 + rep, SYNTH_BIASED, seed=rep, param=5 * rep)
        if rep == 2:                      # hop 1 gets no reads in this interval
            keep = hops != 1
            reads, hops = reads[keep], hops[keep]
        avg_w, smp_w, db_w = expected(port_oracle, plan, w, reads, hops)
        for r, h in zip(reads, hops):
            g.submit(int(h), r)
        avg, smp, db = g.collect_all()
        assert np.array_equal(avg, avg_w) and np.array_equal(smp, smp_w)
        n = 256
        i1 = int(n * 0.2 * 0.5)          # csv_dbm's crop (rtl_power.c:741-748): bins i1 .. n-1-i1 of the swapped spectrum
        want = np.empty_like(db_w)
        if state is None:
            state = np.full((3, n - 2 * i1), np.nan)
        for h in range(3):
            a = avg_w[h].copy()
            a[0] = a[1]                                         # DC nuke before the half swap (:732-738)
            sw = np.concatenate([a[n // 2:], a[:n // 2]])
            with np.errstate(all="ignore"):
                d = sw[i1: n - i1].astype(np.float64) / float(plan["rate"]) / float(smp_w[h])
            if smp_w[h] != 0:
                new = np.where(np.isnan(state[h]), d, state[h] + alpha * (d - state[h]))
                state[h] = new
            else:
                new = d
            with np.errstate(all="ignore"):
                want[h, :-1] = 10 * np.log10(new)
            want[h, -1] = want[h, -2]
        assert db_close(db, want), rep
    g.close()


def test_first_release_cfg_size_still_accepted(scan_mod):
    """struct_size of the first release (without iir_alpha) must keep working: ABI versioning"""
    L = scan_mod.load_library()
    cfg = scan_mod._Cfg()
    cfg.struct_size = scan_mod._Cfg.iir_alpha.offset
    cfg.device, cfg.tune_count, cfg.bin_e, cfg.buf_len = 0, 1, 10, 16384
    cfg.downsample, cfg.boxcar, cfg.rate, cfg.crop = 1, 1, 2400000, 0.0
    cfg.iir_alpha = 123.0       # garbage beyond the declared size must be ignored
    h = ctypes.c_void_p()
    assert L.rtlsdr_gpu_scan_init(ctypes.byref(cfg), ctypes.byref(h)) == 0
    L.rtlsdr_gpu_scan_close(h)
    cfg.struct_size = 12
    assert L.rtlsdr_gpu_scan_init(ctypes.byref(cfg), ctypes.byref(h)) == -2


def _run_sweep_main(args, out, world, tmp_path, timeout=600):
    base = [sys.executable]
    env = dict(os.environ)
    if world > 1:
        base += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                 "--master-port", str(29500 + os.getpid() % 2000), "-m", "rtlsdr_b200.sweep_main",
                 "--backend", "gloo", "--device", "0"]
    else:
        base += ["-m", "rtlsdr_b200.sweep_main"]
    r = subprocess.run(base + args + ["-o", str(out)], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    return out.read_text()


def test_sweep_main_two_ranks_three_intervals(scan_mod, tmp_path):
    """the sharded driver with more than one rank AND more than one interval (pinned input
    cubes and report buffers are reused): rows byte-identical to the 1-rank run.  The two ranks share
    cuda:0 and exchange through gloo (NCCL refuses two ranks on one device); on a multi-GPU box the same
    driver runs over NCCL / peer memory (tools/sweep_check.sh)."""
    args = ["-f", "88M:108M:25k", "-c", "20%", "-w", "hamming", "--sweeps", "5", "--intervals", "3",
            "--synth", "biased", "--seed", "4", "--param", "17"]
    one = _run_sweep_main(args, tmp_path / "one.csv", 1, tmp_path)
    two = _run_sweep_main(args, tmp_path / "two.csv", 2, tmp_path)
    three = _run_sweep_main(args, tmp_path / "three.csv", 3, tmp_path)
    assert len(one) > 1000 and one == two == three
    shuffled = _run_sweep_main(args + ["--random-hops", "5"], tmp_path / "shuffled.csv", 2, tmp_path)
    assert shuffled == one


def _run_cli(args, out, env_extra):
    from rtlsdr_b200 import _build
    _build.build_host()
    exe = os.path.join(_build.HOST_BUILD, "rtl_power_gpu")
    env = dict(os.environ, RTLSDR_SYNTH_MODE="biased", RTLSDR_SYNTH_SEED="3", RTLSDR_SYNTH_PARAM="21",
               RTL_POWER_PASSES="4", RTL_POWER_REPORTS="3", RTL_POWER_TIMESTAMP="2026-01-01, 00:00:00", **env_extra)
    r = subprocess.run([exe] + args + [str(out)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return out.read_text(), r.stderr


def test_cli_workers_and_random_hops_same_rows(scan_mod, tmp_path):
    """rtl_power_gpu -t N (GPU workers = the reference's 'multiple FFT workers' TODO) and -R (randomised hopping):
    the CSV bytes do not depend on either.  Two / three workers on device 0 here; on a multi-GPU box
    RTLSDR_GPU_DEVICES names distinct devices."""
    import torch
    args = ["-f", "88M:108M:25k", "-c", "20%", "-w", "hamming"]
    one, _ = _run_cli(args, tmp_path / "w1.csv", {})
    two, err = _run_cli(args + ["-t", "2"], tmp_path / "w2.csv", {"RTLSDR_GPU_DEVICES": "0,0"})
    assert "GPU workers: 2" in err
    three, _ = _run_cli(args + ["-t", "3", "-R", "9"], tmp_path / "w3.csv", {"RTLSDR_GPU_DEVICES": "0,0,0"})
    assert len(one) > 1000 and one == two == three
    if torch.cuda.device_count() >= 2:
        real, _ = _run_cli(args + ["-t", "2"], tmp_path / "w2real.csv", {})
        assert real == one
    # more workers than hops (a single-hop scan, 2^13 bins with peak hold): the SWEEPS are dealt to the workers and
    # their raw accumulators merged into worker 0's at every report -- same bytes again
    args = ["-f", "100M:102.4M:300", "-w", "blackman-harris", "-P"]
    one, _ = _run_cli(args, tmp_path / "s1.csv", {})
    three, err = _run_cli(args + ["-t", "3"], tmp_path / "s3.csv", {"RTLSDR_GPU_DEVICES": "0,0,0"})
    assert "every 3-th sweep" in err and len(one) > 1000 and one == three


def test_cli_iir_rows(scan_mod, port_oracle, tmp_path):
    """-s iir through the CLI: first report equals the plain one, later reports are smoothed"""
    args = ["-f", "433M:435M:4k", "-w", "blackman"]
    plain, _ = _run_cli(args, tmp_path / "avg.csv", {})
    iir, _ = _run_cli(args + ["-s", "iir"], tmp_path / "iir.csv", {"RTL_POWER_IIR_ALPHA": "0.5"})
    p, q = plain.splitlines(), iir.splitlines()
    assert len(p) == len(q) == 3
    assert p[0] == q[0] and p[1] != q[1]
    # second report = 10 log10(0.5 (d1 + d2)) within print precision
    d = [np.array([float(x) for x in line.split(", ")[6:]]) for line in p]
    got = np.array([float(x) for x in q[1].split(", ")[6:]])
    want = 10 * np.log10(0.5 * (10 ** (d[0] / 10) + 10 ** (d[1] / 10)))
    assert np.allclose(got, want, atol=0.02)


def test_kat_rows_through_host_cube(scan_mod):
    """the bench's verification path in small: synthetic cube generated by the host library (threads), pinned
    submit_batch, FNV of the bins == SURVEY.md 8(c) rows for the two sharded bench workloads"""
    from rtlsdr_b200.planner import fnv1a_int64 as fnv_c, plan_scan, synth_cube
    for freq, fir, want in (("24M:1457.6M:700", None, 0x7b1c7343a9686225), ("24M:1766M:1k", 9, 0x074f712a23c886d1)):
        plan = plan_scan(freq, 0.0, fir).as_dict()
        plan["peak_hold"] = 0
        tc, b = plan["tune_count"], plan["buf_len"]
        pinned = scan_mod.PinnedBuffer(tc * b)
        synth_cube(pinned.ptr, SYNTH_XORSHIFT, 0, 0, tc, 0, tc, 0, 1, b)
        g = scan_mod.GpuScan.from_plan(plan, window_coefs=scan_mod.window_coefs("rectangle", 4096))
        g.submit_batch(0, tc, 1, pinned.ptr, tc * b, b)
        avg, smp, _ = g.collect_all(want_db=False)
        g.close()
        pinned.free()
        assert fnv_c(avg) == want == fnv1a_int64(avg)
        assert (smp == 2).all()


@pytest.mark.parametrize("bin_e,reads_n,peak", [(17, 150, 0), (17, 77, 1), (15, 300, 0), (18, 40, 0), (13, 700, 1)])
def test_large_path_many_reads_shuffled_hops(scan_mod, port_oracle, bin_e, reads_n, peak):
    """large-FFT path on one batch of many hop visits in shuffled order (submit_reads): unequal hop sizes, hop
    changes inside round C's 16-read runs, sums / peaks of repeated reads"""
    n = 1 << bin_e
    tc = 3
    plan = plan_dict(bin_e, buf_len=2 * n, peak_hold=peak, tune_count=tc, crop=0.25)
    w = port_oracle.window_coefs("blackman-harris", n)
    rng = np.random.default_rng(bin_e + reads_n)
    # a handful of distinct reads, repeated (the oracle's time goes with the number of DISTINCT reads below)
    base_reads, _ = make_reads(port_oracle.lib, dict(plan, tune_count=1), 6, SYNTH_BIASED, seed=bin_e, param=27)
    base_reads[1, ::5] = 255
    pick = rng.integers(0, len(base_reads), reads_n)
    hops = np.sort(rng.integers(0, tc, reads_n)).astype(np.int32)     # unequal hop sizes, hop changes mid-CTA
    # expected: per hop, sum / max over its reads of the per-read spectrum
    one = {}
    for j in range(len(base_reads)):
        a, _, _ = expected(port_oracle, dict(plan, tune_count=1), w, base_reads[j:j + 1], np.zeros(1, np.int32))
        one[j] = a[0]
    want = np.zeros((tc, n), dtype=np.int64)
    smp = np.zeros(tc, dtype=np.int32)
    for j, h in zip(pick, hops):
        want[h] = np.maximum(want[h], one[j]) if peak else want[h] + one[j]
        smp[h] += 1
    # ONE batch of reads_n hop visits in shuffled order (submit_reads): the library sorts them by hop and walks
    # them chunk by chunk
    order = rng.permutation(reads_n)
    pinned = scan_mod.PinnedBuffer(reads_n * 2 * n)
    pinned.view(np.uint8, (reads_n, 2 * n))[:] = base_reads[pick[order]]
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w)
    g.submit_reads(hops[order], pinned.ptr)
    avg, got_smp, _ = g.collect_all(want_db=False)
    g.close()
    pinned.free()
    assert np.array_equal(avg, want)
    assert np.array_equal(got_smp, smp)


def test_bench_two_ranks_on_one_gpu(scan_mod, tmp_path):
    """bench.py's multi-rank control flow (hop sharding, verification against the known answer and the 1-rank run,
    timed rounds, host-buffer leg, second sharded workload) with two ranks sharing cuda:0 and gloo as the
    exchange (BENCH_GLOO_ONE_GPU=1) at a reduced size: every collective must be entered by both ranks."""
    import json
    env = dict(os.environ, BENCH_GLOO_ONE_GPU="1", BENCH_E2E_WEIGHTS="3,1")   # host-fed leg with unequal hop shares
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + os.getpid() % 500), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "4",
           "--warmup", "3", "--sweeps", "8", "--no-cpu"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    assert line["config"]["hops_per_gpu"] == [256, 256]
    v = line["verify"]
    assert v["ok"] and v["kat_ok"] and v["interval_ok"] and v["kat_fnv"] == "7b1c7343a9686225"
    assert v["interval_ranks"] == 2 and v["interval_fnv"] == v["interval_fnv_check"]
    c3 = line["companions"][0]
    assert c3["verify"]["ok"] and c3["hops_per_gpu"] == [312, 311]
    c4 = line["companions"][1]     # the single-hop scan with its READS sharded and merged on rank 0
    assert c4["verify"]["ok"] and c4["reads_per_gpu"] == [1024, 1024] and c4["value"] > 0
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] >= 4
    assert line["e2e"]["hops_per_gpu"] == [384, 128] and line["e2e"]["report_fnv_equals_verified_interval"] is True


@pytest.mark.parametrize("bin_e,peak", [(12, 0), (10, 1), (14, 0), (0, 0)])
def test_async_report_back_to_back_intervals(scan_mod, port_oracle, bin_e, peak):
    """RTLSDR_GPU_FLAG_ASYNC_REPORT: collect_device() reports on the handle's report stream and flips to the second
    accumulator set; five intervals are queued back to back without any host synchronisation, every report must
    equal the oracle's for ITS interval (no leakage between the two sets, read-and-zero intact), and a host
    collect afterwards sees only what was submitted after the last device collect"""
    import torch
    n = 1 << bin_e
    tc = 3
    buf_len = max(16384, 2 * n)
    plan = plan_dict(bin_e, buf_len=buf_len, peak_hold=peak, tune_count=tc, crop=0.2 if bin_e else 0.0, rate=1000000 if not bin_e else 2400000)
    w = port_oracle.window_coefs("hamming", n) if bin_e else None
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w, async_report=True)
    assert g.get_report_stream() != g.get_stream()
    keep, outs, wants = [], [], []
    for it in range(5):
        reads, hops = make_reads(port_oracle.lib, plan, 2 + it, SYNTH_BIASED, seed=100 + it, param=9 * it - 20)
        wants.append(expected(port_oracle, plan, w if bin_e else np.zeros(1, np.int32), reads, hops))
        dev = torch.from_numpy(reads).cuda()
        keep.append(dev)
        o_avg = torch.full((tc, n), -1, dtype=torch.int64, device="cuda")
        o_smp = torch.full((tc,), -1, dtype=torch.int32, device="cuda")
        o_db = torch.zeros((tc, g.db_count), dtype=torch.float64, device="cuda")
        outs.append((o_avg, o_smp, o_db))
        torch.cuda.synchronize()          # inputs / outputs exist before the handle's streams touch them
        g.submit_device(0, tc, 2 + it, dev.data_ptr(), tc * buf_len, buf_len)
        g.collect_device(o_avg.data_ptr(), o_smp.data_ptr(), o_db.data_ptr())
    # something submitted after the last device collect, reported by a HOST collect
    reads, hops = make_reads(port_oracle.lib, plan, 2, SYNTH_BIASED, seed=200, param=3)
    for r, h in zip(reads, hops):
        g.submit(int(h), r)
    tail = g.collect_all()
    torch.cuda.synchronize()
    for it, (want, (o_avg, o_smp, o_db)) in enumerate(zip(wants, outs)):
        assert np.array_equal(o_avg.cpu().numpy(), want[0]), it
        assert np.array_equal(o_smp.cpu().numpy(), want[1]), it
        assert db_close(o_db.cpu().numpy(), want[2]), it
    want_tail = expected(port_oracle, plan, w if bin_e else np.zeros(1, np.int32), reads, hops)
    assert np.array_equal(tail[0], want_tail[0]) and np.array_equal(tail[1], want_tail[1]) and db_close(tail[2], want_tail[2])
    g.close()


def test_flag_signal_and_sleeping_wait(scan_mod):
    """the report hand-off kernels: a wait on stream A completes only after both flags were raised on stream B
    (wrap-safe compare), and a wait that is never satisfied gives up after its time-out and says so"""
    import time
    import torch
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    marker = torch.zeros(1, dtype=torch.int32, device="cuda")
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    scan_mod.flag_wait(a.cuda_stream, flags.data_ptr(), 2, 5, 5000, flags.data_ptr() + 12)
    with torch.cuda.stream(a):
        marker.fill_(1)
    scan_mod.flag_signal(b.cuda_stream, flags.data_ptr(), 5)
    b.synchronize()
    time.sleep(0.05)
    assert not a.query()                      # flag 1 still missing: the waiter sleeps, the marker is unset
    scan_mod.flag_signal(b.cuda_stream, flags.data_ptr() + 4, 7)   # 7 >= 5
    a.synchronize()
    assert int(marker.item()) == 1 and flags.tolist() == [5, 7, 0, 0]
    # time-out: nobody raises flag 2
    t0 = time.perf_counter()
    scan_mod.flag_wait(a.cuda_stream, flags.data_ptr() + 8, 1, 1, 60, flags.data_ptr() + 12)
    a.synchronize()
    assert 0.04 < time.perf_counter() - t0 < 2.0
    assert flags.tolist() == [5, 7, 0, 1]
    with pytest.raises(scan_mod.ScanError):
        scan_mod.flag_wait(a.cuda_stream, flags.data_ptr() + 1, 1, 1)      # misaligned


def test_short_reads_keep_the_hops_previous_tail(scan_mod, port_oracle):
    """RTLSDR_GPU_FLAG_SHORT_READS: a short rtlsdr_read_sync overwrites the first n_read bytes of tunes[hop].buf8,
    the rest still holds the hop's previous read, and the reference processes the whole buffer
    (rtl_power.c:657-659).  Without the flag a short length is an error."""
    plan = plan_dict(10, tune_count=2, crop=0.0)
    b = plan["buf_len"]
    w = port_oracle.window_coefs("hamming", 1024)
    reads, hops = make_reads(port_oracle.lib, plan, 3, SYNTH_BIASED, seed=77, param=15)
    lens = [b, b, 5000, b, b, 1]            # hop 0: full, short (5000), full; hop 1: full, full, 1 byte
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w, short_reads=True)
    shadow = np.zeros((2, b), dtype=np.uint8)
    seen = []
    for r, h, n in zip(reads, hops, lens):
        shadow[h, :n] = r[:n]
        seen.append(shadow[h].copy())
        scan_mod.load_library().rtlsdr_gpu_scan_submit(g.h, int(h), r.ctypes.data, n)
    got = g.collect_all()
    g.close()
    want = expected(port_oracle, plan, w, np.stack(seen), hops)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and db_close(got[2], want[2])
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w)
    assert scan_mod.load_library().rtlsdr_gpu_scan_submit(g.h, 0, reads[0].ctypes.data, 5000) == -4
    assert scan_mod.load_library().rtlsdr_gpu_scan_submit(g.h, 0, reads[0].ctypes.data, b + 16) == -4
    g.close()


@pytest.mark.parametrize("bin_e,peak,buf_len,async_report", [(10, 0, 16384, False), (12, 1, 16384, True),
                                                             (13, 1, 16384, False), (0, 0, 16384, False)])
def test_merge_device_reads_of_one_hop_on_three_handles(scan_mod, port_oracle, bin_e, peak, buf_len, async_report):
    """VERDICT r1 missing #6: the READS of the same hops split over several handles (one per GPU on a multi-GPU
    box, three on cuda:0 here), raw accumulators collected on the device and folded into the first handle with
    rtlsdr_gpu_scan_merge_device(): bins, counts AND dB identical to the oracle's for all the reads, i.e. to one
    handle -- int64 sums / peak-hold maxima are exact in any grouping (rtl_power.c:708-717)."""
    import torch
    n, tc, passes = 1 << bin_e, 2, 7
    plan = plan_dict(bin_e, buf_len=buf_len, peak_hold=peak, tune_count=tc, crop=0.1 if bin_e else 0.0)
    win = port_oracle.window_coefs("hamming", n) if bin_e else None
    reads, hops = make_reads(port_oracle.lib, plan, passes, SYNTH_BIASED, seed=bin_e + 3, param=29)
    owin = win if win is not None else np.zeros(1, np.int32)      # the oracle wants an array even for rms bins
    want_avg, want_smp, want_db = expected(port_oracle, plan, owin, reads, hops)
    shares = [(0, 3), (3, 5), (5, 7)]                      # sweeps per handle
    gs = [scan_mod.GpuScan.from_plan(plan, window_coefs=win, async_report=async_report) for _ in shares]
    try:
        words = tc * n + (tc + 1) // 2                     # [avg tc*N int64 | samples tc x int32]
        slots = torch.zeros(len(shares), words, dtype=torch.int64, device="cuda")
        dev = torch.from_numpy(np.ascontiguousarray(reads)).cuda()
        torch.cuda.synchronize()
        for k, (g, (lo, hi)) in enumerate(zip(gs, shares)):
            g.submit_device(0, tc, hi - lo, dev[lo * tc:].data_ptr(), tc * buf_len, buf_len)
            base = slots[k].data_ptr()
            g.collect_device(base, base + tc * n * 8, None)
        torch.cuda.synchronize()                           # the reports may be on the handles' report streams
        # every handle (the merging one included) is empty again; fold all three sets into handle 0
        gs[0].merge_device(slots.data_ptr(), slots.data_ptr() + tc * n * 8, len(shares), words * 8)
        avg, smp, db = gs[0].collect_all()
        assert np.array_equal(avg, want_avg) and np.array_equal(smp, want_smp)
        assert db_close(db, want_db)
        # read-and-zero still holds after a merge, and per-hop collects see merged counts too
        gs[0].merge_device(slots.data_ptr(), slots.data_ptr() + tc * n * 8, 2, words * 8)
        a1, s1, _ = gs[0].collect(1)
        sub = np.concatenate([reads[0:3 * tc], reads[3 * tc:5 * tc]]), np.concatenate([hops[0:3 * tc], hops[3 * tc:5 * tc]])
        w_avg, w_smp, _ = expected(port_oracle, plan, owin, *sub)
        assert np.array_equal(a1, w_avg[1]) and s1 == w_smp[1]
        a0, s0, _ = gs[0].collect(0)
        assert np.array_equal(a0, w_avg[0]) and s0 == w_smp[0]
        avg, smp, _ = gs[0].collect_all()
        assert not avg.any() and not smp.any()
    finally:
        for g in gs:
            g.close()


def test_merge_device_argument_checks(scan_mod):
    import torch
    g = scan_mod.GpuScan(1, 4, 16384)
    try:
        buf = torch.zeros(64, dtype=torch.int64, device="cuda")
        L = g.lib
        assert L.rtlsdr_gpu_scan_merge_device(None, buf.data_ptr(), buf.data_ptr(), 1, 0) == -1
        assert L.rtlsdr_gpu_scan_merge_device(g.h, None, buf.data_ptr(), 1, 0) == -1
        assert L.rtlsdr_gpu_scan_merge_device(g.h, buf.data_ptr(), buf.data_ptr(), 0, 0) == -2
        assert L.rtlsdr_gpu_scan_merge_device(g.h, buf.data_ptr(), buf.data_ptr(), 2, 0) == -2
        assert L.rtlsdr_gpu_scan_merge_device(g.h, buf.data_ptr() + 4, buf.data_ptr(), 1, 0) == -8
        assert L.rtlsdr_gpu_scan_merge_device(g.h, buf.data_ptr(), buf.data_ptr(), 2, 12) == -8
    finally:
        g.close()


@pytest.mark.parametrize("args", [
    ["-f", "100M:102.4M:2400", "--sweeps", "7", "--intervals", "3", "--synth", "biased", "--seed", "2", "--param", "33"],
    ["-f", "100M:102.4M:300", "-w", "blackman-harris", "-P", "--sweeps", "5", "--intervals", "2", "--seed", "6"],
    ["-f", "100M:104M:1M", "--sweeps", "6", "--intervals", "2", "--synth", "biased", "--param", "9"],
])
def test_sweep_main_read_sharded_single_hop(scan_mod, tmp_path, args):
    """--shard reads (BASELINE configs 1 and 4 are single-hop scans: nothing to shard by hop): 1, 2 or 3 ranks on
    cuda:0 over gloo, several intervals through the alternating buffers -- CSV byte-identical to the plain
    hop-sharded 1-rank run (splits 7 -> 3+2+2, 5 -> 3+2, 6 -> 2+2+2; 1024 bins, 2^13 bins with peak hold, rms bins)."""
    plain = _run_sweep_main(args, tmp_path / "plain.csv", 1, tmp_path)
    world = 2 if "-P" in args else 3                        # (one multi-rank launch per case keeps the suite short)
    many = _run_sweep_main(args + ["--shard", "reads"], tmp_path / "rn.csv", world, tmp_path)
    assert len(plain) > 100 and plain == many
    if "-P" not in args and "--param" in args and args[args.index("--param") + 1] == "33":
        one = _run_sweep_main(args + ["--shard", "reads"], tmp_path / "r1.csv", 1, tmp_path)
        assert one == plain
