"""GPU parity tests proper: the CUDA path, called through the C ABI
(include/rtlsdr_gpu_scan.h), against the oracles on identical synthetic bytes.
Bit-exact for int64 bins and sample counts; dB within 1e-6 relative."""
import numpy as np
import pytest

from oracles import (SYNTH_BIASED, SYNTH_CONST, SYNTH_COUNTER, SYNTH_TONE, SYNTH_XORSHIFT, WINDOWS,
                     fnv1a_int64)
from scan_cases import KAT_ROWS, db_close, expected, make_reads, plan_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scan_mod():
    import rtlsdr_b200.scan as s
    s.load_library()  # must exist on a GPU box: no fallback
    return s


def run_gpu(scan_mod, plan, window, reads, hops, how="submit"):
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=window)
    try:
        if how == "submit":
            for r, h in zip(reads, hops):
                g.submit(int(h), r)
        else:
            import torch
            tc = plan["tune_count"]
            passes = len(reads) // tc
            dev = torch.from_numpy(np.ascontiguousarray(reads)).cuda()
            torch.cuda.synchronize()
            b = plan["buf_len"]
            g.submit_device(0, tc, passes, dev.data_ptr(), tc * b, b)
            g.sync()
        avg, smp, db = g.collect_all()
        # accumulators are zero after a collect (rtl_power.c:761-764)
        avg2, smp2, _ = g.collect_all(want_db=False)
        assert not avg2.any() and not smp2.any()
        return avg, smp, db
    finally:
        g.close()


@pytest.mark.parametrize("bin_e", list(range(1, 13)))
@pytest.mark.parametrize("peak", [0, 1])
def test_u8_path_all_sizes(scan_mod, port_oracle, bin_e, peak):
    n = 1 << bin_e
    rng = np.random.default_rng(bin_e * 2 + peak)
    plan = plan_dict(bin_e, peak_hold=peak, tune_count=3, crop=0.1 if bin_e > 3 else 0.0)
    window = port_oracle.window_coefs(WINDOWS[bin_e % len(WINDOWS)], n)
    if bin_e % 5 == 0:
        window = rng.integers(-70000, 70000, n).astype(np.int32)  # only the low 16 bits can matter
    reads, hops = make_reads(port_oracle.lib, plan, 4, SYNTH_BIASED, seed=bin_e, param=11)
    reads[1, :] = 255
    reads[2, :] = 0
    reads[5, 0::2] = 255
    want = expected(port_oracle, plan, window, reads, hops)
    got = run_gpu(scan_mod, plan, window, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


@pytest.mark.parametrize("window", WINDOWS)
@pytest.mark.parametrize("mode,param", [(SYNTH_XORSHIFT, 0), (SYNTH_COUNTER, 0), (SYNTH_CONST, 127),
                                        (SYNTH_CONST, 255), (SYNTH_TONE, 126)])
def test_windows_and_input_classes(scan_mod, port_oracle, window, mode, param):
    plan = plan_dict(10, tune_count=2, rate=2400000)
    w = port_oracle.window_coefs(window, 1024)
    reads, hops = make_reads(port_oracle.lib, plan, 3, mode, seed=5, param=param)
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops, how="device")
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


@pytest.mark.parametrize("row", KAT_ROWS, ids=[f"{r[0]}-{r[2]}-F{r[3]}-P{r[4]}" for r in KAT_ROWS])
def test_survey_known_answers(scan_mod, port_oracle, row):
    """FNV-1a of the reference's int64 bins, generated from the
    unmodified reference; the GPU path must reproduce every hash."""
    from rtlsdr_b200.planner import plan_scan
    freq, crop, window, fir, peak, passes, mode, fnv = row
    plan = plan_scan(freq, crop, None if fir < 0 else fir).as_dict()
    plan["peak_hold"] = peak
    w = port_oracle.window_coefs(window, 1 << plan["bin_e"])
    reads, hops = make_reads(port_oracle.lib, plan, passes, mode, seed=0, param=0)
    how = "device" if plan["tune_count"] > 100 else "submit"
    avg, smp, db = run_gpu(scan_mod, plan, w, reads, hops, how=how)
    assert fnv1a_int64(avg) == fnv


@pytest.mark.parametrize("freq,window,fir,peak", [
    ("100M:100.1M:100", "rectangle", -1, 0),   # boxcar ds=28, B=57344
    ("100M:100.1M:100", "blackman", 9, 0),     # 4 x fifth_order + 9-tap FIR
    ("100M:100.1M:100", "youssef", 0, 1),      # 4 x fifth_order, no FIR, peak hold
    ("100M:100.5M:10k", "bartlett", -1, 0),    # ds=5, clamped buffer, partial tail block
    ("100M:100.3M:3k", "hamming", -1, 1),      # ds=9
    ("100M:100.9M:30k", "hamming", -1, 0),     # ds=3, odd l_len
    ("100M:100.01M:50", "hamming", 9, 0),      # 8 passes, B=131072
    ("100M:100.02M:100", "hamming", -1, 0),    # boxcar ds=140: thread-per-slot kernel, 16-byte loads
    ("100M:100.03M:100", "blackman", -1, 1),   # boxcar ds=93
])
@pytest.mark.parametrize("mode,param", [(SYNTH_XORSHIFT, 0), (SYNTH_BIASED, 40), (SYNTH_CONST, 255)])
def test_decimating_paths(scan_mod, port_oracle, freq, window, fir, peak, mode, param):
    from rtlsdr_b200.planner import plan_scan
    plan = plan_scan(freq, 0.0, None if fir < 0 else fir).as_dict()
    plan["peak_hold"] = peak
    w = port_oracle.window_coefs(window, 1 << plan["bin_e"])
    reads, hops = make_reads(port_oracle.lib, plan, 3, mode, seed=9, param=param)
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


@pytest.mark.parametrize("peak", [0, 1])
def test_rms_path(scan_mod, port_oracle, peak):
    plan = plan_dict(0, tune_count=5, peak_hold=peak, rate=1000000)
    reads, hops = make_reads(port_oracle.lib, plan, 4, SYNTH_BIASED, seed=2, param=25)
    want = expected(port_oracle, plan, np.zeros(1, np.int32), reads, hops)
    got = run_gpu(scan_mod, plan, None, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


def test_per_hop_collect_and_reaccumulate(scan_mod, port_oracle):
    """collect(hop) zeroes only that hop; later submits accumulate from zero."""
    plan = plan_dict(9, tune_count=3)
    w = port_oracle.window_coefs("hamming", 512)
    reads, hops = make_reads(port_oracle.lib, plan, 2, SYNTH_XORSHIFT, seed=1)
    want = expected(port_oracle, plan, w, reads, hops)
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w)
    try:
        for r, h in zip(reads, hops):
            g.submit(int(h), r)
        a1, s1, d1 = g.collect(1)
        assert np.array_equal(a1, want[0][1]) and s1 == want[1][1] and db_close(d1, want[2][1])
        for r, h in zip(reads, hops):
            g.submit(int(h), r)
        a1b, s1b, _ = g.collect(1)
        assert np.array_equal(a1b, want[0][1]) and s1b == want[1][1]
        a0, s0, _ = g.collect(0)
        assert np.array_equal(a0, 2 * want[0][0]) and s0 == 2 * want[1][0]
    finally:
        g.close()


def test_error_codes(scan_mod):
    g = scan_mod.GpuScan(2, 10, 16384)
    try:
        with pytest.raises(scan_mod.ScanError) as e:
            g.submit(2, np.zeros(16384, np.uint8))
        assert e.value.code == -3
        with pytest.raises(scan_mod.ScanError) as e:
            g.submit(0, np.zeros(100, np.uint8))
        assert e.value.code == -4
    finally:
        g.close()
    with pytest.raises(scan_mod.ScanError) as e:
        scan_mod.GpuScan(0, 10, 16384)
    assert e.value.code == -2


@pytest.mark.parametrize("bin_e,peak,window", [(13, 0, "hamming"), (14, 1, "blackman"), (15, 0, "rectangle"),
                                               (16, 0, "youssef"), (17, 1, "blackman-harris"),
                                               (18, 0, "bartlett"), (20, 0, "hann-poisson"), (21, 1, "hamming")])
def test_large_fft_path(scan_mod, port_oracle, bin_e, peak, window):
    """N >= 8192: one FFT block per read, three-round transform through global scratch."""
    n = 1 << bin_e
    plan = plan_dict(bin_e, buf_len=2 * n, peak_hold=peak, tune_count=2, crop=0.25)
    w = port_oracle.window_coefs(window, n)
    reads, hops = make_reads(port_oracle.lib, plan, 2, SYNTH_BIASED, seed=bin_e, param=-17)
    reads[1, : n // 2] = 255
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


def test_large_fft_with_decimation(scan_mod, port_oracle):
    """narrow scan with many bins: boxcar ds=3 feeding a 2^13-point transform."""
    from rtlsdr_b200.planner import plan_scan
    plan = plan_scan("100M:100.9M:120", 0.0).as_dict()
    assert plan["bin_e"] == 13 and plan["downsample"] == 3
    plan["peak_hold"] = 0
    w = port_oracle.window_coefs("hamming", 1 << plan["bin_e"])
    reads, hops = make_reads(port_oracle.lib, plan, 3, SYNTH_BIASED, seed=4, param=21)
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])
    plan = plan_scan("100M:100.9M:120", 0.0, 9).as_dict()  # -F 9: 1 fifth_order pass + FIR
    plan["peak_hold"] = 1
    w = port_oracle.window_coefs("blackman", 1 << plan["bin_e"])
    reads, hops = make_reads(port_oracle.lib, plan, 2, SYNTH_TONE, seed=4, param=100)
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])


def test_cli_csv_matches_reference_rows(scan_mod, port_oracle, tmp_path):
    """rtl_power_gpu (host/rtl_power_gpu.c): same command line, same CSV bytes as the
    reference would print for the same synthetic bytes (fixed timestamp, sweep-count interval)."""
    import os
    import subprocess
    from rtlsdr_b200 import _build
    from rtlsdr_b200.planner import plan_scan
    _build.build_host()
    exe = os.path.join(_build.HOST_BUILD, "rtl_power_gpu")
    out = tmp_path / "scan.csv"
    env = dict(os.environ, RTLSDR_SYNTH_MODE="biased", RTLSDR_SYNTH_SEED="3", RTLSDR_SYNTH_PARAM="21",
               RTL_POWER_PASSES="4", RTL_POWER_TIMESTAMP="2026-01-01, 00:00:00")
    r = subprocess.run([exe, "-f", "88M:108M:25k", "-c", "20%", "-w", "hamming", "-1", str(out)],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "Number of frequency hops: 9" in r.stderr
    plan = plan_scan("88M:108M:25k", 0.2)
    pd = plan.as_dict()
    pd["peak_hold"] = 0
    w = port_oracle.window_coefs("hamming", 1 << pd["bin_e"])
    reads, hops = make_reads(port_oracle.lib, pd, 4, SYNTH_BIASED, seed=3, param=21)
    avg, smp, db = expected(port_oracle, pd, w, reads, hops)
    want = "".join("2026-01-01, 00:00:00, " + plan.csv_row(h, int(smp[h]), db[h]) for h in range(pd["tune_count"]))
    assert out.read_text() == want


@pytest.mark.parametrize("bin_e,ds", [(1, 7), (2, 3), (3, 5), (4, 2), (6, 11), (12, 2)])
def test_decimating_small_n(scan_mod, port_oracle, bin_e, ds):
    """packed decimated images: several reads per working set, N down to 2 (alignment padding)"""
    n = 1 << bin_e
    buf_len = max(16384, 2 * n * ds)
    plan = plan_dict(bin_e, buf_len=buf_len, downsample=ds, tune_count=3, peak_hold=bin_e % 2, crop=0.0)
    w = port_oracle.window_coefs("blackman", n)
    reads, hops = make_reads(port_oracle.lib, plan, 5, SYNTH_BIASED, seed=bin_e, param=45)
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


@pytest.mark.parametrize("passes,fir", [(1, 9), (2, 0), (3, 9), (4, 5), (5, 9), (6, 9), (7, 9), (8, 9), (9, 0)])
def test_fifth_order_chain_every_depth(scan_mod, port_oracle, passes, fir):
    """-F path at every decimation depth: fused tile kernel (P <= 7) and per-pass kernels"""
    bin_e = 7
    n, ds = 1 << bin_e, 1 << passes
    buf_len = max(16384, 2 * n * ds)
    plan = plan_dict(bin_e, buf_len=buf_len, downsample=ds, downsample_passes=passes, boxcar=0,
                     comp_fir_size=fir, tune_count=2, peak_hold=passes % 2)
    w = port_oracle.window_coefs("hamming", n)
    reads, hops = make_reads(port_oracle.lib, plan, 3, SYNTH_BIASED, seed=passes, param=30)
    reads[1, 100:300] = 255
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


def test_level_stats_soft_agc_counts(scan_mod, port_oracle):
    """optional byte statistics (what softagc() counts per buffer, librtlsdr.c:3299-3306)"""
    plan = plan_dict(10, tune_count=3)
    reads, hops = make_reads(port_oracle.lib, plan, 4, SYNTH_TONE, seed=3, param=140)  # clipped tone: 0 / 255 bytes
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=None, level_stats=True)
    try:
        for r, h in zip(reads, hops):
            g.submit(int(h), r)
        for h in range(3):
            b = reads[hops == h]
            over, high, nbytes = g.level_stats(h)
            assert over == np.count_nonzero((b == 0) | (b == 255))
            assert high == np.count_nonzero((b < 64) | (b > 191))
            assert nbytes == b.size
        g.collect(1)
        assert g.level_stats(1) == (0, 0, 0)
        assert g.level_stats(0)[2] == reads[hops == 0].size
    finally:
        g.close()
    g = scan_mod.GpuScan.from_plan(plan)
    with pytest.raises(scan_mod.ScanError):
        g.level_stats(0)   # not enabled
    g.close()


def test_cli_async_callback_mode_and_replay(scan_mod, port_oracle, tmp_path):
    """rtl_power_gpu fed through rtlsdr_read_async callbacks, and from a recorded rtl_tcp capture,
    prints the same rows as the read_sync path"""
    import os
    import struct
    import subprocess
    from rtlsdr_b200 import _build
    _build.build_host()
    exe = os.path.join(_build.HOST_BUILD, "rtl_power_gpu")
    base_env = dict(os.environ, RTLSDR_SYNTH_MODE="biased", RTLSDR_SYNTH_SEED="6", RTLSDR_SYNTH_PARAM="-15",
                    RTL_POWER_PASSES="3", RTL_POWER_TIMESTAMP="2026-01-01, 00:00:00")
    args = ["-f", "433M:435M:4k", "-w", "blackman", "-P", "-1"]
    outs = {}
    for name, extra in (("sync", {}), ("async", {"RTL_POWER_ASYNC": "1"})):
        out = tmp_path / f"{name}.csv"
        r = subprocess.run([exe] + args + [str(out)], env=dict(base_env, **extra), capture_output=True, text=True,
                           timeout=120)
        assert r.returncode == 0, r.stderr
        outs[name] = out.read_text()
    assert outs["sync"] == outs["async"] and len(outs["sync"]) > 1000
    # the same bytes recorded as an rtl_tcp capture and replayed
    from rtlsdr_b200.planner import plan_scan
    plan = plan_scan("433M:435M:4k").as_dict()
    reads, _ = make_reads(port_oracle.lib, plan, 3, SYNTH_BIASED, seed=6, param=-15)
    cap = tmp_path / "capture.rtltcp"
    cap.write_bytes(b"RTL0" + struct.pack(">II", 5, 29) + reads.tobytes())
    out = tmp_path / "replay.csv"
    env = dict(base_env, RTLSDR_SYNTH_MODE="replay", RTLSDR_SYNTH_REPLAY=str(cap))
    r = subprocess.run([exe] + args + [str(out)], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == outs["sync"]


@pytest.mark.parametrize("bin_e", [0, 6, 12, 14])
def test_collect_device_reports_and_zeroes(scan_mod, port_oracle, bin_e):
    """collect_device: dB rows, raw bins and sample counts written into caller device buffers by the
    epilogue kernel, accumulators cleared on the device (fused for N <= 8192, memset above)."""
    import torch
    n = 1 << bin_e
    plan = plan_dict(bin_e, buf_len=max(16384, 2 * n), tune_count=4, crop=0.2 if bin_e else 0.0, rate=2000000)
    w = port_oracle.window_coefs("hamming", n) if bin_e else None
    reads, hops = make_reads(port_oracle.lib, plan, 3, SYNTH_BIASED, seed=bin_e, param=9)
    want = expected(port_oracle, plan, w if w is not None else np.zeros(1, np.int32), reads, hops)
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w)
    try:
        d_avg = torch.full((4, n), -1, dtype=torch.int64, device="cuda")
        d_smp = torch.full((4,), -1, dtype=torch.int32, device="cuda")
        d_db = torch.zeros((4, g.db_count), dtype=torch.float64, device="cuda")
        for rep in range(2):   # second round: accumulators really were zeroed
            for r, h in zip(reads, hops):
                g.submit(int(h), r)
            g.collect_device(d_avg.data_ptr(), d_smp.data_ptr(), d_db.data_ptr())
            g.sync()
            assert np.array_equal(d_avg.cpu().numpy(), want[0])
            assert np.array_equal(d_smp.cpu().numpy(), want[1])
            assert db_close(d_db.cpu().numpy(), want[2])
        avg, smp, _ = g.collect_all(want_db=False)
        assert not avg.any() and not smp.any()
    finally:
        g.close()


def test_argument_errors_on_device_paths(scan_mod):
    import torch
    g = scan_mod.GpuScan(4, 10, 16384)
    try:
        dev = torch.zeros(4 * 16384 + 64, dtype=torch.uint8, device="cuda")
        for args, code in (((0, 4, 1, dev.data_ptr() + 1, 4 * 16384, 16384), -8),     # misaligned base
                           ((0, 4, 1, dev.data_ptr(), 4 * 16384 + 8, 16384), -8),     # misaligned stride
                           ((2, 3, 1, dev.data_ptr(), 4 * 16384, 16384), -3),         # hops 2..4 of 4
                           ((0, 4, 0, dev.data_ptr(), 4 * 16384, 16384), -2)):        # no passes
            with pytest.raises(scan_mod.ScanError) as e:
                g.submit_device(*args)
            assert e.value.code == code, args
        with pytest.raises(scan_mod.ScanError) as e:
            g.collect(4)
        assert e.value.code == -3
        # an untouched handle reports zeros and "-inf"-free rows are the caller's business
        avg, smp, db = g.collect_all()
        assert not avg.any() and not smp.any()
    finally:
        g.close()
    # u8 fast path needs the planner's buffer length (16384 whenever 2N <= 16384, rtl_power.c:501-504)
    with pytest.raises(scan_mod.ScanError) as e:
        scan_mod.GpuScan(1, 10, 32768)
    assert e.value.code == -2


def test_sweep_main_single_rank_rows(scan_mod, port_oracle, tmp_path):
    """rtlsdr_b200.sweep_main (the hop-sharded driver) at world size 1: rows equal the oracle's"""
    import os
    import subprocess
    import sys
    from rtlsdr_b200.planner import plan_scan
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "sweep.csv"
    r = subprocess.run([sys.executable, "-m", "rtlsdr_b200.sweep_main", "-f", "88M:108M:25k", "-c", "20%", "-w",
                        "hamming", "--sweeps", "3", "--intervals", "2", "--synth", "biased", "--seed", "4",
                        "--param", "17", "-o", str(out)], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    plan = plan_scan("88M:108M:25k", 0.2)
    pd = plan.as_dict()
    pd["peak_hold"] = 0
    w = port_oracle.window_coefs("hamming", 1 << pd["bin_e"])
    reads, hops = make_reads(port_oracle.lib, pd, 6, SYNTH_BIASED, seed=4, param=17)
    per = 3 * pd["tune_count"]
    want = ""
    for i in range(2):
        avg, smp, db = expected(port_oracle, pd, w, reads[i * per:(i + 1) * per], hops[i * per:(i + 1) * per])
        want += "".join("2026-01-01, 00:00:00, " + plan.csv_row(h, int(smp[h]), db[h]) for h in range(pd["tune_count"]))
    assert out.read_text() == want


@pytest.mark.parametrize("peak", [0, 1])
def test_mixed_submission_paths_stress(scan_mod, port_oracle, peak):
    """random interleaving of submit (ring), submit_batch (pinned), submit_device, flush, per-hop
    collect and a caller-owned stream; every hop's report must equal the oracle over exactly the
    reads submitted since its previous collect"""
    import torch
    rng = np.random.default_rng(99 + peak)
    bin_e, tc, b = 8, 5, 16384
    n = 1 << bin_e
    plan = plan_dict(bin_e, tune_count=tc, peak_hold=peak, crop=0.1, rate=2500000)
    w = port_oracle.window_coefs("blackman-harris", n)
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w, ring_bytes=8 * b)   # tiny ring: many flushes
    pinned = scan_mod.PinnedBuffer(6 * tc * b)
    pending = {h: [] for h in range(tc)}

    def check(h):
        avg, smp, db = g.collect(h)
        reads = np.array(pending[h], dtype=np.uint8).reshape(-1, b) if pending[h] else np.zeros((0, b), np.uint8)
        one = dict(plan, tune_count=1)
        want = expected(port_oracle, one, w, reads, np.zeros(len(reads), np.int32))
        assert np.array_equal(avg, want[0][0]), h
        assert smp == want[1][0]
        if len(reads):
            assert db_close(db, want[2][0])
        pending[h] = []

    try:
        user_stream = torch.cuda.Stream()
        for step in range(60):
            op = rng.integers(0, 6)
            if op == 0:      # single reads through the ring
                for _ in range(rng.integers(1, 12)):
                    h = int(rng.integers(0, tc))
                    r = rng.integers(0, 256, b, dtype=np.uint8)
                    g.submit(h, r)
                    pending[h].append(r)
            elif op == 1:    # pinned strided batch over a hop range
                h0 = int(rng.integers(0, tc)); hc = int(rng.integers(1, tc - h0 + 1)); p = int(rng.integers(1, 7))
                cube = pinned.view(np.uint8, (p, hc, b))
                cube[:] = rng.integers(0, 256, (p, hc, b), dtype=np.uint8)
                g.submit_batch(h0, hc, p, pinned.ptr, hc * b, b)
                g.sync()     # the pinned block is reused by the next batch
                for pi in range(p):
                    for k in range(hc):
                        pending[h0 + k].append(cube[pi, k].copy())
            elif op == 2:    # device-resident batch
                h0 = int(rng.integers(0, tc)); hc = int(rng.integers(1, tc - h0 + 1)); p = int(rng.integers(1, 5))
                host = rng.integers(0, 256, (p, hc, b), dtype=np.uint8)
                dev = torch.from_numpy(host).cuda()
                torch.cuda.synchronize()
                g.submit_device(h0, hc, p, dev.data_ptr(), hc * b, b)
                g.sync()
                for pi in range(p):
                    for k in range(hc):
                        pending[h0 + k].append(host[pi, k])
            elif op == 3:
                g.flush()
            elif op == 4:
                check(int(rng.integers(0, tc)))
            else:            # move the handle to a caller stream and back
                g.set_stream(user_stream.cuda_stream if step % 2 else None)
        for h in range(tc):
            check(h)
    finally:
        g.close()


@pytest.mark.parametrize("seed", list(range(24)))
def test_random_configurations_fuzz(scan_mod, port_oracle, seed):
    """random (bin_e, decimator, window, peak, crop, hop count, submission order) against the oracle"""
    rng = np.random.default_rng(1000 + seed)
    kind = rng.choice(["plain", "boxcar", "halfband", "rms", "large"], p=[0.35, 0.25, 0.2, 0.05, 0.15])
    peak = int(rng.integers(0, 2))
    tc = int(rng.integers(1, 6))
    crop = float(rng.choice([0.0, 0.1, 0.33, 0.5]))
    rate = int(rng.integers(900001, 2800000))
    window = str(rng.choice(WINDOWS))
    if kind == "plain":
        bin_e = int(rng.integers(1, 13))
        plan = plan_dict(bin_e, peak_hold=peak, tune_count=tc, crop=crop, rate=rate)
    elif kind == "boxcar":
        bin_e = int(rng.integers(1, 12))
        ds = int(rng.integers(2, 200))
        plan = plan_dict(bin_e, buf_len=max(16384, 2 * (1 << bin_e) * ds), downsample=ds, peak_hold=peak,
                         tune_count=tc, crop=crop, rate=rate)
    elif kind == "halfband":
        bin_e = int(rng.integers(2, 11))
        p = int(rng.integers(1, 10))
        plan = plan_dict(bin_e, buf_len=max(16384, 2 * (1 << bin_e) << p), downsample=1 << p, downsample_passes=p,
                         boxcar=0, comp_fir_size=int(rng.choice([0, 9])), peak_hold=peak, tune_count=tc, crop=crop,
                         rate=rate)
    elif kind == "rms":
        bin_e = 0
        plan = plan_dict(0, peak_hold=peak, tune_count=tc, crop=0.0, rate=rate)
    else:
        bin_e = int(rng.integers(13, 17))
        plan = plan_dict(bin_e, buf_len=2 << bin_e, peak_hold=peak, tune_count=min(tc, 2), crop=crop, rate=rate)
    n = 1 << plan["bin_e"]
    w = port_oracle.window_coefs(window, n) if plan["bin_e"] else np.zeros(1, np.int32)
    passes = int(rng.integers(1, 5))
    mode, param = [(SYNTH_XORSHIFT, 0), (SYNTH_BIASED, int(rng.integers(-60, 60))), (SYNTH_TONE, int(rng.integers(20, 200))),
                   (SYNTH_COUNTER, 0)][int(rng.integers(0, 4))]
    reads, hops = make_reads(port_oracle.lib, plan, passes, mode, seed=seed, param=param)
    order = rng.permutation(len(reads))            # hops arrive in any order
    reads, hops = reads[order], hops[order]
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w if plan["bin_e"] else None, reads, hops)
    assert np.array_equal(got[0], want[0]), (kind, plan)
    assert np.array_equal(got[1], want[1]), (kind, plan)
    assert db_close(got[2], want[2]), (kind, plan)


@pytest.mark.parametrize("bin_e,ds", [(8, 2), (8, 13), (9, 28), (10, 28), (10, 64), (11, 5), (12, 3), (12, 9)])
@pytest.mark.parametrize("peak", [0, 1])
def test_fused_boxcar_path(scan_mod, port_oracle, bin_e, ds, peak):
    """narrow boxcar scans with one FFT block per read go through the single fused kernel"""
    n = 1 << bin_e
    plan = plan_dict(bin_e, buf_len=2 * n * ds, downsample=ds, tune_count=3, peak_hold=peak, crop=0.2)
    w = port_oracle.window_coefs("hamming", n)
    reads, hops = make_reads(port_oracle.lib, plan, 6, SYNTH_BIASED, seed=bin_e * 100 + ds, param=-25)
    reads[4, :] = 255
    reads[7, ::2] = 0
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


@pytest.mark.parametrize("bin_e,ds", [(8, 2), (8, 13), (8, 56), (9, 28), (10, 28), (10, 64), (11, 5), (11, 24), (12, 3), (12, 16)])
@pytest.mark.parametrize("peak", [0, 1])
@pytest.mark.parametrize("stream", ["0", "1", "2", "3", "5"])
def test_boxcar_stream_kernel_forced(scan_mod, port_oracle, monkeypatch, bin_e, ds, peak, stream):
    """both narrow-scan kernels (single-role / warp-specialised producer-boxcar-transform) on the same inputs;
    many short segments so that ring slots, image buffers and barrier phases wrap several times"""
    monkeypatch.setenv("RTLSDR_GPU_BOXCAR_STREAM", stream)
    n = 1 << bin_e
    plan = plan_dict(bin_e, buf_len=2 * n * ds, downsample=ds, tune_count=2, peak_hold=peak, crop=0.1)
    w = port_oracle.window_coefs("blackman-harris", n)
    passes = 650  # 1300 reads -> ~300 segments of 4-5 reads, two per CTA of the one-CTA-per-SM kernel
    reads, hops = make_reads(port_oracle.lib, plan, passes, SYNTH_BIASED, seed=bin_e * 1000 + ds, param=31)
    reads[5, :] = 255
    reads[8, 1::2] = 0
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops, how="device")
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])


@pytest.mark.parametrize("passes,fir", [(1, 0), (2, 9), (3, 9), (4, 9), (5, 0)])
@pytest.mark.parametrize("tile_kernel", [False, True])
def test_fifth_order_streaming_big_batch(scan_mod, port_oracle, monkeypatch, passes, fir, tile_kernel):
    """-F chain on one large device batch: the register-streaming kernel picks long spans (256 final samples per
    thread with 2400 reads) + head tiles; the tile kernel alone must give the same bins"""
    if tile_kernel:
        monkeypatch.setenv("RTLSDR_GPU_NO_HB_STREAM", "1")
    bin_e = 7
    n, ds = 1 << bin_e, 1 << passes
    buf_len = max(16384, 2 * n * ds)
    plan = plan_dict(bin_e, buf_len=buf_len, downsample=ds, downsample_passes=passes, boxcar=0,
                     comp_fir_size=fir, tune_count=2, peak_hold=0, crop=0.3)
    w = port_oracle.window_coefs("bartlett", n)
    reads, hops = make_reads(port_oracle.lib, plan, 1200, SYNTH_BIASED, seed=77 + passes, param=21)
    reads[7, 100:300] = 255
    reads[11, :40] = 0
    want = expected(port_oracle, plan, w, reads, hops)
    got = run_gpu(scan_mod, plan, w, reads, hops, how="device")
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
    assert db_close(got[2], want[2])
