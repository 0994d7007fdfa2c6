"""The five BASELINE.json configurations at (or near) full size on the GPU.

Where the oracle finishes in seconds the comparison is exact over the whole
workload; for the biggest ones a subset of hops is checked against the oracle
and the rest through size-independent properties of the accumulation
(additivity over disjoint read sets, idempotent peak hold, order independence)."""
import numpy as np
import pytest

from oracles import SYNTH_BIASED, SYNTH_XORSHIFT, fnv1a_int64
from scan_cases import db_close, expected, make_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scan_mod():
    import rtlsdr_b200.scan as s
    s.load_library()
    return s


def device_scan(scan_mod, plan, window, reads, passes, hop_first=0, hop_count=None, g=None):
    """reads: uint8 [passes, tune_count, buf_len] (pass-major, like a sweep)"""
    import torch
    own = g is None
    if own:
        g = scan_mod.GpuScan.from_plan(plan, window_coefs=window)
    tc, b = plan["tune_count"], plan["buf_len"]
    hop_count = tc if hop_count is None else hop_count
    dev = torch.from_numpy(reads).cuda()
    g.submit_device(hop_first, hop_count, passes, dev.data_ptr() + hop_first * b, tc * b, b)
    g.sync()
    if not own:
        return None
    out = g.collect_all()
    g.close()
    return out


def sweep_reads(lib, plan, passes, mode, seed, param=0):
    reads, hops = make_reads(lib, plan, passes, mode, seed, param)
    return reads.reshape(passes, plan["tune_count"], plan["buf_len"]), reads, hops


def test_config1_single_hop_full(scan_mod, port_oracle):
    """rtl_power -f 100M:102.4M:2400 -i 1: 1 hop, 1024 bins, rectangle, 293 sweeps"""
    from rtlsdr_b200.planner import plan_scan
    plan = plan_scan("100M:102.4M:2400").as_dict()
    plan["peak_hold"] = 0
    w = port_oracle.window_coefs("rectangle", 1024)
    cube, reads, hops = sweep_reads(port_oracle.lib, plan, 293, SYNTH_XORSHIFT, 1)
    want = expected(port_oracle, plan, w, reads, hops)
    got = device_scan(scan_mod, plan, w, cube, 293)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and db_close(got[2], want[2])


def test_config2_fm_band_full(scan_mod, port_oracle):
    """-f 88M:108M:1k -c 20% -w hamming -i 10: 9 hops x 4096 bins x 377 sweeps = the bench workload"""
    from rtlsdr_b200.planner import plan_scan
    plan = plan_scan("88M:108M:1k", 0.2).as_dict()
    plan["peak_hold"] = 0
    w = port_oracle.window_coefs("hamming", 4096)
    cube, reads, hops = sweep_reads(port_oracle.lib, plan, 377, SYNTH_BIASED, 2, 7)
    want = expected(port_oracle, plan, w, reads, hops)
    got = device_scan(scan_mod, plan, w, cube, 377)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and db_close(got[2], want[2])
    # the host-buffer path (submit per read, ring + H2D) must give the same bins
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w, ring_bytes=4 << 20)
    for r, h in zip(reads, hops):
        g.submit(int(h), r)
    avg2, smp2, _ = g.collect_all(want_db=False)
    g.close()
    assert np.array_equal(avg2, want[0]) and np.array_equal(smp2, want[1])


def test_config3_wideband_623_hops(scan_mod, port_oracle):
    """-f 24M:1766M:1k -F 9: 623 hops x 4096 bins, FIR/boxcar inert at this width, 6 sweeps"""
    from rtlsdr_b200.planner import plan_scan
    plan = plan_scan("24M:1766M:1k", 0.0, 9).as_dict()
    assert plan["tune_count"] == 623 and plan["downsample"] == 1
    plan["peak_hold"] = 0
    w = port_oracle.window_coefs("rectangle", 4096)
    cube, reads, hops = sweep_reads(port_oracle.lib, plan, 6, SYNTH_XORSHIFT, 3)
    want = expected(port_oracle, plan, w, reads, hops)
    got = device_scan(scan_mod, plan, w, cube, 6)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and db_close(got[2], want[2])


def test_config4_large_fft_peak_hold(scan_mod, port_oracle):
    """-f 100M:102.4M:19 -w blackman-harris -P: 2^17 bins, peak hold; 48 of the 1099 sweeps
    against the oracle, plus idempotence: replaying the same reads cannot change a peak."""
    from rtlsdr_b200.planner import plan_scan
    plan = plan_scan("100M:102.4M:19").as_dict()
    assert plan["bin_e"] == 17
    plan["peak_hold"] = 1
    w = port_oracle.window_coefs("blackman-harris", 1 << 17)
    cube, reads, hops = sweep_reads(port_oracle.lib, plan, 48, SYNTH_BIASED, 4, 20)
    want = expected(port_oracle, plan, w, reads, hops)
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w)
    device_scan(scan_mod, plan, w, cube, 48, g=g)
    device_scan(scan_mod, plan, w, cube, 48, g=g)   # replay: peaks idempotent, samples double
    avg, smp, _ = g.collect_all(want_db=False)
    g.close()
    assert np.array_equal(avg, want[0])
    assert np.array_equal(smp, 2 * want[1])


def test_config5_512_hop_streams_properties(scan_mod, port_oracle):
    """-f 24M:1457.6M:700: 512 hops x 4096 bins x 64 sweeps (512 MiB).  Every 37th hop exactly
    against the oracle; all hops: additivity over two disjoint halves of the sweeps and
    independence from how hops are grouped into submissions."""
    from rtlsdr_b200.planner import plan_scan
    import torch
    plan = plan_scan("24M:1457.6M:700").as_dict()
    assert plan["tune_count"] == 512
    plan["peak_hold"] = 0
    tc, b, passes = 512, plan["buf_len"], 64
    w = port_oracle.window_coefs("rectangle", 4096)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    dev = torch.randint(0, 256, (passes, tc, b), dtype=torch.uint8, device="cuda", generator=gen)
    g = scan_mod.GpuScan.from_plan(plan, window_coefs=w)
    g.submit_device(0, tc, passes, dev.data_ptr(), tc * b, b)
    full, smp, _ = g.collect_all(want_db=False)
    assert (smp == passes * 2).all()
    # halves of the sweeps, and hops submitted in two separate groups
    g.submit_device(0, tc, passes // 2, dev.data_ptr(), tc * b, b)
    first, _, _ = g.collect_all(want_db=False)
    g.submit_device(0, 200, passes // 2, dev[passes // 2:].data_ptr(), tc * b, b)
    g.submit_device(200, 312, passes // 2, dev[passes // 2:].data_ptr() + 200 * b, tc * b, b)
    second, _, _ = g.collect_all(want_db=False)
    g.close()
    assert np.array_equal(first + second, full)
    host = dev[:, ::37].cpu().numpy()            # hops 0, 37, 74, ...
    sub = dict(plan, tune_count=host.shape[1])
    reads = host.reshape(-1, b)
    hops = np.tile(np.arange(host.shape[1], dtype=np.int32), passes)
    want, _, _ = expected(port_oracle, sub, w, reads, hops)
    assert np.array_equal(full[::37], want)
    assert fnv1a_int64(full[::37]) == fnv1a_int64(want)
