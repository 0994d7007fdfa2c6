"""Host side of the drop-in (C, host/): planner, CSV formatter, synthetic source."""
import ctypes
import os
import subprocess
import threading

import numpy as np
import pytest

from oracles import SYNTH_COUNTER, SYNTH_XORSHIFT, synth_bytes
from rtlsdr_b200.planner import host_library, plan_scan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RANGES = [("100M:102.4M:2400", 0.0, None), ("88M:108M:1k", 0.2, None), ("24M:1766M:1k", 0.0, 9),
          ("100M:102.4M:19", 0.0, None), ("24M:1457.6M:700", 0.0, None), ("100M:100.1M:100", 0.0, None),
          ("100M:100.1M:100", 0.0, 9), ("100M:100.5M:10k", 0.5, None), ("100M:110M:1M", 0.3, None),
          ("100M:100.01M:50", 0.0, 5), ("50M:60M:25k", 0.1, 0), ("433M:434M:500", 0.25, None),
          ("0.5G:0.6G:30k", 0.0, None)]


def test_baseline_config_plans():
    """the planner facts SURVEY.md 8(a) lists for the five BASELINE configs"""
    p = plan_scan("100M:102.4M:2400")
    assert (p.tune_count, p.bin_e, p.buf_len, p.downsample, p.rate) == (1, 10, 16384, 1, 2400000)
    p = plan_scan("88M:108M:1k", 0.2)
    assert (p.tune_count, p.bin_e, p.buf_len, p.rate) == (9, 12, 16384, 2777777)
    p = plan_scan("24M:1766M:1k", 0.0, 9)
    assert (p.tune_count, p.bin_e, p.downsample, p.downsample_passes, p.rate) == (623, 12, 1, 0, 2796147)
    p = plan_scan("100M:102.4M:19")
    assert (p.tune_count, p.bin_e, p.buf_len) == (1, 17, 262144)
    p = plan_scan("24M:1457.6M:700")
    assert (p.tune_count, p.bin_e) == (512, 12)
    p = plan_scan("100M:100.1M:100")
    assert (p.bin_e, p.downsample, p.buf_len) == (10, 28, 57344)
    p = plan_scan("100M:100.1M:100", 0.0, 9)
    assert (p.downsample, p.downsample_passes, p.buf_len, p.rate) == (16, 4, 32768, 1600000)


@pytest.mark.parametrize("freq,crop,fir", RANGES)
def test_planner_matches_reference(ref_oracle, freq, crop, fir):
    p = plan_scan(freq, crop, fir).as_dict()
    r = ref_oracle.configure(freq, crop, "rectangle", -1 if fir is None else fir, 0)
    for k in ("tune_count", "bin_e", "buf_len", "downsample", "downsample_passes", "rate", "crop", "boxcar",
              "comp_fir_size", "freqs"):
        assert p[k] == r[k], (k, p[k], r[k])


def test_suffix_parsers():
    L = host_library()
    assert L.rp_atofs(b"2.4M") == 2.4e6 and L.rp_atofs(b"1k") == 1e3 and L.rp_atofs(b"1.766G") == 1.766e9
    assert L.rp_atofs(b"2400") == 2400.0 and L.rp_atofs(b"88m ") == 88e6
    assert L.rp_atoft(b"10") == 10.0 and L.rp_atoft(b"5m") == 300.0 and L.rp_atoft(b"2h") == 7200.0
    assert abs(L.rp_atofp(b"20%") - 0.2) < 1e-15 and L.rp_atofp(b"0.5") == 0.5


def test_bad_ranges():
    with pytest.raises(ValueError):
        plan_scan("100M:102M")
    with pytest.raises(ValueError):
        plan_scan("0:5000G:1k")  # needs more than 1500 hops


def test_synth_source_sync_stream():
    """read_sync semantics: settle dump after a retune, pure function of (hop, pass)."""
    L = host_library()
    L.rtlsdr_open.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32]
    L.synth_configure.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
    L.synth_set_hops.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.synth_set_block_len.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.rtlsdr_set_center_freq.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    L.rtlsdr_get_center_freq.argtypes = [ctypes.c_void_p]
    L.rtlsdr_get_center_freq.restype = ctypes.c_uint32
    L.rtlsdr_read_sync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    dev = ctypes.c_void_p()
    assert L.rtlsdr_open(ctypes.byref(dev), 0) == 0
    assert L.rtlsdr_open(ctypes.byref(ctypes.c_void_p()), 3) == -1
    freqs = (ctypes.c_int * 3)(100, 200, 300)
    L.synth_set_hops(dev, freqs, 3)
    L.synth_set_block_len(dev, 16384)
    L.synth_configure(dev, SYNTH_XORSHIFT, 5, 0)
    buf = np.zeros(16384, np.uint8)
    n = ctypes.c_int()
    for p in range(2):
        for h, f in enumerate((100, 200, 300)):
            L.rtlsdr_set_center_freq(dev, f)
            assert L.rtlsdr_get_center_freq(dev) == f
            dump = np.zeros(4096, np.uint8)
            L.rtlsdr_read_sync(dev, dump.ctypes.data, 4096, ctypes.byref(n))
            assert n.value == 4096 and (dump == 0x7F).all()
            L.rtlsdr_read_sync(dev, buf.ctypes.data, 16384, ctypes.byref(n))
            assert n.value == 16384
            assert np.array_equal(buf, synth_bytes(L, SYNTH_XORSHIFT, 5, 0, 3, h, p, 16384))
    assert L.rtlsdr_read_sync(None, None, 10, None) == -1


def test_synth_source_async_callback_contract():
    """read_async: callback on the calling thread, buffers re-armed, cancel from the callback."""
    L = host_library()
    CB = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_ubyte), ctypes.c_uint32, ctypes.c_void_p)
    L.rtlsdr_read_async.argtypes = [ctypes.c_void_p, CB, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]
    L.rtlsdr_cancel_async.argtypes = [ctypes.c_void_p]
    L.rtlsdr_open.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32]
    L.synth_configure.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
    L.synth_set_block_len.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.synth_set_hops.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    dev = ctypes.c_void_p()
    L.rtlsdr_open(ctypes.byref(dev), 0)
    L.synth_set_hops(dev, (ctypes.c_int * 1)(100), 1)
    L.synth_set_block_len(dev, 16384)
    L.synth_configure(dev, SYNTH_COUNTER, 0, 0)
    got, tids, addrs = [], set(), []

    def cb(buf, length, ctx):
        tids.add(threading.get_ident())
        addrs.append(ctypes.addressof(buf.contents))
        got.append(bytes(buf[:8]) + bytes([length >> 8]))
        if len(got) == 7:
            assert L.rtlsdr_cancel_async(dev) == 0
    assert L.rtlsdr_cancel_async(dev) == -2          # not streaming
    assert L.rtlsdr_read_async(None, CB(cb), None, 0, 0) == -1
    rc = L.rtlsdr_read_async(dev, CB(cb), None, 3, 1000)  # 1000 is not a multiple of 512 -> default 32768
    assert rc == 0 and len(got) == 7
    assert tids == {threading.get_ident()}
    assert all(g[8] == (32768 >> 8) for g in got)
    assert addrs[0] == addrs[3] == addrs[6] and addrs[0] != addrs[1]   # ring of 3 buffers, re-armed
    assert got[0][:8] == bytes(range(8))                                 # counter pattern restarts per 16 KiB block


def test_cli_builds():
    from rtlsdr_b200 import _build
    _build.build_cuda()
    _build.build_host(force=False)
    exe = os.path.join(ROOT, "host", "_build", "rtl_power_gpu")
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "No frequency range provided." in r.stderr
