"""The C-ABI library builds for sm_100a, loads, and exports exactly what
include/rtlsdr_gpu_scan.h declares.  No compute calls: there is no GPU here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtlsdr_gpu_scan.h")


@pytest.fixture(scope="module")
def lib():
    from rtlsdr_b200 import _build
    import rtlsdr_b200.scan as rs
    _build.build_cuda()
    return rs.load_library()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"RTLSDR_GPU_API[^;(]*?\b(rtlsdr_gpu_scan_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    # and nothing else leaks out of the library (hidden default visibility, like CMakeLists.txt:54)
    from rtlsdr_b200.scan import lib_path
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(names), exported ^ set(names)


def test_library_contains_sm100a_code_only():
    from rtlsdr_b200.scan import lib_path
    out = subprocess.run(["cuobjdump", "-lelf", lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_null_and_config_errors_without_gpu(lib):
    h = ctypes.c_void_p()
    assert lib.rtlsdr_gpu_scan_init(None, ctypes.byref(h)) == -1
    assert lib.rtlsdr_gpu_scan_submit(None, 0, None, 0) == -1
    assert lib.rtlsdr_gpu_scan_collect(None, 0, None, None, None) == -1
    assert lib.rtlsdr_gpu_scan_flush(None) == -1
    assert lib.rtlsdr_gpu_scan_db_count(None) == -1
    assert lib.rtlsdr_gpu_scan_strerror(-3) == b"hop index out of range"
    lib.rtlsdr_gpu_scan_close(None)  # no-op like free(NULL)


def test_no_cpu_fallback(lib):
    """Without a usable sm_100 device init must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import rtlsdr_b200.scan as rs
    with pytest.raises(rs.ScanError) as e:
        rs.GpuScan(1, 10, 16384)
    assert e.value.code in (-5, -6)


def test_host_table_helpers_match_oracle(lib, port_oracle):
    import rtlsdr_b200.scan as rs
    from oracles import WINDOWS
    for m in (1, 2, 5, 10, 13):
        assert np.array_equal(rs.sine_table(m), port_oracle.sine_table(m))
    for w in WINDOWS:
        for n in (2, 64, 4096):
            assert np.array_equal(rs.window_coefs(w, n), port_oracle.window_coefs(w, n)), (w, n)
    out = np.zeros(8, dtype=np.int32)
    assert lib.rtlsdr_gpu_scan_window(b"nonsense", 8, out.ctypes.data_as(ctypes.c_void_p)) == -1
    assert (out == 256).all()  # unknown names silently stay rectangle (rtl_power.c:826-843)
