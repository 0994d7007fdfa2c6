"""ctypes front ends for the two parity oracles (TEST INFRASTRUCTURE).

* ``RefOracle``  -- oracle/_ref/librtlpower_ref.so: the UNMODIFIED reference
  rtl_power.c object driven like its own main() (oracle/ref_harness.c).  Built
  here from /root/reference; on the GPU box only the prebuilt .so exists.
* ``PortOracle`` -- oracle/_build/liboracle.so: oracle/scan_oracle.c, our C
  restatement, rebuildable anywhere with gcc.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "librtlpower_ref.so")
PORT_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

SYNTH_XORSHIFT, SYNTH_COUNTER, SYNTH_CONST, SYNTH_BIASED, SYNTH_TONE, SYNTH_REPLAY = range(6)

WINDOWS = ["rectangle", "hamming", "blackman", "blackman-harris", "hann-poisson",
           "youssef", "kaiser", "bartlett"]


def build_oracles(quiet=True):
    """make -C oracle (port always; ref only where /root/reference exists)."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, stdout=out, stderr=out)
    if os.path.exists("/root/reference/src/rtl_power.c"):
        subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, stdout=out, stderr=out)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def synth_bytes(lib, mode, seed, param, tune_count, hop, pass_idx, length):
    out = np.empty(length, dtype=np.uint8)
    lib.synth_generate(mode, ctypes.c_uint64(seed), param, tune_count, hop,
                       ctypes.c_uint64(pass_idx), _ptr(out), ctypes.c_size_t(length))
    return out


class OracleCfg(ctypes.Structure):
    _fields_ = [("bin_e", ctypes.c_int), ("buf_len", ctypes.c_int),
                ("downsample", ctypes.c_int), ("downsample_passes", ctypes.c_int),
                ("boxcar", ctypes.c_int), ("comp_fir_size", ctypes.c_int),
                ("peak_hold", ctypes.c_int),
                ("window", ctypes.c_void_p), ("sine", ctypes.c_void_p)]


class PortOracle:
    """oracle/scan_oracle.c"""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build_oracles()
        L = self.lib = ctypes.CDLL(PORT_SO)
        L.oracle_fix_mpy.restype = ctypes.c_int16
        L.oracle_fix_mpy.argtypes = [ctypes.c_int16, ctypes.c_int16]
        L.oracle_rms_power.restype = ctypes.c_int64
        L.oracle_rms_power.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int]
        L.oracle_cic9.restype = ctypes.POINTER(ctypes.c_int * 10)
        L.oracle_epilogue.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def sine_table(self, m):
        out = np.zeros(max((1 << m) * 3 // 4, 1), dtype=np.int16)
        self.lib.oracle_sine_table(m, _ptr(out))
        return out[: (1 << m) * 3 // 4]

    def window_coefs(self, name, n):
        out = np.zeros(n, dtype=np.int32)
        self.lib.oracle_window_coefs(name.encode(), n, _ptr(out))
        return out

    def fix_fft(self, iq, m, sine=None, log2_nwave=None):
        iq = np.ascontiguousarray(iq, dtype=np.int16).copy()
        if log2_nwave is None:
            log2_nwave = m
        if sine is None:
            sine = self.sine_table(log2_nwave)
        sine = np.ascontiguousarray(np.concatenate([sine, np.zeros(4, np.int16)]))
        rc = self.lib.oracle_fix_fft(_ptr(iq), m, _ptr(sine), log2_nwave)
        assert rc == 0
        return iq

    def fifth_order(self, data, length):
        d = np.ascontiguousarray(data, dtype=np.int16).copy()
        self.lib.oracle_fifth_order(_ptr(d), length)
        return d

    def generic_fir(self, data, length, passes):
        d = np.ascontiguousarray(data, dtype=np.int16).copy()
        self.lib.oracle_generic_fir(_ptr(d), length, self.lib.oracle_cic9(passes))
        return d

    def remove_dc(self, data, length):
        d = np.ascontiguousarray(data, dtype=np.int16).copy()
        self.lib.oracle_remove_dc(_ptr(d), length)
        return d

    def rms_power(self, buf, avg0=0, peak=0):
        b = np.ascontiguousarray(buf, dtype=np.uint8)
        return int(self.lib.oracle_rms_power(_ptr(b), len(b), avg0, peak))

    def scan(self, plan, window, reads, hops, tune_count=None, sine=None):
        """Accumulate `reads` (array [n_reads, buf_len] of uint8) into per-hop
        spectra; hops[i] is the hop index of read i.  Returns (avg, samples)."""
        n = 1 << plan["bin_e"]
        if tune_count is None:
            tune_count = int(max(hops)) + 1 if len(hops) else 1
        if sine is None:
            sine = self.sine_table(plan["bin_e"])
        sine = np.ascontiguousarray(np.concatenate([sine, np.zeros(4, np.int16)]))
        window = np.ascontiguousarray(window, dtype=np.int32)
        cfg = OracleCfg(plan["bin_e"], plan["buf_len"], plan["downsample"],
                        plan["downsample_passes"], plan["boxcar"], plan["comp_fir_size"],
                        plan["peak_hold"], window.ctypes.data, sine.ctypes.data)
        avg = np.zeros((tune_count, n), dtype=np.int64)
        samples = np.zeros(tune_count, dtype=np.int32)
        work = np.zeros(plan["buf_len"] + 16, dtype=np.int16)
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        for i, h in enumerate(hops):
            s = ctypes.c_int(int(samples[h]))
            self.lib.oracle_scan_read(ctypes.byref(cfg), _ptr(reads[i]), _ptr(work),
                                      _ptr(avg[h]), ctypes.byref(s))
            samples[h] = s.value
        return avg, samples

    def epilogue(self, avg, bin_e, crop, rate, samples):
        a = np.ascontiguousarray(avg, dtype=np.int64).copy()
        db = np.zeros((1 << bin_e) + 2, dtype=np.float64)
        k = self.lib.oracle_epilogue(_ptr(a), bin_e, crop, rate, samples, _ptr(db))
        return a, db[:k]


class RefOracle:
    """The unmodified reference object (one configuration at a time, global state)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            build_oracles()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = self.lib = ctypes.CDLL(REF_SO)
        L.ref_configure.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.c_char_p,
                                    ctypes.c_int, ctypes.c_int]
        L.ref_source.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
        L.ref_source_replay.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
        L.ref_fnv.restype = ctypes.c_uint64
        L.ref_crop.restype = ctypes.c_double
        L.ref_scan_timed.restype = ctypes.c_double
        L.ref_rms_power.restype = ctypes.c_long
        L.ref_rms_power.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_int]
        L.ref_real_conj.restype = ctypes.c_long
        L.ref_real_conj.argtypes = [ctypes.c_int16, ctypes.c_int16]
        L.ref_window.restype = ctypes.c_double
        L.ref_window.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        self.plan = None
        self._keep = None

    @staticmethod
    def available():
        return os.path.exists(REF_SO) or os.path.exists("/root/reference/src/rtl_power.c")

    def configure(self, freq, crop=0.0, window="rectangle", fir=-1, peak=0):
        n = self.lib.ref_configure(freq.encode(), crop, window.encode(), fir, peak)
        p = (ctypes.c_int * 8)()
        self.lib.ref_plan(p)
        self.plan = dict(tune_count=p[0], bin_e=p[1], buf_len=p[2], downsample=p[3],
                         downsample_passes=p[4], rate=p[5], boxcar=p[6], comp_fir_size=p[7],
                         peak_hold=1 if peak else 0, crop=float(self.lib.ref_crop()),
                         freqs=[self.lib.ref_hop_freq(i) for i in range(n)],
                         window=window)
        return self.plan

    def source(self, mode, seed=0, param=0):
        self.lib.ref_source(mode, seed, param)

    def source_replay(self, pool):
        """pool: uint8 [n_reads, buf_len]; read r = pass * tune_count + hop."""
        pool = np.ascontiguousarray(pool, dtype=np.uint8)
        self._keep = pool
        self.lib.ref_source_replay(_ptr(pool), pool.shape[1], pool.shape[0])

    def scan(self, passes):
        self.lib.ref_scan(passes)

    def scan_timed(self, passes):
        return float(self.lib.ref_scan_timed(passes))

    def avg(self):
        n = 1 << self.plan["bin_e"]
        out = np.zeros((self.plan["tune_count"], n), dtype=np.int64)
        for h in range(self.plan["tune_count"]):
            self.lib.ref_avg(h, _ptr(out[h]))
        return out

    def samples(self):
        return np.array([self.lib.ref_samples(h) for h in range(self.plan["tune_count"])],
                        dtype=np.int32)

    def fnv(self):
        return int(self.lib.ref_fnv())

    def window_coefs(self):
        out = np.zeros(1 << self.plan["bin_e"], dtype=np.int32)
        self.lib.ref_window_coefs(_ptr(out))
        return out

    def sinewave(self):
        n = (1 << self.plan["bin_e"]) * 3 // 4
        out = np.zeros(max(n, 1), dtype=np.int16)
        self.lib.ref_sinewave(_ptr(out))
        return out[:n]

    def csv(self, hop):
        cap = (1 << self.plan["bin_e"]) * 16 + 256
        buf = ctypes.create_string_buffer(cap)
        n = self.lib.ref_csv(hop, buf, cap)
        assert 0 <= n < cap
        return buf.value.decode()

    # unit-level
    def sine_table(self, m):
        self.lib.ref_sine_table(m)
        self.plan = dict(bin_e=m)
        n = (1 << m) * 3 // 4
        out = np.zeros(max(n, 1), dtype=np.int16)
        self.lib.ref_sinewave(_ptr(out))
        return out[:n]

    def fix_fft(self, iq, m):
        """requires sine_table(m) or configure() beforehand"""
        iq = np.ascontiguousarray(iq, dtype=np.int16).copy()
        rc = self.lib.ref_fix_fft(_ptr(iq), m)
        assert rc == 0
        return iq

    def fifth_order(self, data, length):
        d = np.ascontiguousarray(data, dtype=np.int16).copy()
        self.lib.ref_fifth_order(_ptr(d), length)
        return d

    def generic_fir(self, data, length, passes):
        d = np.ascontiguousarray(data, dtype=np.int16).copy()
        self.lib.ref_generic_fir(_ptr(d), length, passes)
        return d

    def remove_dc(self, data, length):
        d = np.ascontiguousarray(data, dtype=np.int16).copy()
        self.lib.ref_remove_dc(_ptr(d), length)
        return d

    def rms_power(self, buf, avg0=0, peak=0):
        b = np.ascontiguousarray(buf, dtype=np.uint8)
        return int(self.lib.ref_rms_power(_ptr(b), len(b), avg0, peak))


def fnv1a_int64(avg):
    """FNV-1a over int64 words in natural order."""
    h = 14695981039346656037
    for v in np.ascontiguousarray(avg, dtype=np.int64).ravel().view(np.uint64):
        h ^= int(v)
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h
