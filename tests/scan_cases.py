"""Shared helpers for the parity tests: inputs from the synthetic source, expected
values from the oracles (compiled reference when present, C restatement otherwise)."""
import ctypes

import numpy as np

from oracles import (PortOracle, RefOracle, SYNTH_BIASED, SYNTH_CONST, SYNTH_COUNTER, SYNTH_TONE,
                     SYNTH_XORSHIFT, synth_bytes)

# the survey's known-answer rows: range, crop, window, -F, -P, passes, mode, fnv
KAT_ROWS = [
    ("100M:102.4M:2400", 0.0, "rectangle", -1, 0, 1, SYNTH_COUNTER, 0x37a8a961936a7ea5),
    ("100M:102.4M:2400", 0.0, "rectangle", -1, 0, 3, SYNTH_XORSHIFT, 0x8cded15669b6ff79),
    ("88M:108M:1k", 0.2, "hamming", -1, 0, 2, SYNTH_XORSHIFT, 0x2588fc849e07d9dd),
    ("24M:1766M:1k", 0.0, "rectangle", 9, 0, 1, SYNTH_XORSHIFT, 0x074f712a23c886d1),
    ("100M:102.4M:19", 0.0, "blackman-harris", -1, 1, 3, SYNTH_XORSHIFT, 0xd6d2ca5ee525c021),
    ("24M:1457.6M:700", 0.0, "rectangle", -1, 0, 1, SYNTH_XORSHIFT, 0x7b1c7343a9686225),
    ("100M:100.1M:100", 0.0, "rectangle", -1, 0, 2, SYNTH_XORSHIFT, 0x85e7f3bafff5a40d),
    ("100M:100.1M:100", 0.0, "blackman", 9, 0, 2, SYNTH_XORSHIFT, 0x66bf8be862e26755),
    ("100M:100.1M:100", 0.0, "youssef", 0, 1, 2, SYNTH_XORSHIFT, 0x27896368d3b29815),
    ("100M:100.5M:10k", 0.0, "bartlett", -1, 0, 2, SYNTH_XORSHIFT, 0x024cf55ee35ac3e9),
    ("100M:100.5M:10k", 0.5, "hann-poisson", -1, 0, 2, SYNTH_XORSHIFT, 0x8eb7f704af145b99),
]


def make_reads(lib, plan, passes, mode, seed=0, param=0):
    """uint8 [passes * tune_count, buf_len] in sweep order (pass-major) + hop per read."""
    tc, b = plan["tune_count"], plan["buf_len"]
    reads = np.empty((passes * tc, b), dtype=np.uint8)
    hops = np.empty(passes * tc, dtype=np.int32)
    i = 0
    for p in range(passes):
        for h in range(tc):
            reads[i] = synth_bytes(lib, mode, seed, param, tc, h, p, b)
            hops[i] = h
            i += 1
    return reads, hops


def expected(port: PortOracle, plan, window, reads, hops):
    """(avg [tc, N], samples [tc], db [tc, db_count]) from the C restatement."""
    avg, smp = port.scan(plan, window, reads, hops, plan["tune_count"])
    dbs = []
    for h in range(plan["tune_count"]):
        _, db = port.epilogue(avg[h], plan["bin_e"], plan["crop"], plan["rate"], int(smp[h]))
        dbs.append(db)
    return avg, smp, np.stack(dbs)


def db_close(got, want, rel=1e-6):
    """dB agreement within `rel` relative (BASELINE.json north_star); inf/nan must match."""
    got, want = np.asarray(got), np.asarray(want)
    fin = np.isfinite(want)
    if not np.array_equal(np.isfinite(got), fin):
        return False
    if not np.array_equal(got[~fin], want[~fin], equal_nan=True):
        return False
    return bool(np.all(np.abs(got[fin] - want[fin]) <= rel * np.maximum(np.abs(want[fin]), 1e-300)))


def plan_dict(bin_e, buf_len=16384, downsample=1, downsample_passes=0, boxcar=1, comp_fir_size=0,
              peak_hold=0, rate=2400000, crop=0.0, tune_count=1):
    return dict(tune_count=tune_count, bin_e=bin_e, buf_len=buf_len, downsample=downsample,
                downsample_passes=downsample_passes, boxcar=boxcar, comp_fir_size=comp_fir_size,
                peak_hold=peak_hold, rate=rate, crop=crop)
