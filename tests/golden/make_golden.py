#!/usr/bin/env python
"""Generate tests/golden/ref_vectors.npz and kat.json FROM THE REFERENCE ITSELF.

Runs only where /root/reference exists (this container): it drives the
unmodified rtl_power.c object through oracle/_ref (oracle/ref_harness.c) and
stores small input/output vectors, so that the oracle restatement and the CUDA
path can be pinned on boxes where the reference cannot travel.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracles import (PortOracle, RefOracle, SYNTH_BIASED, SYNTH_COUNTER, SYNTH_TONE, SYNTH_XORSHIFT,  # noqa: E402
                     WINDOWS, build_oracles, synth_bytes)
from scan_cases import KAT_ROWS  # noqa: E402


def int16_vector(lib, count, seed, kind):
    """deterministic int16 test vector from the C synthetic source (not numpy's RNG)"""
    raw = synth_bytes(lib, SYNTH_XORSHIFT, seed, 0, 1, 0, 0, 2 * count)
    v = raw.view(np.int16).copy()
    if kind == 1:      # full-scale corners: exercises the int16 wrap inside fix_fft
        v = np.where(v >= 0, 32767, -32768).astype(np.int16)
    elif kind == 2:    # u8-derived range after a rectangle window
        v = (((raw[:count].astype(np.int32) - 127) * 256) & 0xFFFF).astype(np.uint16).view(np.int16)
    return v


def main():
    build_oracles()
    ref, port = RefOracle(), PortOracle()
    out = {}
    meta = {"fix_fft": [], "fifth_order": [], "generic_fir": [], "remove_dc": [], "rms": [], "scan": []}

    for m in (1, 2, 3, 4, 5, 6, 8, 10, 12):
        ref.sine_table(m)
        out[f"sine_{m}"] = ref.sine_table(m)
        for kind in (0, 1, 2):
            x = int16_vector(port.lib, 2 << m, 100 + m, kind)
            key = f"fft_m{m}_k{kind}"
            out[key + "_in"] = x
            out[key + "_out"] = ref.fix_fft(x, m)
            meta["fix_fft"].append({"key": key, "m": m})

    for i, length in enumerate((12, 13, 64, 250, 1024)):
        x = int16_vector(port.lib, length + 8, 200 + i, i % 2)
        key = f"fifth_{length}"
        out[key + "_in"] = x
        out[key + "_out"] = ref.fifth_order(x, length)
        meta["fifth_order"].append({"key": key, "length": length})
    for p in range(1, 11):
        x = int16_vector(port.lib, 96, 300 + p, p % 2)
        key = f"fir_{p}"
        out[key + "_in"] = x
        out[key + "_out"] = ref.generic_fir(x, 90, p)
        meta["generic_fir"].append({"key": key, "length": 90, "table": p})
    for i, length in enumerate((9, 64, 65, 1000)):
        x = (int16_vector(port.lib, length + 4, 400 + i, 0) // 4 + 1500).astype(np.int16)
        key = f"dc_{length}"
        out[key + "_in"] = x
        out[key + "_out"] = ref.remove_dc(x, length)
        meta["remove_dc"].append({"key": key, "length": length})
    for i, (mode, param) in enumerate(((SYNTH_XORSHIFT, 0), (SYNTH_BIASED, 60), (SYNTH_COUNTER, 0))):
        b = synth_bytes(port.lib, mode, 500 + i, param, 1, 0, 0, 16384)
        meta["rms"].append({"mode": mode, "seed": 500 + i, "param": param,
                            "sum": ref.rms_power(b, 1234, 0), "peak": ref.rms_power(b, 10**9, 1)})

    for w in WINDOWS:
        for n in (32, 1024):
            ref.configure("100M:101M:60k" if n == 32 else "100M:102.4M:2400", 0.0, w)
            assert (1 << ref.plan["bin_e"]) == n, ref.plan
            out[f"win_{w}_{n}"] = ref.window_coefs()

    # small end-to-end scans incl. the CSV rows the reference prints
    cases = [("100M:100.5M:10k", 0.5, "hann-poisson", -1, 0, 2, SYNTH_XORSHIFT, 0),
             ("100M:100.5M:10k", 0.0, "bartlett", -1, 0, 2, SYNTH_BIASED, 25),
             ("100M:101M:60k", 0.1, "hamming", -1, 1, 3, SYNTH_TONE, 120),
             ("100M:100.1M:100", 0.0, "blackman", 9, 0, 1, SYNTH_XORSHIFT, 0),
             ("100M:104M:1M", 0.0, "rectangle", -1, 0, 2, SYNTH_BIASED, 33),
             ("88M:108M:25k", 0.2, "hamming", -1, 0, 2, SYNTH_XORSHIFT, 0)]
    for i, (freq, crop, w, fir, peak, passes, mode, param) in enumerate(cases):
        plan = ref.configure(freq, crop, w, fir, peak)
        ref.source(mode, 7, param)
        ref.scan(passes)
        key = f"scan_{i}"
        out[key + "_avg"] = ref.avg()
        out[key + "_samples"] = ref.samples()
        rows = [ref.csv(h) for h in range(plan["tune_count"])]
        meta["scan"].append({"key": key, "freq": freq, "crop": crop, "window": w, "fir": fir, "peak": peak,
                             "passes": passes, "mode": mode, "seed": 7, "param": param,
                             "plan": {k: plan[k] for k in ("tune_count", "bin_e", "buf_len", "downsample",
                                                           "downsample_passes", "rate", "boxcar",
                                                           "comp_fir_size", "crop", "freqs")},
                             "csv": rows})

    # the survey's known-answer rows, re-derived here from the reference
    kat = []
    for freq, crop, w, fir, peak, passes, mode, fnv in KAT_ROWS:
        ref.configure(freq, crop, w, fir, peak)
        ref.source(mode, 0, 0)
        ref.scan(passes)
        got = ref.fnv()
        assert got == fnv, (freq, hex(got), hex(fnv))
        kat.append({"freq": freq, "crop": crop, "window": w, "fir": fir, "peak": peak, "passes": passes,
                    "mode": mode, "fnv": f"{got:016x}", "samples0": int(ref.samples()[0]),
                    "avg0_head": [int(v) for v in ref.avg()[0][:3]]})

    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as f:
        json.dump(meta, f, indent=1)
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "ref_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
