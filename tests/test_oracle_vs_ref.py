"""The C restatement against the compiled UNMODIFIED reference (oracle/_ref).
Skipped where neither /root/reference nor a prebuilt oracle/_ref exists."""
import numpy as np
import pytest

from oracles import (SYNTH_BIASED, SYNTH_CONST, SYNTH_COUNTER, SYNTH_TONE, SYNTH_XORSHIFT, WINDOWS,
                     fnv1a_int64)
from scan_cases import KAT_ROWS, expected, make_reads


def test_tables(ref_oracle, port_oracle):
    for m in range(1, 19):
        assert np.array_equal(ref_oracle.sine_table(m), port_oracle.sine_table(m))
    for w in WINDOWS + ["no-such-window"]:
        for f in ("100M:102.4M:2400", "100M:100.5M:10k", "88M:108M:1k"):
            pl = ref_oracle.configure(f, 0, w)
            assert np.array_equal(ref_oracle.window_coefs(), port_oracle.window_coefs(w, 1 << pl["bin_e"]))


def test_fix_fft_incl_int16_wrap(ref_oracle, port_oracle):
    """fix_fft of the restatement == the reference object's, m = 1..17 (rtl_power.c:483 allows up to 21; the large
    sizes are what the GPU's multi-round path is checked against), incl. full-scale inputs that hit the int16 wrap"""
    rng = np.random.default_rng(1)
    for m in range(1, 18):
        ref_oracle.sine_table(m)
        n = 1 << m
        ph = 2 * np.pi * max(n // 8, 1) * np.arange(n) / n + np.pi / 4
        corner = np.empty(2 * n, np.int16)
        corner[0::2] = np.clip(np.round(46340 * np.cos(ph)), -32768, 32767)
        corner[1::2] = np.clip(np.round(46340 * np.sin(ph)), -32768, 32767)
        for x in (rng.integers(-32768, 32768, 2 * n).astype(np.int16), np.full(2 * n, -32768, np.int16),
                  (rng.integers(0, 2, 2 * n) * 65535 - 32768).astype(np.int16), corner):
            assert np.array_equal(ref_oracle.fix_fft(x, m), port_oracle.fix_fft(x, m)), m


def test_fix_fft_largest_sizes(ref_oracle, port_oracle):
    """m = 18..21, the rest of what rtl_power.c:483 allows (the GPU's rounds B and C for N > 2^17 are checked against
    the restatement at these sizes): a random and an alternating full-scale input each"""
    rng = np.random.default_rng(3)
    for m in range(18, 22):
        ref_oracle.sine_table(m)
        n = 1 << m
        for x in (rng.integers(-32768, 32768, 2 * n).astype(np.int16),
                  (rng.integers(0, 2, 2 * n) * 65535 - 32768).astype(np.int16)):
            assert np.array_equal(ref_oracle.fix_fft(x, m), port_oracle.fix_fft(x, m)), m


def test_filters(ref_oracle, port_oracle):
    rng = np.random.default_rng(2)
    for length in (12, 13, 16, 64, 1000, 16383, 16384):
        x = rng.integers(-32768, 32768, length + 16).astype(np.int16)
        assert np.array_equal(ref_oracle.fifth_order(x, length), port_oracle.fifth_order(x, length))
        assert np.array_equal(ref_oracle.remove_dc(x, length), port_oracle.remove_dc(x, length))
        if length >= 20:
            for p in range(1, 11):
                assert np.array_equal(ref_oracle.generic_fir(x, length, p), port_oracle.generic_fir(x, length, p))
    for t in range(8):
        b = np.clip(rng.integers(0, 256, 16384) + (t - 4) * 20, 0, 255).astype(np.uint8)
        for pk in (0, 1):
            assert ref_oracle.rms_power(b, 777, pk) == port_oracle.rms_power(b, 777, pk)


@pytest.mark.parametrize("row", KAT_ROWS, ids=[f"{r[0]}-{r[2]}-F{r[3]}-P{r[4]}" for r in KAT_ROWS])
def test_reference_reproduces_survey_rows(ref_oracle, row):
    freq, crop, window, fir, peak, passes, mode, fnv = row
    ref_oracle.configure(freq, crop, window, fir, peak)
    ref_oracle.source(mode, 0, 0)
    ref_oracle.scan(passes)
    assert ref_oracle.fnv() == fnv
    assert fnv1a_int64(ref_oracle.avg()) == fnv


@pytest.mark.parametrize("freq,bin_e,window,peak", [("100M:102.4M:150", 15, "blackman-harris", 0),
                                                    ("100M:102.4M:38", 17, "hamming", 1),
                                                    ("100M:102.4M:600", 13, "bartlett", 0)])
def test_whole_scans_large_bin_counts(ref_oracle, port_oracle, freq, bin_e, window, peak):
    """whole-scan restatement vs reference at 2^13 / 2^15 / 2^17 bins (the port's large sizes rested on
    a single known-answer hash)"""
    for mode, param in ((SYNTH_BIASED, -31), (SYNTH_TONE, 127), (SYNTH_CONST, 255)):
        pl = ref_oracle.configure(freq, 0.1, window, -1, peak)
        assert pl["bin_e"] == bin_e
        ref_oracle.source(mode, 5, param)
        ref_oracle.scan(2)
        reads, hops = make_reads(port_oracle.lib, pl, 2, mode, 5, param)
        avg, smp, db = expected(port_oracle, pl, ref_oracle.window_coefs(), reads, hops)
        assert np.array_equal(avg, ref_oracle.avg()), (freq, mode)
        assert np.array_equal(smp, ref_oracle.samples())
        from rtlsdr_b200.planner import plan_scan
        p = plan_scan(freq, 0.1)
        assert p.csv_row(0, int(smp[0]), db[0]) == ref_oracle.csv(0)


@pytest.mark.parametrize("freq,crop,window,fir,peak", [
    ("100M:102.4M:2400", 0.0, "rectangle", -1, 0), ("88M:108M:1k", 0.2, "hamming", -1, 0),
    ("100M:100.1M:100", 0.0, "rectangle", -1, 0), ("100M:100.1M:100", 0.0, "blackman", 9, 0),
    ("100M:100.1M:100", 0.0, "youssef", 0, 1), ("100M:100.5M:10k", 0.0, "bartlett", -1, 0),
    ("100M:100.5M:10k", 0.5, "hann-poisson", -1, 0), ("100M:110M:1M", 0.0, "rectangle", -1, 1),
    ("100M:100.3M:3k", 0.0, "hamming", -1, 0), ("100M:100.01M:50", 0.0, "hamming", 9, 1),
    ("100M:100.9M:120", 0.0, "kaiser", -1, 0)])
def test_whole_scans(ref_oracle, port_oracle, freq, crop, window, fir, peak):
    for mode, param in ((SYNTH_XORSHIFT, 0), (SYNTH_BIASED, 9), (SYNTH_CONST, 255), (SYNTH_CONST, 0),
                        (SYNTH_TONE, 120), (SYNTH_COUNTER, 0)):
        pl = ref_oracle.configure(freq, crop, window, fir, peak)
        ref_oracle.source(mode, 7, param)
        ref_oracle.scan(2)
        reads, hops = make_reads(port_oracle.lib, pl, 2, mode, 7, param)
        avg, smp, db = expected(port_oracle, pl, ref_oracle.window_coefs(), reads, hops)
        assert np.array_equal(avg, ref_oracle.avg()), (freq, mode)
        assert np.array_equal(smp, ref_oracle.samples())
        # text rows: same bytes as csv_dbm prints (also zeroes the reference's accumulators)
        from rtlsdr_b200.planner import plan_scan
        p = plan_scan(freq, crop, None if fir < 0 else fir)
        for h in range(min(pl["tune_count"], 3)):
            assert p.csv_row(h, int(smp[h]), db[h]) == ref_oracle.csv(h)
