"""Recorded-IQ front doors: raw rtl_sdr files, rtl_sdr WAV files and captured
rtl_tcp streams all yield the same payload bytes (host/replay_file.c).  CPU only."""
import ctypes
import struct

import numpy as np

from rtlsdr_b200.planner import host_library


class ReplayInfo(ctypes.Structure):
    _fields_ = [("format", ctypes.c_int), ("payload_offset", ctypes.c_uint64), ("payload_bytes", ctypes.c_uint64),
                ("sample_rate", ctypes.c_uint32), ("tuner_type", ctypes.c_uint32), ("gain_count", ctypes.c_uint32)]


def wav_bytes(payload, rate, data_size=None):
    """header layout of src/convenience/wavewrite.c:120-246 (8-bit stereo PCM + an 'auxi' chunk)"""
    fmt = struct.pack("<4sIHHIIHH", b"fmt ", 16, 1, 2, rate, rate * 2, 2, 8)
    aux = b"auxi" + struct.pack("<I", 20) + bytes(20)
    data = b"data" + struct.pack("<I", len(payload) if data_size is None else data_size) + payload
    body = b"WAVE" + fmt + aux + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def test_three_formats_same_payload(tmp_path):
    L = host_library()
    L.replay_probe.argtypes = [ctypes.c_char_p, ctypes.POINTER(ReplayInfo)]
    L.replay_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ReplayInfo)]
    L.replay_load.restype = ctypes.c_void_p
    rng = np.random.default_rng(8)
    payload = rng.integers(0, 256, 3 * 16384 + 1000, dtype=np.uint8).tobytes()
    files = {
        "raw": payload,
        "wav": wav_bytes(payload, 2400000),
        "wav_unfinished": wav_bytes(payload, 2400000, data_size=0),   # recorder killed: dataSize still 0
        "rtl_tcp": b"RTL0" + struct.pack(">II", 5, 29) + payload,
    }
    want_fmt = {"raw": 0, "wav": 1, "wav_unfinished": 1, "rtl_tcp": 2}
    for name, blob in files.items():
        p = tmp_path / name
        p.write_bytes(blob)
        info = ReplayInfo()
        assert L.replay_probe(str(p).encode(), ctypes.byref(info)) == 0
        assert info.format == want_fmt[name]
        assert info.payload_bytes == len(payload), name
        n = ctypes.c_size_t()
        ptr = L.replay_load(str(p).encode(), 16384, ctypes.byref(n), ctypes.byref(info))
        assert ptr and n.value == 3
        got = ctypes.string_at(ptr, 3 * 16384)
        assert got == payload[: 3 * 16384], name
        if name.startswith("wav"):
            assert info.sample_rate == 2400000
        if name == "rtl_tcp":
            assert (info.tuner_type, info.gain_count) == (5, 29)
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFF" + struct.pack("<I", 100) + b"WAVE" + b"junk" + struct.pack("<I", 4000))
    assert L.replay_probe(str(bad).encode(), ctypes.byref(ReplayInfo())) == -2
    assert L.replay_probe(b"/nonexistent/file", ctypes.byref(ReplayInfo())) == -1
