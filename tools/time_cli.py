#!/usr/bin/env python
"""Wall-clock throughput of the C drop-in path: host/_build/rtl_power_gpu (planner, sweep loop,
rtlsdr_read_sync from the synthetic source, rtlsdr_gpu_scan_submit per read, collect per hop, CSV)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "host", "_build", "rtl_power_gpu")
reports = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for mode in ("counter", "xorshift"):
    env = dict(os.environ, RTLSDR_SYNTH_MODE=mode, RTL_POWER_PASSES="377", RTL_POWER_REPORTS=str(reports),
               RTL_POWER_TIMESTAMP="2026-01-01, 00:00:00")
    t0 = time.perf_counter()
    subprocess.run([exe, "-f", "88M:108M:1k", "-c", "20%", "-w", "hamming", "/dev/null"], env=env, check=True,
                   stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    samples = reports * 377 * 9 * 8192
    print(json.dumps({"source": mode, "reports": reports, "wall_s": dt, "Msamples_per_s_incl_process_start": samples / dt / 1e6}))
