#!/usr/bin/env python
"""Host-to-device bandwidth of every GPU of the box, alone and together (is the
end-to-end path limited by shared PCIe uplinks?).

    python tools/pcie_topo.py [--mib 1024] [--reps 6] > gpurun_out/pcie_topo.json

One process, one pinned buffer and one copy stream per GPU (allocated after binding the allocating thread to
the GPU's NUMA node when the box reports one).  Measures, with CUDA events:
  alone[g]      H2D GB/s of GPU g while every other GPU idles
  pairs[g][h]   per-GPU H2D GB/s while g and h copy at the same time (a pair behind one switch uplink
                shows about half of `alone`)
  all           per-GPU H2D GB/s with every GPU copying
  d2h_alone[g]  the other direction
plus `nvidia-smi topo -m` and the NUMA node / PCI bus id of every GPU.
"""
import argparse
import json
import os
import subprocess
import sys


def numa_of(bus):
    p = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
    try:
        return int(open(p).read().strip())
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=6)
    args = ap.parse_args()
    import torch
    n = torch.cuda.device_count()
    nbytes = args.mib << 20
    info = []
    try:
        import pynvml
        pynvml.nvmlInit()
        for g in range(n):
            h = pynvml.nvmlDeviceGetHandleByIndex(g)
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            info.append({"gpu": g, "bus": bus, "numa_node": numa_of(bus)})
    except Exception as exc:  # noqa: BLE001
        info = [{"error": repr(exc)}]
    all_cpus = os.sched_getaffinity(0)
    host, dev, streams = [], [], []
    for g in range(n):
        node = info[g].get("numa_node") if g < len(info) else None
        if node is not None and node >= 0:
            try:
                cpus = []
                for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus += list(range(int(a), int(b or a) + 1))
                os.sched_setaffinity(0, set(cpus) & all_cpus or all_cpus)
            except Exception:  # noqa: BLE001
                pass
        t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        t.fill_(g + 1)          # first touch on the bound node
        host.append(t)
        with torch.cuda.device(g):
            dev.append(torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{g}"))
            streams.append(torch.cuda.Stream(device=g))
    os.sched_setaffinity(0, all_cpus)

    def run(gpus, d2h=False):
        """per-GPU GB/s with all of `gpus` copying concurrently"""
        ev = {}
        for g in gpus:
            with torch.cuda.device(g):
                ev[g] = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for g in gpus:          # warm-up
            with torch.cuda.device(g), torch.cuda.stream(streams[g]):
                (host[g] if d2h else dev[g]).copy_(dev[g] if d2h else host[g], non_blocking=True)
        for g in gpus:
            streams[g].synchronize()
        for g in gpus:
            with torch.cuda.device(g), torch.cuda.stream(streams[g]):
                ev[g][0].record(streams[g])
                for _ in range(args.reps):
                    (host[g] if d2h else dev[g]).copy_(dev[g] if d2h else host[g], non_blocking=True)
                ev[g][1].record(streams[g])
        out = {}
        for g in gpus:
            streams[g].synchronize()
            out[g] = round(args.reps * nbytes / (ev[g][0].elapsed_time(ev[g][1]) * 1e-3) / 1e9, 2)
        return out

    res = {"gpus": n, "mib_per_copy": args.mib, "reps": args.reps, "devices": info,
           "cpus_visible": len(all_cpus)}
    res["alone"] = {g: run([g])[g] for g in range(n)}
    res["d2h_alone"] = {g: run([g], d2h=True)[g] for g in range(n)}
    res["pairs"] = {f"{g},{h}": run([g, h]) for g in range(n) for h in range(g + 1, n)}
    for k in (2, 4, 8):
        if k <= n:
            r = run(list(range(k)))
            res[f"first_{k}_together"] = {"per_gpu": r, "sum": round(sum(r.values()), 1)}
    try:
        res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception as exc:  # noqa: BLE001
        res["topo"] = repr(exc)
    try:
        res["lscpu_numa"] = [line for line in subprocess.run(["lscpu"], capture_output=True, text=True, timeout=30).stdout.splitlines()
                             if "NUMA" in line or "Model name" in line or line.startswith("CPU(s)")]
    except Exception:  # noqa: BLE001
        pass
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
