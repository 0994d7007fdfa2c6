#!/bin/bash
# one GPU session: parity tests, smoke, the bench with the driver's flags, launch list under ncu
# usage: tools/gpu_round.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -5 gpurun_out/${TAG}_pytest.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.txt 2>&1; tail -2 gpurun_out/${TAG}_smoke.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "timed_steps", "gpu_launches")}, d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["verify"], d["clocks"])
    for c in d.get("companions", []):
        print(c.get("workload", c)[:60] if isinstance(c, dict) else c, c.get("value"), c.get("frac_of_hbm_peak"), c.get("verify", {}).get("ok") if isinstance(c.get("verify"), dict) else "")
    print(d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
