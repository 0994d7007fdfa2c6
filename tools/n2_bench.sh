# 2 GPUs: bench.py with the driver's flags (hop-sharded headline + the read-sharded companion)
TAG=${1:-r03m}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["verify"]["ok"])
for c in d["companions"]:
    print(c.get("workload", "")[:70], c.get("value"), c.get("ms_per_step"), c.get("verify", {}).get("ok"), c.get("reads_per_gpu"), c.get("error"))
PY
