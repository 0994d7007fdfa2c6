#!/bin/bash
# multi-GPU session (gpurun --gpus N): the bench at N with the driver's flags, the reference arm, the sharded drivers
# usage: tools/gpu_round_n.sh TAG N
TAG=${1:-r02}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "timed_steps", "gpu_launches", "scaling")}, d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"])
    print(d["verify"]); print(d["config"]["exchange"], d["config"]["host_pinning"])
    for c in d.get("companions", []):
        print(c.get("workload", "")[:60], c.get("value"), c.get("verify"))
except Exception as e:
    print("bench parse failed", e)
PY
RTLSDR_B200_NCCL_GATHER=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}_nccl.json 2> gpurun_out/${TAG}_bench_n${N}_nccl.err
echo "bench (NCCL gather) N=$N rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n${N}_nccl.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['verify'].get('ok'), d['config']['exchange'][:40])"
bash tools/sweep_check.sh $TAG $N
