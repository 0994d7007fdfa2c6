mkdir -p gpurun_out
BENCH_E2E_WEIGHTS="3,2" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02s_bench_n2_weighted.json 2> gpurun_out/r02s_bench_n2_weighted.err
echo "rc=$?"; tail -c 800 gpurun_out/r02s_bench_n2_weighted.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02s_bench_n2_weighted.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["verify"]["ok"], d["e2e"])
PY
