mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02v_bench_n8.json 2> gpurun_out/r02v_bench_n8.err
echo "rc=$?"; tail -c 600 gpurun_out/r02v_bench_n8.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02v_bench_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["verify"]["ok"], d["e2e"])
print(d["companions"][0]["value"], d["host_issue_ms_per_step"])
PY
