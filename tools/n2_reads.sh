# 2 GPUs: single-hop scans with their READS sharded (sweep_main --shard reads, peer-written and NCCL-gathered partial
# accumulators) and the bench's read-sharded companion (verified against a 1-rank scan inside bench.py)
TAG=${1:-r03h}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ARGS="-f 100M:102.4M:19 -P -w blackman-harris --sweeps 24 --intervals 3 --seed 5"
timeout 150 $TR --master-port 29611 -m rtlsdr_b200.sweep_main $ARGS --shard reads -o gpurun_out/${TAG}_reads_peer.csv 2> gpurun_out/${TAG}_reads_peer.err
RTLSDR_B200_NCCL_GATHER=1 timeout 150 $TR --master-port 29612 -m rtlsdr_b200.sweep_main $ARGS --shard reads -o gpurun_out/${TAG}_reads_nccl.csv 2> gpurun_out/${TAG}_reads_nccl.err
timeout 120 python -m rtlsdr_b200.sweep_main $ARGS -o gpurun_out/${TAG}_one.csv 2> gpurun_out/${TAG}_one.err
md5sum gpurun_out/${TAG}_reads_peer.csv gpurun_out/${TAG}_reads_nccl.csv gpurun_out/${TAG}_one.csv | tee gpurun_out/${TAG}_sweep_check.txt
grep -h "sweep_main:" gpurun_out/${TAG}_reads_peer.err gpurun_out/${TAG}_reads_nccl.err | cut -c1-200 | tee -a gpurun_out/${TAG}_sweep_check.txt
rm -f gpurun_out/${TAG}_*.csv
timeout 400 $TR --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["verify"]["ok"])
for c in d["companions"]:
    print(c.get("workload", "")[:70], c.get("value"), c.get("ms_per_step"), c.get("verify", {}).get("ok"), c.get("reads_per_gpu"), c.get("error"))
PY
