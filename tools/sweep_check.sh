#!/bin/bash
# Hop-sharded drivers on N GPUs of one box (run under `gpurun --gpus N`, N >= 2):
#  - rtlsdr_b200.sweep_main: CSV byte-identical on 1 GPU, N GPUs with peer-written reports and N GPUs with the NCCL
#    gather, three intervals (pinned input cubes and report buffers are reused), plus randomised hop order
#  - host/rtl_power_gpu -t N (C side, one handle per device): CSV byte-identical to -t 1
# usage: tools/sweep_check.sh TAG N
TAG=${1:-r02}; N=${2:-2}
mkdir -p gpurun_out
O=gpurun_out/${TAG}
A="-f 24M:300M:1k -c 20% -w hamming --sweeps 4 --intervals 3"
TR="timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 python -m rtlsdr_b200.sweep_main $A -o ${O}_sweep_n1.csv
$TR --master-port 29521 -m rtlsdr_b200.sweep_main $A -o ${O}_sweep_peer.csv 2> ${O}_sweep_peer.err
RTLSDR_B200_NCCL_GATHER=1 $TR --master-port 29522 -m rtlsdr_b200.sweep_main $A -o ${O}_sweep_nccl.csv 2> ${O}_sweep_nccl.err
$TR --master-port 29523 -m rtlsdr_b200.sweep_main $A --random-hops 3 -o ${O}_sweep_shuffled.csv 2> ${O}_sweep_shuffled.err
grep -h "interval reports" ${O}_sweep_peer.err ${O}_sweep_nccl.err
E="RTLSDR_SYNTH_MODE=biased RTLSDR_SYNTH_SEED=3 RTLSDR_SYNTH_PARAM=21 RTL_POWER_PASSES=4 RTL_POWER_REPORTS=3"
env $E RTL_POWER_TIMESTAMP="2026-01-01, 00:00:00" host/_build/rtl_power_gpu -f 24M:300M:1k -c 20% -w hamming ${O}_cli_t1.csv 2> ${O}_cli_t1.err
env $E RTL_POWER_TIMESTAMP="2026-01-01, 00:00:00" host/_build/rtl_power_gpu -f 24M:300M:1k -c 20% -w hamming -t $N ${O}_cli_tN.csv 2> ${O}_cli_tN.err
grep -h "GPU workers" ${O}_cli_tN.err
{
  wc -l ${O}_sweep_n1.csv ${O}_cli_t1.csv
  md5sum ${O}_sweep_n1.csv ${O}_sweep_peer.csv ${O}_sweep_nccl.csv ${O}_sweep_shuffled.csv
  md5sum ${O}_cli_t1.csv ${O}_cli_tN.csv
} | tee ${O}_sweep_check.txt
rm -f ${O}_sweep_*.csv ${O}_cli_*.csv
