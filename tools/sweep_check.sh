# hop-sharded sweep driver: CSV must be byte-identical on 1 GPU, 2 GPUs (peer-written reports) and 2 GPUs (NCCL gather)
mkdir -p gpurun_out
A="-f 24M:300M:1k -c 20% -w hamming --sweeps 4 --intervals 2"
python -m rtlsdr_b200.sweep_main $A -o gpurun_out/sweep_n1.csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 -m rtlsdr_b200.sweep_main $A -o gpurun_out/sweep_n2_peer.csv 2> gpurun_out/sweep_n2_peer.err
RTLSDR_B200_NCCL_GATHER=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 -m rtlsdr_b200.sweep_main $A -o gpurun_out/sweep_n2_nccl.csv 2> gpurun_out/sweep_n2_nccl.err
wc -l gpurun_out/sweep_n1.csv; md5sum gpurun_out/sweep_n1.csv gpurun_out/sweep_n2_peer.csv gpurun_out/sweep_n2_nccl.csv | cut -c1-60
grep -n "rror" gpurun_out/sweep_n2_peer.err | head -5
