mkdir -p gpurun_out
CFGS="${CFGS:-10:28 10:16 10:20 10:40 10:56 12:28 8:28 10:12}"
for cfg in $CFGS; do
  for st in ${MODES:-0 1}; do
    P=$(python -c "
be,ds=map(int,'$cfg'.split(':')); print(max(64,(400<<20)//(2*(1<<be)*ds)))")
    RTLSDR_GPU_BOXCAR_STREAM=$st timeout 120 python tools/scan_bench.py --boxcar $cfg --passes $P --steps 20 --no-kernel-time 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('RES', '$cfg', $st, round(d['ms_per_step'],4), round(d['Msamples_per_s']), round(d['frac_of_6542.7'],3))"
  done
done 2>&1 | tee -a gpurun_out/s2_stream_ab.txt
