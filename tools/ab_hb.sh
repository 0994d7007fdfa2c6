for cfg in "100M:100.1M:100 -F 9 -w blackman" "100M:100.4M:100 -F 9" "100M:100.05M:100 -F 0" "100M:100.2M:100 -F 9" "100M:100.8M:100 -F 9"; do
  for e in "" "RTLSDR_GPU_NO_HB_STREAM=1"; do
    env $e python tools/scan_bench.py --range $cfg --passes 4096 --steps 10 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('HB', '$cfg', '$e', 'P=%d' % d['plan']['downsample_passes'], round(d['ms_per_step'],4), round(d['Msamples_per_s']), round(d['frac_of_6542.7'],3), d['launches_per_step'])"
  done
done
