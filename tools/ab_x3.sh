# A/B timing of fp16x3 kernel build variants (GPU box): bash tools/ab_x3.sh base collect ...
for v in "$@"; do
  if [ $v = base ]; then unset NFB_LIB_PATH; else export NFB_LIB_PATH=nerfies_b200/_variants/libnfb_$v.so; fi
  timeout 200 python bench.py --precision ${PREC:-fp16x3} --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-also --no-parity 2>/dev/null </dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']/1e6,1), 'M/s fine_ms', round(d['roofline']['kernel_ms'],2), 'coarse_ms', round(d['roofline']['coarse_kernel_ms'],2), 'clk', d['clocks']['sm_mhz'], 'W', d['clocks'].get('power_w'), d['clocks'].get('power_w_median'), d['clocks']['reasons'])"
done
