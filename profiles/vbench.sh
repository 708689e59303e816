for v in base lcap5 lcap4 base; do
  if [ "$v" = base ]; then unset SNPGPU_LIB; else export SNPGPU_LIB=$PWD/variants/libsnpgpu_$v.so; fi
  python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],3), 'ms', round(d['roofline']['avg_launch_ms'],5))"
done
