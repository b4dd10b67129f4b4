#!/bin/bash
# On the GPU box: bench.py once per library variant given (build/variants/<name>.so), one summary line each.
cd "$(dirname "$0")/.."
for name in base "$@"; do
  if [ "$name" = base ]; then unset MMO_B200_LIB; else export MMO_B200_LIB=$PWD/build/variants/$name.so; fi
  python bench.py --steps 20 --warmup 3 --no-aux --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$name', 'value %.2fM' % (d['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], 'K1 %.3f ms' % r['kernel_ms_per_launch'], 'fix %.3f ms' % r['hard_fix_ms_per_launch'], 'frac %.4f' % r['frac'], 'e2e %.2fM' % (d['e2e']['value']/1e6), 'clk', d['clocks']['sm_mhz'])"
done
