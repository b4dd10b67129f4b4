#!/bin/bash
# On the GPU box: the C5 leg (tools/bench_legs.py --only c5) once per library variant (build/variants/<name>.so).
cd "$(dirname "$0")/.."
for name in base "$@"; do
  if [ "$name" = base ]; then unset MMO_B200_LIB; else export MMO_B200_LIB=$PWD/build/variants/$name.so; fi
  python tools/bench_legs.py --only c5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['c5_screen']; k=d['kernel_ms_rank0']
print('$name', 'ms %.2f' % d['ms'], 'Mconf/s %.2f' % (d['conformers_per_s']/1e6), 'prep %.2f pair %.2f fix %.2f' % (k['prepare_and_sort'], k['direct_items_kernel'], k['fp64_close_contact_pass']), 'frac %.4f pair-alone %.4f' % (d['roofline']['frac'], d['roofline']['pair_kernel_alone_frac']), 'best', d['best_E'], d['best_conformer'])"
done
