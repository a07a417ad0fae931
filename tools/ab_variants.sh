#!/bin/bash
# A/B timing of variant builds of the library (tools: build.build_variant -> variants/*.so).  Usage: tools/ab_variants.sh out.log lib1.so lib2.so ...
out=$1; shift
: > $out
for lib in default "$@"; do
  for rep in 1 2; do
    if [ "$lib" = default ]; then unset AVSE_B200_LIB; else export AVSE_B200_LIB=$PWD/$lib; fi
    echo "== $lib rep $rep" >> $out
    python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu ${AB_EXTRA:-} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        inv = d.get('inverse', {})
        print('step_ms %.4f fwd_kernel_ms %.4f frac %.4f inv_ms %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], inv.get('ms_per_step')))
    elif l: print(l)
" >> $out
  done
done
cat $out
