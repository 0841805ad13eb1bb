#!/bin/bash
# Compare builds of libmadrona_b200 on one box (same clocks, same run):
#   scripts/ab_bench.sh "libA.so libB.so ..." [bench args...]
LIBS=$1; shift
for rep in 1 2; do
  for lib in $LIBS; do
    MADRONA_B200_LIB=$lib python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'ms/step %.4f  e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
  done
done
