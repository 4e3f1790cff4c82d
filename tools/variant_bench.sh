#!/bin/bash
# usage: tools/variant_bench.sh FILE.cu "-DX=1" "-DX=2" ...   rebuilds FILE's object with each flag set and benches
F=$1; shift
for V in "$@"; do
  rm -f rufus_b200/csrc/$F.o
  make lib EXTRA_NVFLAGS="$V" > /dev/null 2>&1 || { echo "build failed for $V"; continue; }
  echo "=== $V"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print('value', round(d['value'],2), 'frac', round(d['roofline']['frac'],4), {k:round(v['ms_avg'],4) for k,v in d['kernels_ms'].items() if k in ('partition_kernel','tile_insert_kernel','filter_hits_kernel')}, d['result'])"
done
rm -f rufus_b200/csrc/$F.o; make lib > /dev/null 2>&1
