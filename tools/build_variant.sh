#!/bin/bash
# Development helper: build librufus_gpu with extra -D flags into rufus_b200/variants/<name>.so (A/B kernel variants;
# selected at run time with RG_LIB=rufus_b200/variants/<name>.so). Usage: tools/build_variant.sh name -DSK_CTAS=2 ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/rufus_b200/variants; tmp=$(mktemp -d)
mkdir -p $out
for f in $root/rufus_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$root/include -I$root/rufus_b200/csrc "$@" -c $f -o $tmp/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $tmp/*.o
rm -rf $tmp
echo built $out/$name.so
