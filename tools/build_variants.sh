#!/bin/bash
# Builds experiment variants of the C-ABI library into build/ (git-ignored, travels to the GPU box). Each variant is the
# current csrc/ tree with extra -D flags or with one file replaced:   tools/build_variants.sh name "flags" [file=replacement]...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; flags=$2; shift 2
W=$(mktemp -d)
mkdir -p $W/visrag_b200 $W/include
cp -r $ROOT/visrag_b200/csrc $W/visrag_b200/csrc
cp $ROOT/include/visrag_b200.h $W/include/
for rep in "$@"; do cp "${rep#*=}" "$W/visrag_b200/csrc/${rep%%=*}"; done
rm -f $W/visrag_b200/csrc/*.o
make -s -C $W/visrag_b200/csrc -j8 NVCCFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v $flags" > /dev/null
grep -A3 "attention4_tcgen05_kernelILi80ELb1" $W/visrag_b200/csrc/attention.ptxas.log | grep spill || true
cp $W/visrag_b200/libvisrag_b200.so $ROOT/build/libvr_$name.so
rm -rf $W
echo "built build/libvr_$name.so"
