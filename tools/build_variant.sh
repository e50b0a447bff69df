#!/bin/bash
# build_variant.sh <git-rev> <name>: builds coupe_b200/csrc of a git revision into gpurun_out-independent
# coupe_b200/lib/variants/lib<name>.so (A/B timing on one GPU box; variants are git-ignored).
set -e
rev=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p $tmp/coupe_b200/csrc $tmp/include $root/coupe_b200/lib/variants
git -C $root show $rev:coupe_b200/csrc/engine.cu > $tmp/coupe_b200/csrc/engine.cu
git -C $root show $rev:coupe_b200/csrc/ffi.cu > $tmp/coupe_b200/csrc/ffi.cu
git -C $root show $rev:coupe_b200/csrc/rcb_kernels.cuh > $tmp/coupe_b200/csrc/rcb_kernels.cuh
git -C $root show $rev:coupe_b200/csrc/Makefile > $tmp/coupe_b200/csrc/Makefile
git -C $root show $rev:include/coupe.h > $tmp/include/coupe.h
git -C $root show $rev:include/coupe_b200.h > $tmp/include/coupe_b200.h
make -C $tmp/coupe_b200/csrc all > /dev/null
cp $tmp/coupe_b200/lib/libcoupe_b200.so $root/coupe_b200/lib/variants/lib$name.so
rm -rf $tmp
echo built $root/coupe_b200/lib/variants/lib$name.so
