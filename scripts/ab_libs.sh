#!/bin/bash
# A/B experimental builds of libdwg.so (build_exp/lib_*.so, see sampler_b200/_build.py DWG_LIB) on one workload:
#   scripts/ab_libs.sh OUT.log "lib1 lib2 ..." [ab_sweep args...]
out=$1; libs=$2; shift 2
: > $out
for l in $libs; do
  echo "== $l" >> $out
  DWG_LIB=$PWD/build_exp/lib_$l.so python scripts/ab_sweep.py "$@" 2>&1 | grep '"variant"' >> $out
done
cat $out
