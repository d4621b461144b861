#!/bin/bash
# Developer A/B: build libisr.so variants with extra -D flags into csrc/variants/<name>.so
#   scripts/build_variants.sh name1 "-DISR_OPT_X=1" name2 "-DISR_OPT_Y=1 ..."
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/imagesequenceregistrationfor6dposeestimationlabeling_b200/csrc
mkdir -p $CSRC/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (
    tmp=/tmp/isr_variant_$name
    rm -rf $tmp && mkdir -p $tmp/pkg/csrc $tmp/include
    cp $CSRC/*.cu $CSRC/*.cuh $CSRC/Makefile $tmp/pkg/csrc/
    cp $ROOT/include/isr.h $tmp/include/
    sed -i 's#-I../../include#-I'$tmp'/include#' $tmp/pkg/csrc/Makefile
    make -C $tmp/pkg/csrc -j4 EXTRA_NVCCFLAGS="$flags" > $tmp/build.log 2>&1 || { tail -20 $tmp/build.log; exit 1; }
    cp $tmp/pkg/csrc/libisr.so $CSRC/variants/$name.so
    echo "built $name ($flags)"
  ) &
done
wait
