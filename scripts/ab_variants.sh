#!/bin/bash
# Developer A/B on one box: the verification bench, the ICP probes and ADD-S for every variant build.
#   scripts/ab_variants.sh name1 name2 ...   ("cur" = csrc/libisr.so)
V=imagesequenceregistrationfor6dposeestimationlabeling_b200/csrc/variants
for n in "$@"; do
  if [ "$n" = cur ]; then unset ISR_LIBISR_PATH; else export ISR_LIBISR_PATH=$PWD/$V/$n.so; fi
  v=$(python bench.py --steps 8 --warmup 3 --skip-cpu --skip-extra --skip-icp 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']))")
  i8=$(python scripts/probe_icp_allshards.py 8 2>&1 | tail -1 | sed 's/.*slowest //')
  i1=$(python scripts/probe_icp_allshards.py 1 2>&1 | tail -1 | sed 's/.*slowest //')
  a=$(python scripts/probe_adds.py 2>&1 | tail -1 | sed 's/,.*//;s/.*: //')
  echo "$n: verify $v | icp8 $i8 | icp1 $i1 | adds $a"
done
