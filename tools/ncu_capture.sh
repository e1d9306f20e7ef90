#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each, after the same command exited 0 without ncu);
# the reports stay on the box, tools/ncu_summary.py turns them into the text + JSON summaries that go to profiles/.
# usage: tools/ncu_capture.sh <tag> <commit> [C3 C4 C5 OSD ...]
TAG=$1; COMMIT=$2; shift 2
WHAT=${@:-C3 C4 C5}
O=gpurun_out
cap() {  # name kernel-regex workload variant kernel_rev syndromes  bench-args...
  local name=$1 rx=$2 wl=$3 var=$4 rev=$5 syn=$6; shift 6
  python bench.py "$@" > /dev/null 2> $O/${TAG}_${name}_plain.err || { echo "$name: plain run failed"; tail -3 $O/${TAG}_${name}_plain.err; return; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o /tmp/${TAG}_${name} python bench.py "$@" > /dev/null 2> $O/${TAG}_${name}_ncu.err
  python tools/ncu_summary.py /tmp/${TAG}_${name}.ncu-rep $O/${TAG}_${name}_sass.txt --json $O/${TAG}_ncu_${name}.json \
      workload=$wl variant=$var kernel_rev=$rev syndromes=$syn commit=$COMMIT "command=bench.py $*" > $O/${TAG}_${name}_ncu_full.txt 2>&1
  head -24 $O/${TAG}_${name}_ncu_full.txt | cut -c1-130
}
COMMON="--steps 1 --warmup 1 --no-cpu --no-sweep --no-e2e"
for w in $WHAT; do
  case $w in
    C3)  cap c3 bp_smem C3 exact 2 10000000 $COMMON ;;
    C3MS) cap c3_minsum bp_smem C3 minsum 2 10000000 $COMMON --variant minsum ;;
    C4)  cap c4 bp_persistent C4 exact 1 1000000 $COMMON --workload C4 --batch 1000000 ;;
    C5)  cap c5 bp_persistent C5 exact 1 65536 $COMMON --workload C5 --batch 65536 ;;   # (a 262144-syndrome launch takes ncu more than ten minutes to replay)
    C2)  cap c2 bp_smem C2 exact 2 1000000 $COMMON --workload C2 ;;
  esac
done
