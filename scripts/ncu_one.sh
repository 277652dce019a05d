#!/bin/bash
# `ncu --set full` capture of k_mh_run on one bench workload.  Usage: scripts/ncu_one.sh tag workload [env...]
tag=$1; wl=$2; shift; shift
out=gpurun_out; mkdir -p $out
B="python bench.py --no-other --no-cpu-baseline --no-e2e --workload $wl --steps 1 --warmup 1"
env "$@" $B > $out/${tag}_plain.log 2>&1 &&
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mh_run -s 1 -c 1 -f -o $out/${tag} $B > $out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $out/${tag}_plain.log | cut -c1-300
