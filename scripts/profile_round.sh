#!/bin/bash
# One gpurun call: launch list of the default bench and `ncu --set full` captures of k_mh_run on the
# three bench workloads (B200_PROFILING.md recipe).  Usage: scripts/profile_round.sh r02a
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
B="python bench.py --no-other --no-cpu-baseline --no-e2e"
$B --steps 2 --warmup 1 > $out/${tag}_plain_cfg2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_cfg2.csv \
    $B --steps 2 --warmup 1 > $out/${tag}_ncu_launches.log 2>&1
for wl in ggm_ar_p50_n100_M10000_R100_c1024 ggm_ar_p500_n1000_M10000_R100_c1024 loglin_ar_p100_n100000_M10000_R100_c1024; do
  short=$(echo $wl | cut -d_ -f1,3)
  $B --workload $wl --steps 1 --warmup 1 > $out/${tag}_plain_${short}.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mh_run -s 1 -c 1 -f -o $out/${tag}_mh_${short} \
      $B --workload $wl --steps 1 --warmup 1 > $out/${tag}_ncu_${short}.log 2>&1
  echo "$short ncu rc=$?"
done
python scripts/scorer_roofline.py > $out/${tag}_scorer_roofline.txt 2>&1
echo "scorer rc=$?"
ls -la $out
