#!/bin/bash
# device-arm bench of the three workloads under a list of environment settings.
# Usage: scripts/knobs.sh tag "ENV1=a ENV2=b" "ENV1=c" ...   (use "-" for the default environment)
tag=$1; shift
out=gpurun_out; mkdir -p $out
for cfg in "$@"; do
  name=$(echo "$cfg" | tr ' =/.' '____')
  for wl in ${WL:-ggm_ar_p50_n100_M10000_R100_c1024 ggm_ar_p500_n1000_M10000_R100_c1024 loglin_ar_p100_n100000_M10000_R100_c1024}; do
    short=$(echo $wl | cut -d_ -f1,3)
    if [ "$cfg" = "-" ]; then e=""; else e="$cfg"; fi
    env $e python bench.py --no-other --no-cpu-baseline --no-e2e --workload $wl --steps 3 --warmup 1 > $out/${tag}_${name}_${short}.json 2> $out/${tag}_${name}_${short}.err
    python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_${name}_${short}.json"))
    r=d["roofline"]
    print("%-52s %-12s %8.1f M/s  %8.1f ms  ident %s  hit %.4f evaluated %.0f derived %s"%("$cfg","$short",d["value"]/1e6,d["ms_per_step"],d["repeat_runs_identical"],r["memo_hit_rate"],r["sets_evaluated_per_step"],r.get("sets_derived_per_step")))
except Exception as ex:
    print("$cfg $short FAILED", ex)
PY
  done
done
