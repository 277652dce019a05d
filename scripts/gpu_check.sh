#!/bin/bash
# GPU tests, then the bench with and without the two-warps-per-chain kernels.  Usage: scripts/gpu_check.sh tag [nospec]
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "pytest rc=$?"; tail -12 $out/${tag}_gputests.log
timeout 400 python bench.py --no-cpu-baseline > $out/${tag}_bench_spec.json 2> $out/${tag}_bench_spec.err; echo "rc=$?"
if [ "$2" = "nospec" ]; then PDG_SPEC=0 timeout 400 python bench.py --no-cpu-baseline > $out/${tag}_bench_nospec.json 2> $out/${tag}_bench_nospec.err; echo "rc=$?"; fi
python - <<PY
import json
for f in ("spec","nospec"):
    try:
        d=json.load(open("$out/${tag}_bench_%s.json"%f))
    except Exception as e:
        print(f, "no json", e); continue
    print(f, "cfg2 %.1f M (e2e %.1f) ms %.1f ident %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["repeat_runs_identical"]), d.get("ess"))
    for k,o in d["other_workloads"].items(): print("   ",k,"%.1f M (e2e %.1f) ms %.1f ident %s"%(o["value"]/1e6,o["e2e"]["value"]/1e6,o["ms_per_step"],o["repeat_runs_identical"]))
PY
tail -3 $out/${tag}_bench_spec.err
