#!/bin/bash
# A/B on one box: early TMEM release in the filter's epilogue (FRG_TC_EARLY) on top of the grouped rare path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" python bench.py --batch 128 --sweep 64,128,256,512,1024 --steps 60 --warmup 5 --no-cpu --no-extra-configs --no-check --e2e-callers 1 > gpurun_out/abe_$name.json 2> gpurun_out/abe_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/abe_$name.err; return; }
  python - "$name" <<'PY'
import json,sys
d=json.load(open("gpurun_out/abe_%s.json" % sys.argv[1]))
print("%-22s" % sys.argv[1], " ".join("F=%d %.4f" % (s["batch"], s["ms_per_step"]) for s in d["sweep"]), flush=True)
PY
}
for rep in 1 2 3 4; do
run grouped_$rep A=1
run early_$rep FRG_TC_EARLY=1
done
