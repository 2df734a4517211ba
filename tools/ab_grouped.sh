#!/bin/bash
# A/B on one box: grouped slow path (FRG_TC_GROUPED) x pre-pass stride, sustained clocks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" python bench.py --batch 128 --sweep 64,128,256,512,1024 --steps 60 --warmup 5 --no-cpu --no-extra-configs --no-check --e2e-callers 1 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/ab_$name.err; return; }
  python - "$name" <<'PY'
import json,sys
d=json.load(open("gpurun_out/ab_%s.json" % sys.argv[1]))
print("%-22s" % sys.argv[1], " ".join("F=%d %.4f" % (s["batch"], s["ms_per_step"]) for s in d["sweep"]), flush=True)
PY
}
for rep in 1 2 3; do
run base_s32_$rep A=1
run grouped_s32_$rep FRG_TC_GROUPED=1
run base_s16_$rep FRG_TC_PRE_MIN_ROWS=62000
run grouped_s16_$rep FRG_TC_GROUPED=1 FRG_TC_PRE_MIN_ROWS=62000
done
