#!/bin/bash
# tunables sweep for small batches (sustained clocks, per-point preload)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" python bench.py --batch 128 --sweep 64,128,256 --steps 40 --warmup 5 --no-cpu --no-extra-configs --no-check --e2e-callers 1 > gpurun_out/tune_$name.json 2> gpurun_out/tune_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/tune_$name.err; return; }
  python - "$name" <<'PY'
import json,sys
d=json.load(open("gpurun_out/tune_%s.json" % sys.argv[1]))
print(sys.argv[1], " ".join("F=%d %.4f ms (dom %.4f, %s MHz)" % (s["batch"], s["ms_per_step"], s["kernel_ms"], s["sm_mhz"]) for s in d["sweep"]), flush=True)
PY
}
run default A=1
run default2 A=1
run prechunks74 FRG_TC_PRE_CHUNKS=74
run prechunks37 FRG_TC_PRE_CHUNKS=37
run stride64 FRG_TC_PRE_MIN_ROWS=8000
run stride16 FRG_TC_PRE_MIN_ROWS=40000
run fused FRG_TC_FUSED=1
run fused_div16 FRG_TC_FUSED=1 FRG_TC_PROBE_DIV=16
run nopdl FRG_PDL=0
