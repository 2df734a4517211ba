#!/bin/bash
# pre-pass sample size (FRG_TC_PRE_MIN_ROWS -> stride 8/16/32/64 at 1 M rows) against batch size, sustained clocks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for rep in 1 2; do
for mr in 125000 62000 31000 15000; do
  FRG_TC_PRE_MIN_ROWS=$mr python bench.py --batch 128 --sweep 64,128,256,512,1024 --steps 60 --warmup 5 --no-cpu --no-extra-configs --no-check --e2e-callers 1 > gpurun_out/stride_$mr.json 2> gpurun_out/stride_$mr.err || { echo "$mr FAILED"; tail -3 gpurun_out/stride_$mr.err; continue; }
  python - "$mr" <<'PY'
import json,sys
d=json.load(open("gpurun_out/stride_%s.json" % sys.argv[1]))
print("min_rows", sys.argv[1], " ".join("F=%d %.4f" % (s["batch"], s["ms_per_step"]) for s in d["sweep"]), flush=True)
PY
done; done
