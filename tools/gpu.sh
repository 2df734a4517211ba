#!/bin/bash
# Gate every GPU call on a clean local build + CPU test pass: a broken build must not burn box time.
set -e
cd /root/repo
python -c "from facerecognition_infrenceengine_b200 import build; build.build()" >/dev/null
python -m pytest tests -x -q -m "not gpu" >/dev/null 2>&1 || { echo "CPU tests fail - not going to the GPU"; exit 1; }
exec /usr/local/graft/bin/gpurun "$@"
