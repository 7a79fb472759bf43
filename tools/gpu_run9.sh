#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
for f in ${FLAGS:-0 2}; do
echo "=== flags $f"
BW_MEGA_FLAGS=$f timeout 300 python tools/mega_trace.py 2>&1 | tee gpurun_out/t9_trace_$f.log | grep -E "step span|work avg|sum slowest|staged" | head -40
BW_MEGA_FLAGS=$f BW_TIME=1 timeout 300 python tools/profile_decode.py 2>&1 | tail -4
done
