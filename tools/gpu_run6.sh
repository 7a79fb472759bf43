#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
for f in 0 4; do echo "=== flags $f"; BW_MEGA_FLAGS=$f python tools/mega_trace.py 2>&1 | grep -E "step span|work avg|sum slowest"; done 2>&1 | tee gpurun_out/flags.log
