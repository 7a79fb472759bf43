#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/mega_trace.py 2>&1 | tee gpurun_out/t11_trace.log | grep -E "step span|work avg|sum slowest|staged" | head -40
BW_TIME=1 timeout 300 python tools/profile_decode.py 2>&1 | tail -4
