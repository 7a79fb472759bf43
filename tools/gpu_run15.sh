#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
timeout 120 tools/microbench/sync_bench 2>&1 | tee gpurun_out/sync_bench2.log
timeout 1200 python -m pytest tests -q -p no:cacheprovider -m gpu 2>&1 | grep -v Warning | tail -6 | tee gpurun_out/t15_tests.log
BW_MEGA_FLAGS=32 timeout 300 python tools/mega_trace.py 2>&1 | grep -v Warning | tee gpurun_out/t15_trace_v2.log | head -12
