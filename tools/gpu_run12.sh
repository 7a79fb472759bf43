#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
timeout 900 python -m pytest tests -q -p no:cacheprovider -m gpu 2>&1 | tail -8 | tee gpurun_out/t12_tests.log
BW_TIME=1 timeout 300 python tools/profile_decode.py 2>&1 | tail -4
BW_NO_FUSED_SELECT=1 BW_TIME=1 timeout 300 python tools/profile_decode.py 2>&1 | tail -2
