#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -q -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/t10_tests.log
BW_NO_GRAPH=1 BW_STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 5 -c 1 -o gpurun_out/prof_mega4 python tools/profile_decode.py > gpurun_out/t10_ncu.log 2>&1; tail -2 gpurun_out/t10_ncu.log
