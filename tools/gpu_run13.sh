#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
T=tests/test_pipeline_gpu.py::test_pipeline_vs_live_oracle_small30
timeout 300 python -m pytest $T -q -p no:cacheprovider --tb=short 2>&1 | grep -v Warning | tail -25 > gpurun_out/t13_small30_mega.log
BW_NO_MEGA=1 timeout 300 python -m pytest $T -q -p no:cacheprovider --tb=short 2>&1 | grep -v Warning | tail -25 > gpurun_out/t13_small30_perop.log
tail -12 gpurun_out/t13_small30_mega.log; tail -6 gpurun_out/t13_small30_perop.log
BW_AB="0:1,0:4,0:8,4:1,8:1,8:4,12:4" timeout 300 python tools/mega_ab.py 2>&1 | tail -9 | tee gpurun_out/t13_ab1.log
BW_AB="0:1,16:1,16:4,28:4,28:8,24:4" timeout 300 python tools/mega_ab.py 2>&1 | tail -8 | tee gpurun_out/t13_ab2.log
BW_MEGA_FLAGS=28 BW_MEGA_REP=4 timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider --tb=short 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/t13_model_flags28.log
