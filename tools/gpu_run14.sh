#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
BW_MEGA_FLAGS=32 timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider --tb=short 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/t14_model_v2.log
BW_AB="0:1,32:1,0:1,32:1" timeout 300 python tools/mega_ab.py 2>&1 | tail -5 | tee gpurun_out/t14_ab.log
timeout 300 python -m pytest tests/test_pipeline_gpu.py -q -p no:cacheprovider --tb=short 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/t14_pipeline.log
BW_MEGA_FLAGS=32 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/t14_bench_v2.json | cut -c1-400
