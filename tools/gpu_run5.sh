#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-500)" | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
PT="python -m pytest -q -p no:cacheprovider --timeout 300"
run model_decode 900 $PT tests/test_model_gpu.py -k "teacher_forced or batch_rows"
run time_mega 300 env BW_TIME=1 python tools/profile_decode.py
run trace 300 python tools/mega_trace.py
cat gpurun_out/summary.txt; tail -22 gpurun_out/trace.log
