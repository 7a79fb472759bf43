#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400)" | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
PT="python -m pytest -q -p no:cacheprovider --timeout 300"
run ops 400 $PT tests/test_ops_gpu.py -k "gemv or layernorm"
run model 900 $PT tests/test_model_gpu.py
run time_graph 300 env BW_TIME=1 python tools/profile_decode.py
run launches 600 env BW_NO_GRAPH=1 BW_STEPS=2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/profile_decode.py
run ncu_gemv 600 env BW_NO_GRAPH=1 BW_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 900 -c 6 -o gpurun_out/prof_gemv python tools/profile_decode.py
run ncu_xattn 600 env BW_NO_GRAPH=1 BW_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:cross_attn -s 96 -c 2 -o gpurun_out/prof_xattn python tools/profile_decode.py
run bench 900 python bench.py --steps 5 --warmup 3
cat gpurun_out/summary.txt
