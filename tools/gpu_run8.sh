#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-900)" | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run pytest_gpu 1500 python -m pytest tests -q -p no:cacheprovider --timeout 600 -m gpu
run bench 900 python bench.py --steps 5 --warmup 3
run bench_perop 900 env BW_NO_MEGA=1 python bench.py --steps 5 --warmup 3
run bench_ref 900 python bench.py --impl reference --steps 2 --warmup 1
run ncu_mega 600 env BW_NO_GRAPH=1 BW_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 5 -c 1 -o gpurun_out/prof_mega3 python tools/profile_decode.py
run launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1v3.csv python bench.py --steps 1 --warmup 1
cat gpurun_out/summary.txt
