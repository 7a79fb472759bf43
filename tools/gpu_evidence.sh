#!/bin/bash
# One gpurun call that produces the round's evidence: GPU parity tests, the bench line, the ncu launch list of the bench
# command, one `ncu --set full` capture of the persistent decoder-step kernel and its barrier timeline.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.csv
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -q -p no:cacheprovider -m gpu 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_bench_under_ncu.log | cut -c1-300
BW_STEPS=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 20 -c 1 \
  -o gpurun_out/${TAG}_mega python tools/profile_decode.py > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
timeout 300 python tools/mega_trace.py 2>&1 | tee gpurun_out/${TAG}_trace.log | tail -30
ls -la gpurun_out | head -30
