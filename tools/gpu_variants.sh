#!/bin/bash
# Round-2 opener: parity + timing of every compile-time variant of decode_mega_kernel (decode_mega.cu V_* bits) and of
# decode_mega3 in one gpurun call.  Variant bits: 1 no trace, 2 relaxed barriers over tagged activations, 4 per-head
# readiness counters in front of the attention phases, 8 barrier counter sharded 4 ways, 16 producer-only arrival behind
# the attention phases, 32 up to 32 decoder steps per launch.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
BW_AB="64:0,64:1,64:3,64:5,64:9,64:17,64:21,64:23,64:33,64:55,64:7,64:15,64:0" timeout 600 python tools/mega_ab.py 2>&1 | tail -15 | tee gpurun_out/variants_ab.log
for v in 3 5 17 23 33 55 15; do
  echo "== variant $v: model parity tests"
  BW_MEGA_VARIANT=$v timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider --tb=short -k "teacher_forced or batch_rows" 2>&1 | grep -v Warning | tail -4 | tee gpurun_out/variants_tests_$v.log
done
echo "== decode_mega3 (BW_MEGA_FLAGS=192: attention fused with its out-projection, 4 grid barriers per layer)"
BW_MEGA_FLAGS=192 timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider --tb=short -k "teacher_forced or batch_rows" 2>&1 | grep -v Warning | tail -6 | tee gpurun_out/variants_tests_mega3.log
BW_AB="64:0,192:0,192:33,64:0,192:0,192:33" timeout 300 python tools/mega_ab.py 2>&1 | tail -7 | tee gpurun_out/variants_ab_mega3.log
