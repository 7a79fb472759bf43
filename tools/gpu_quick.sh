#!/bin/bash
# A short, hang-proof check (every step under its own tight timeout): default paths first, then the opt-in encoder features.
#     gpurun --timeout 480 -- 'bash tools/gpu_quick.sh <tag>'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
O=gpurun_out/${1:-quick}
echo "== default: pair GEMM / attention / GELU op tests"
timeout 150 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -m gpu -x -k "pair or attn_enc or gelu" 2>&1 | tail -2 | tee ${O}_ops.log
echo "== default: encoder parity at large-v3 dims"
timeout 200 python -m pytest tests/test_large_gpu.py -q -p no:cacheprovider -m gpu -s -k "encoder" 2>&1 | grep -E "^\[|passed|failed" | cut -c1-200 | tee ${O}_large_encoder.log
echo "== opt-in: 16 epilogue warps (op tests)"
BW_GEMM2_EPI16=1 timeout 100 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -m gpu -x -k "epilogues or gemm_plain and pair" 2>&1 | tail -2 | tee ${O}_epi16_ops.log
echo "== opt-in timings: B = 1 and B = 64"
for spec in "all:1:BW_GEMM2_EPI16=1 BW_ENC_GRAPH=1 BW_ENC_PDL=1" "all:64:BW_GEMM2_EPI16=1 BW_ENC_GRAPH=1 BW_ENC_PDL=1" "graph+pdl:1:BW_ENC_GRAPH=1 BW_ENC_PDL=1" \
            "epi16:64:BW_GEMM2_EPI16=1" "graph:1:BW_ENC_GRAPH=1" "default:1:BW_ENC_GRAPH=0"; do
  IFS=: read -r name a envs <<< "$spec"
  echo "-- $name A=$a" | tee -a ${O}_times.log
  env $envs BW_A=$a BW_G=1 BW_CHUNK_S=30 BW_TIME=1 BW_STEPS=2 timeout 75 python tools/profile_decode.py 2>&1 | grep -E "encode|decode step|rror" | cut -c1-120 | tee -a ${O}_times.log
done
