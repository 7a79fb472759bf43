#!/bin/bash
# First-contact run on the B200 box: every group in its own process (a trapped kernel kills only its group),
# everything bounded by `timeout`, logs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$? $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400)" | tee -a gpurun_out/summary.txt
}
rm -f gpurun_out/summary.txt
run probe 120 python -c "
import torch; p=torch.cuda.get_device_properties(0); print(p); print('smem optin', p.shared_memory_per_block_optin, 'sms', p.multi_processor_count)
import subprocess; print(subprocess.run(['nvidia-smi'],capture_output=True,text=True).stdout)
import os; print('cpus', os.cpu_count())"
PT="python -m pytest -q -p no:cacheprovider --timeout 240"
run ops_simt 400 $PT tests/test_ops_gpu.py -k "simt or layernorm or gemv"
run gemm_tc128 300 $PT tests/test_ops_gpu.py -k "gemm_plain and tc128"
run gemm_tc64 300 $PT tests/test_ops_gpu.py -k "gemm_plain and tc64"
run gemm_tc256 300 $PT tests/test_ops_gpu.py -k "gemm_plain and tc256"
run gemm_tcauto 300 $PT tests/test_ops_gpu.py -k "(gemm_plain and tcauto) or (gemm_epilogues and tc)"
run attn_tc 300 $PT tests/test_ops_gpu.py -k "attn_enc and tc"
run logmel 400 $PT tests/test_model_gpu.py -k "logmel"
run model_simt 600 $PT tests/test_model_gpu.py -k "encoder_parity and simt"
run model_tc 600 $PT tests/test_model_gpu.py -k "encoder_parity and tc"
run model_decode 600 $PT tests/test_model_gpu.py -k "teacher_forced or batch_rows"
run smoke 300 python __graft_entry__.py --smoke
run bench 900 python bench.py --steps 3 --warmup 3
cat gpurun_out/summary.txt
