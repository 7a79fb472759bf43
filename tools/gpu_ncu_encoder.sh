#!/bin/bash
# `ncu --set full` captures of the kernels in front of the decoder (log-mel, tcgen05 GEMM, tcgen05 encoder attention):
# tensor-pipe utilisation and DRAM throughput per kernel.  One GPU, one short process per capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD TRANSFORMERS_OFFLINE=1 HF_HUB_OFFLINE=1 TOKENIZERS_PARALLELISM=false
TAG=${1:-r2}
for spec in "gemm_tc:40:4" "attn_enc_tc:8:2" "logmel:2:2" "layernorm_rows:8:2"; do
  IFS=: read -r pat skip cnt <<< "$spec"
  BW_STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt \
    -o gpurun_out/${TAG}_$pat python tools/profile_decode.py > gpurun_out/${TAG}_ncu_$pat.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$pat.log
  ncu -i gpurun_out/${TAG}_$pat.ncu-rep --page raw --csv 2>/dev/null | python - "$pat" <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    ix = [hdr.index(w) for w in want if w in hdr]
    for r in rows[2:]:
        print(" | ".join(f"{hdr[i]}={r[i]}" for i in ix))
PY
done
