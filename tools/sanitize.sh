#!/bin/bash
# compute-sanitizer over a reduced -m gpu subset: memcheck, racecheck and initcheck on the kernels
# with hand-rolled synchronisation -- the tcgen05 GEMM (mbarrier pipelines), the cooperative LSTM (global step barrier),
# MLPG (shared-memory sweeps), the fused step.  Run on a GPU box:  bash tools/sanitize.sh  -> gpurun_out/sanitize_*.log
# Sizes are small: the sanitizer slows kernels down by 10-100x.
set -u
mkdir -p gpurun_out
SEL='test_sequence_mask_bit_exact or test_masked_mse_golden or test_multi_stream_mlpg_golden or test_mlpg_sizes_vs_f64_banded or test_linear_layer_fwd_bwd or test_lstm_golden_forward or test_fused_gan_step_small_vs_oracle or test_gan_step_golden or test_sru_layer_vs_port or test_edge_shapes'
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 1200 -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$?" | tee -a gpurun_out/sanitize_summary.log
  grep -E "ERROR SUMMARY|passed|failed|Race reported|Invalid|Uninitialized" gpurun_out/sanitize_$tool.log | tail -5 | tee -a gpurun_out/sanitize_summary.log
done
