#!/bin/bash
# compute-sanitizer over a REDUCED -m gpu subset: memcheck / initcheck on the streaming kernels, the
# MLPG kernels (shared-memory strips), the cooperative LSTM (hand-rolled global step barrier), the SRU scan and one small
# tcgen05 step; racecheck on the MLPG and LSTM kernels (the ones that hand data between threads through shared memory).
# Run on a GPU box:  bash tools/sanitize.sh  -> gpurun_out/sanitize_*.log + sanitize_summary.log
# The sanitizer slows kernels down 10-100x: a first attempt over ten test functions incl. the parametrised tcgen05 GEMM
# tests did not finish memcheck inside a 25-minute slot, hence the small selection and the per-tool time limit.
set -u
mkdir -p gpurun_out
: > gpurun_out/sanitize_summary.log
MEM='test_sequence_mask_bit_exact or test_masked_mse_golden or test_multi_stream_mlpg_golden or test_stream_indexing_bit_exact or (test_lstm_golden_forward and simt) or test_edge_shapes'
RACE='test_multi_stream_mlpg_golden or (test_lstm_golden_forward and simt) or test_masked_mse_golden'
run() {   # tool, selection, seconds
  timeout $3 compute-sanitizer --tool $1 --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout $3 -k "$2" > gpurun_out/sanitize_$1.log 2>&1
  echo "== $1 exit=$? (124 = time limit)" | tee -a gpurun_out/sanitize_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Race reported|Invalid __|Uninitialized __" gpurun_out/sanitize_$1.log | tail -6 | tee -a gpurun_out/sanitize_summary.log
}
run memcheck "$MEM" ${SAN_SECONDS:-420}
run racecheck "$RACE" ${SAN_SECONDS:-420}
run initcheck "$MEM" ${SAN_SECONDS:-420}
# second tier (the first tier takes ~6 s per tool): the tcgen05 GEMM / fused-step path at golden sizes, the SRU scan, the
# MLPG substitution kernels at several lengths, the LSTM backward -- memcheck only (racecheck does not model the async proxy
# / mbarrier hand-offs of the TMA + tcgen05 kernels)
MEM2='test_gan_step_golden or test_fused_gan_step_small_vs_oracle or test_mlp_model_golden or test_sru_layer_vs_port or test_mlpg_sizes_vs_f64_banded or (test_lstm_fwd_bwd_vs_torch_cpu) or test_mlp_stack_sigmoid_single_output'
run2() {
  timeout $2 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout $2 -k "$1" > gpurun_out/sanitize_memcheck2.log 2>&1
  echo "== memcheck tier 2 exit=$? (124 = time limit)" | tee -a gpurun_out/sanitize_summary.log
  grep -E "ERROR SUMMARY|passed|failed|Invalid __" gpurun_out/sanitize_memcheck2.log | tail -6 | tee -a gpurun_out/sanitize_summary.log
}
if [ "${SAN_TIER2:-1}" = "1" ]; then run2 "$MEM2" ${SAN_SECONDS2:-600}; fi
# third tier: memcheck over the whole GPU suite at reduced sizes (~25 s under the sanitizer)
if [ "${SAN_TIER3:-1}" = "1" ]; then
  timeout ${SAN_SECONDS3:-560} compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_mode.py tests/test_gpu_properties.py tests/test_gpu_eval_ops.py -m gpu -q --timeout 540 -k "not full_size and not cfg3 and not cfg5 and not cfg2_sized and not cfg2_shapes and not 1000 and not 2000" > gpurun_out/sanitize_memcheck_all.log 2>&1
  echo "== memcheck whole suite (reduced sizes) exit=$? (124 = time limit)" | tee -a gpurun_out/sanitize_summary.log
  grep -E "ERROR SUMMARY|passed|failed|Invalid __" gpurun_out/sanitize_memcheck_all.log | tail -6 | tee -a gpurun_out/sanitize_summary.log
fi
