mkdir -p gpurun_out
timeout 300 python tests/time_torch_eager_step.py > gpurun_out/step_vs_torch_eager.md 2> gpurun_out/step_vs_torch_eager.err
cat gpurun_out/step_vs_torch_eager.md; tail -3 gpurun_out/step_vs_torch_eager.err
timeout 560 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_mode.py tests/test_gpu_properties.py tests/test_gpu_eval_ops.py -m gpu -q --timeout 540 -k "not full_size and not cfg3 and not cfg5 and not cfg2_sized and not cfg2_shapes and not 1000 and not 2000" > gpurun_out/sanitize_memcheck_all.log 2>&1
echo "== memcheck whole suite (reduced sizes) exit=$? (124 = time limit)" | tee -a gpurun_out/sanitize_summary.log
grep -E "ERROR SUMMARY|passed|failed|Invalid __" gpurun_out/sanitize_memcheck_all.log | tail -6 | tee -a gpurun_out/sanitize_summary.log
