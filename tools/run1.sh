mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -40 > gpurun_out/t1.log
GANTTS_B200_F32_STAGE=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q --timeout 900 2>&1 | tail -25 > gpurun_out/t1_stage.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err
GANTTS_B200_F32_STAGE=1 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/b1_stage.json 2> gpurun_out/b1_stage.err
tail -5 gpurun_out/t1.log; tail -3 gpurun_out/t1_stage.log
