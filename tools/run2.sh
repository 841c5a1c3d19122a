mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_mode.py -m gpu -q --timeout 300 -k "chain" 2>&1 | tail -40 > gpurun_out/t2_chain.log
tail -3 gpurun_out/t2_chain.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/t2.log
tail -3 gpurun_out/t2.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/b2.json 2> gpurun_out/b2.err
GANTTS_B200_CHAIN=0 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/b2_nochain.json 2> gpurun_out/b2_nochain.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2a.log 2>&1
