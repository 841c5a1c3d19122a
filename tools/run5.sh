mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/t5.log
tail -3 gpurun_out/t5.log
GANTTS_B200_MLPG_SOLVE=0 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b5_fir.json 2> gpurun_out/b5_fir.err
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/b5.json 2> gpurun_out/b5.err
for c in 1 2; do GANTTS_B200_CHAIN=$c python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b5_chain$c.json 2> gpurun_out/b5_chain$c.err; done
python tools/time_cudnn_lstm.py > gpurun_out/lstm_vs_cudnn.md 2> gpurun_out/lstm_vs_cudnn.err
cat gpurun_out/lstm_vs_cudnn.md
python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b5_cfg3.json 2> gpurun_out/b5_cfg3.err
python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b5_cfg5.json 2> gpurun_out/b5_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 220 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_r2c.log 2>&1
