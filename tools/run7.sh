mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/t7.log
tail -3 gpurun_out/t7.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b7.json 2> gpurun_out/b7.err
GANTTS_B200_TAIL=0 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b7_notail.json 2> gpurun_out/b7_notail.err
GANTTS_B200_MLPG_SOLVE=0 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b7_fir.json 2> gpurun_out/b7_fir.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 240 --csv --log-file gpurun_out/launches_r2e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_r2e.log 2>&1
timeout 1500 bash tools/sanitize.sh > gpurun_out/sanitize_run.log 2>&1
