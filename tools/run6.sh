mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/t6.log
tail -3 gpurun_out/t6.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b6.json 2> gpurun_out/b6.err
GANTTS_B200_BRES=0 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b6_nobres.json 2> gpurun_out/b6_nobres.err
GANTTS_B200_MLPG_SOLVE=0 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b6_fir.json 2> gpurun_out/b6_fir.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 220 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_r2d.log 2>&1
