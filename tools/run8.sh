mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/t8.log
tail -3 gpurun_out/t8.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b8.json 2> gpurun_out/b8.err
GANTTS_B200_MLPG_SOLVE=3 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b8_solve3.json 2> gpurun_out/b8_solve3.err
GANTTS_B200_MLPG_SOLVE=3 ncu --set full --clock-control none --import-source on -k regex:mlpg -s 6 -c 2 -o gpurun_out/prof_mlpg_solve_r2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_mlpg1.log 2>&1
GANTTS_B200_MLPG_SOLVE=0 ncu --set full --clock-control none --import-source on -k regex:mlpg -s 6 -c 2 -o gpurun_out/prof_mlpg_fir_r2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_mlpg2.log 2>&1
