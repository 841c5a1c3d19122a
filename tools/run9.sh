mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "mlpg or fused or golden" 2>&1 | tail -15 > gpurun_out/t9.log
tail -3 gpurun_out/t9.log
for m in 0 2 3; do
GANTTS_B200_MLPG_SOLVE=$m python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b9_solve$m.json 2> gpurun_out/b9_solve$m.err
done
GANTTS_B200_MLPG_SOLVE=3 ncu --set full --clock-control none --import-source on -k regex:mlpg_solve -s 6 -c 2 -o gpurun_out/prof_mlpg_solve_v3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_mlpg3.log 2>&1
grep -h ms_per_step gpurun_out/b9_solve*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['ms_per_step'], d.get('gpu_launches'))
"
