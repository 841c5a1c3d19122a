mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 > gpurun_out/t10.log
tail -3 gpurun_out/t10.log
for m in 2 3; do
GANTTS_B200_MLPG_SOLVE=$m python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b10_solve$m.json 2> gpurun_out/b10_solve$m.err
done
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b10_cfg3.json 2> gpurun_out/b10_cfg3.err
python tools/profile_dropin.py > gpurun_out/dropin_profile.md 2> gpurun_out/dropin_profile.err
GANTTS_B200_MLPG_SOLVE=3 ncu --set full --clock-control none --import-source on -k regex:mlpg_solve -s 6 -c 2 -o gpurun_out/prof_mlpg_solve_v4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_mlpg4.log 2>&1
python - <<'PY'
import json
for f in ['b10_solve2','b10_solve3','b10_cfg3']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        k=d['roofline']['kernels']
        print(f, d['ms_per_step'], {n:(round(v['ms_per_step']*1e3,1),v['launches_per_step']) for n,v in k.items() if 'mlpg' in n or 'lstm' in n})
    except Exception as e: print(f, 'ERR', e)
PY
head -3 gpurun_out/dropin_profile.md
