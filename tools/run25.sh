mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 > gpurun_out/t25.log
tail -3 gpurun_out/t25.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b25.json 2> gpurun_out/b25.err
GANTTS_B200_PDL=0 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b25_nopdl.json 2> gpurun_out/b25_nopdl.err
python - <<'PY'
import json
for f in ['b25','b25_nopdl']:
    d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
    print(f, d['ms_per_step'], d['gpu_launches_per_step'], d['roofline']['gemm_family']['ms_per_step'], d['e2e']['ms_per_step'], d['timed_repeats']['ms_per_step_all'])
PY
