set -x
for bk in 32 64; do
  echo "=== BK=$bk"
  GANTTS_B200_BK=$bk timeout 600 python -m pytest tests -m gpu -x -q -k "linear or mlp or fused or gan_step" 2>&1 | tail -5
  GANTTS_B200_BK=$bk timeout 300 python tools/time_mlp.py 2>&1 | grep -v "^+" | tail -20
  GANTTS_B200_BK=$bk timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1
done
