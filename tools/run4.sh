mkdir -p gpurun_out
GANTTS_B200_CHAIN=7 timeout 900 python -m pytest tests/test_gpu_train_mode.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "chain or fused or sigmoid_single or gan_step" 2>&1 | tail -15 > gpurun_out/t4_chain7.log
tail -3 gpurun_out/t4_chain7.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/t4.log
tail -3 gpurun_out/t4.log
( for cfg in "0 0" "3 0" "7 0" "7 1" "7 2" "7 16" "7 3"; do set -- $cfg; GANTTS_B200_CHAIN=$1 GANTTS_B200_CHAIN_DBG=$2 python tools/time_chain.py 2>/dev/null; done ) > gpurun_out/chain_timing2.log 2>&1
cat gpurun_out/chain_timing2.log
for c in 0 3 7; do GANTTS_B200_CHAIN=$c python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-dropin > gpurun_out/b4_chain$c.json 2> gpurun_out/b4_chain$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 260 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_r2b.log 2>&1
