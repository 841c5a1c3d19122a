mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/t3.log
tail -3 gpurun_out/t3.log
( for cfg in "0 0" "1 0" "1 1" "1 2" "1 4" "1 8" "1 3" "1 15" "1 16" "1 31"; do set -- $cfg; GANTTS_B200_CHAIN=$1 GANTTS_B200_CHAIN_DBG=$2 python tools/time_chain.py 2>/dev/null; done ) > gpurun_out/chain_timing.log 2>&1
cat gpurun_out/chain_timing.log
ITERS=2 ncu --set full --clock-control none --import-source on -k regex:chain_pair -s 40 -c 4 -o gpurun_out/prof_chain_r2 python tools/time_chain.py > gpurun_out/ncu_chain.log 2>&1
