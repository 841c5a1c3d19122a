mkdir -p gpurun_out
export GANTTS_B200_BENCH_TRACE=1
for i in 1 2 3; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540+i)) bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2b_$i.json 2> gpurun_out/b2b_$i.err
echo "== job $i"; grep "trace. rank 0" gpurun_out/b2b_$i.err | head -14 | cut -c1-120
done
