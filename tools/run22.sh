mkdir -p gpurun_out
SAN_SECONDS=120 SAN_SECONDS2=540 bash tools/sanitize.sh
timeout 600 python -m pytest tests/test_gpu_train_mode.py -m gpu -q --timeout 600 -k "without_discriminator or fused" 2>&1 | tail -8
