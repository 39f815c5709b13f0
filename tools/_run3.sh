(timeout 2400 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -60) > gpurun_out/r2_t3.log 2>&1
cat gpurun_out/r2_t3.log
