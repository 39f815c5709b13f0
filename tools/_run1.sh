set -x
nvidia-smi -L; nvidia-smi --query-gpu=persistence_mode,clocks.sm --format=csv
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_t1.log 2>&1
echo "== ref arm"; date
timeout 900 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err; echo "ref rc=$?"; date
nvidia-smi --query-gpu=clocks.sm,temperature.gpu,power.draw --format=csv
dmesg 2>/dev/null | grep -i -E "xid|nvrm" | tail -5
echo "== our arm"
timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_ours1.json 2> gpurun_out/r2_ours1.err; echo "ours rc=$?"; date
tail -c 600 gpurun_out/r2_ours1.err
cat gpurun_out/r2_t1.log
