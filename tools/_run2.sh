set -x
python tools/make_goldens.py --solid-only 2>&1 | tail -8
cp gpurun_out/golden/solid_angle.npz tests/golden/
(timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40) > gpurun_out/r2_t2.log 2>&1
cat gpurun_out/r2_t2.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 400 gpurun_out/r2_bench2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['binding']['march_ms_per_view'], d['checksum'])"
