(timeout 1800 python -m pytest tests/test_gpu_fullsize.py tests/test_mesh_fixtures.py -m gpu -q -k "five or nine or limits or screw" 2>&1 | tail -30) > gpurun_out/r2_t5.log 2>&1
cat gpurun_out/r2_t5.log
