(timeout 1800 python -m pytest tests/test_mesh.py tests/test_mesh_fixtures.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r2_t6.log 2>&1
cat gpurun_out/r2_t6.log
python tools/c4_time.py 2>&1 | tail -6
