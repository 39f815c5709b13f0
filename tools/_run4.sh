(timeout 1200 python -m pytest tests/test_scatter.py tests/test_scatter_oracle.py tests/test_mesh_fixtures.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r2_t4.log 2>&1
cat gpurun_out/r2_t4.log
timeout 600 python bench.py --config c5 --steps 3 --warmup 1 > gpurun_out/r2_c5_n1.json 2> gpurun_out/r2_c5_n1.err; tail -c 300 gpurun_out/r2_c5_n1.err; cat gpurun_out/r2_c5_n1.json
