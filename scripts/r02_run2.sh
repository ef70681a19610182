mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_sharded.py tests/test_gpu_big_configs.py tests/test_gpu_maintenance.py -x -q -m gpu > gpurun_out/r02_tests_v6.log 2>&1; tail -15 gpurun_out/r02_tests_v6.log
python tests/tools/diag_big.py 2000 10000 > gpurun_out/r02_diag_big.log 2>&1; tail -3 gpurun_out/r02_diag_big.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_v6.log 2>&1; tail -1 gpurun_out/r02_bench_v6.log
