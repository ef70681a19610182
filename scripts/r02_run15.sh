mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_group.py -x -q -m gpu 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 > gpurun_out/r02_tests_v12.log; tail -3 gpurun_out/r02_tests_v12.log
