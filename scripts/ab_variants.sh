#!/bin/bash
# A/B of hot-kernel builds: bench.py (no extras) per build; "main" = probqa_b200/lib, others = probqa_b200/lib/exp/<name>
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -3
for v in main ${VARIANTS:-noscreen} main ${VARIANTS:-noscreen}; do
  lib=probqa_b200/lib/exp/$v/libPqaCore.so; [ $v = main ] && lib=probqa_b200/lib/libPqaCore.so
  echo "### $v"
  PQA_B200_LIB=$PWD/$lib python bench.py --no-extras --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g e2e %.4g kernel_ms %.4f'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms']))"
done
