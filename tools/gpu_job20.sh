#!/bin/bash
mkdir -p gpurun_out/j20
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -q -k "warp or fixture or full_size or driver" > gpurun_out/j20/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/j20/pytest.log
tail -4 gpurun_out/j20/pytest.log
timeout 120 python tools/run_warp.py 16 5 2>&1 | tail -1
