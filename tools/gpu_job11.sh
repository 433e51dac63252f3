#!/bin/bash
mkdir -p gpurun_out/j11
O=gpurun_out/j11
timeout 600 python -m pytest tests/test_gpu_forward.py -q -x -k "homography or driver_flow" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -30 $O/pytest.log
