#!/bin/bash
mkdir -p gpurun_out/j22
O=gpurun_out/j22
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_dsic.py -q -k "kitti or dsic_plus" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
