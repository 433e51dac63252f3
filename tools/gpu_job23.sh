#!/bin/bash
mkdir -p gpurun_out/j23
O=gpurun_out/j23
timeout 900 python -m pytest tests/test_gpu_dsic.py tests/test_gpu_forward.py -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python tools/dsic_time.py 8 512 512 3 2>&1 | tail -1
