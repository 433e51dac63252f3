#!/bin/bash
mkdir -p gpurun_out/j19
O=gpurun_out/j19
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -q -k "masked or joint or fixture" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
HESIC_ONE_STREAM=1 timeout 300 python tools/layer_times.py 16 hesic_plus 3 > $O/layer_times_plus.txt 2>&1; head -8 $O/layer_times_plus.txt
