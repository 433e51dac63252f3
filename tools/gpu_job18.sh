#!/bin/bash
mkdir -p gpurun_out/j18
O=gpurun_out/j18
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -q -x -k "enhancement" > $O/memcheck_en.log 2>&1; echo "rc=$?" >> $O/memcheck_en.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_dsic.py -q -x -k "channels_last or group_norm" > $O/memcheck_dsic.log 2>&1; echo "rc=$?" >> $O/memcheck_dsic.log
tail -6 $O/memcheck_en.log; tail -6 $O/memcheck_dsic.log
