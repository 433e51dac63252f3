#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "stencil or conv_parity" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 200 python tools/time_small.py > $O/small.txt 2>&1
HESIC_SMALL_NO_TMA=1 timeout 200 python tools/time_small.py >> $O/small.txt 2>&1
cat $O/small.txt
