#!/bin/bash
TAG=${1:-epi16}; O=gpurun_out/$TAG; mkdir -p $O
L="3 128 5 2 0 512 512 16 1 0"
timeout 120 python tools/time_layer.py $L > $O/t.txt 2>&1
HESIC_TC_EPI8=1 timeout 120 python tools/time_layer.py $L >> $O/t.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:"conv_tc_kernel" -s 2 -c 1 -o $O/conv1_epi16 python tools/run_layer.py $L 3 > $O/ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 3 > $O/bench.json 2> $O/bench.err
cat $O/t.txt; tail -4 $O/pytest.log; cut -c1-240 $O/bench.json
