#!/bin/bash
TAG=${1:-r3j}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_dsic.py -q -x > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gaussian|conv_head" --csv --log-file $O/gauss.csv python bench.py --steps 1 --warmup 1 --cpu-iters 1 --no-extras --sustain-s 0.01 > $O/ncu.log 2>&1
grep -o "gaussian_tile_kernel[^,]*\|conv_head_kernel[^,]*\|\"[0-9.]*\"$" $O/gauss.csv | paste - - | tail -8
