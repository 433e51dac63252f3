#!/bin/bash
# Round evidence: tests, smoke, bench (both models), launch list, ncu --set full captures of each kernel class.
mkdir -p gpurun_out/ev
O=gpurun_out/ev
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_hesic.json 2> $O/bench_hesic.err
timeout 600 python bench.py --steps 20 --warmup 5 --model hesic_plus --cpu-iters 2 > $O/bench_hesic_plus.json 2> $O/bench_hesic_plus.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
HESIC_ONE_STREAM=1 timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layer_times_hesic.txt 2>&1
HESIC_ONE_STREAM=1 timeout 300 python tools/layer_times.py 16 hesic_plus 3 > $O/layer_times_hesic_plus.txt 2>&1
timeout 300 python tools/dsic_time.py 8 512 512 3 > $O/dsic_time.txt 2>&1
timeout 300 python tools/kernel_breakdown.py en 16 > $O/en_breakdown.txt 2>&1
timeout 300 python tools/run_en_layer.py 16 5 all > $O/en_layer_times.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --cpu-iters 1 > $O/bench_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:"conv_tc(_pair)?_kernel" -s 2 -c 1 -o $O/ncu_conv2 python tools/run_dominant.py 4 16 > $O/ncu_conv2.log 2>&1
timeout 600 $NCU -k regex:"warp_kernel|gaussian_tile|conv_small_kernel|conv_head_kernel|eb_kernel|rowpad_kernel|spatial_max|upsample_kernel|sse_kernel" -c 14 -o $O/ncu_elem python bench.py --steps 1 --warmup 0 --cpu-iters 1 > $O/ncu_elem.log 2>&1
timeout 300 $NCU -k regex:en_conv_kernel -s 1 -c 1 -o $O/ncu_en_res1 python tools/run_en_layer.py 16 1 res1 > $O/ncu_en_res1.log 2>&1
tail -2 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/bench_hesic.json | cut -c1-300; cat $O/bench_hesic_plus.json | cut -c1-200; cat $O/bench_reference.json | cut -c1-200; cat $O/dsic_time.txt | tail -1; head -1 $O/en_breakdown.txt
