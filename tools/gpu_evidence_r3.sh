#!/bin/bash
# Round-3 (second session of round 2) evidence in ONE gpurun call: tests, smoke, bench (both arms), per-layer table, launch list, ncu --set full rows.
#   tools/gpu_evidence_r3.sh <tag> [golden]
TAG=${1:-ev3}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
rm -f gpurun_out/parity_stats.jsonl
if [ "$2" == "golden" ]; then
  timeout 600 python tests/golden/make_codec_golden.py gpurun_out/codec_golden > $O/codec_golden.log 2>&1; echo "golden rc=$?" >> $O/codec_golden.log
  cp gpurun_out/codec_golden/codec_* tests/golden/ 2>/dev/null
fi
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cp gpurun_out/parity_stats.jsonl $O/ 2>/dev/null
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_hesic.json 2> $O/bench_hesic.err; echo "bench rc=$?" >> $O/bench_hesic.err
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layer_times_hesic.txt 2>&1
timeout 300 python tools/layer_times.py 16 hesic_plus 3 > $O/layer_times_hesic_plus.txt 2>&1
timeout 300 python tools/kernel_breakdown.py dsic 8 > $O/dsic_breakdown.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --cpu-iters 1 --no-extras --sustain-s 0.01 > $O/bench_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:"conv_tc_pair_kernel" -s 2 -c 1 -o $O/ncu_conv2 python tools/run_dominant.py 4 16 > $O/ncu_conv2.log 2>&1
timeout 300 $NCU -k regex:"conv_tc_first" -s 2 -c 1 -o $O/ncu_first python tools/run_layer.py 3 128 5 2 0 512 512 16 1 0 3 > $O/ncu_first.log 2>&1
for L in "3 128 5 2 0 512 512 16 1 0" "128 128 5 2 0 256 256 16 1 0" "128 128 5 2 1 128 128 16 2 0" "128 960 5 1 0 32 32 16 0 1"; do timeout 120 python tools/time_layer.py $L 20 >> $O/layer_isolated.txt 2>&1; done
timeout 200 python tools/time_small.py >> $O/layer_isolated.txt 2>&1
timeout 600 $NCU -k regex:"warp_rgb_kernel|gaussian_tile|conv_small_kernel|conv_head_kernel|eb_kernel|rowpad4_pair_kernel|spatial_max|upsample|sse_|images_u8" -c 16 -o $O/ncu_elem python bench.py --steps 1 --warmup 0 --cpu-iters 1 --no-extras --sustain-s 0.01 > $O/ncu_elem.log 2>&1
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cut -c1-400 $O/bench_hesic.json; tail -2 $O/bench_hesic.err; tail -3 $O/codec_golden.log 2>/dev/null
