#!/bin/bash
# correctness of every tcgen05 shape group, per-layer times, bench
TAG=${1:-r3}; O=gpurun_out/$TAG; mkdir -p $O
for g in s1 s2 deconv gdn row head big edge; do timeout 300 python tools/tc_check.py $g >> $O/check.txt 2>&1; done
grep -c "^OK" $O/check.txt; grep "^BAD\|rror" $O/check.txt | head
timeout 300 python tools/layer_times.py 16 hesic 3 > $O/layers_hesic.txt 2>&1; head -12 $O/layers_hesic.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 3 > $O/bench.json 2> $O/bench.err; cut -c1-300 $O/bench.json
