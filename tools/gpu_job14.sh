#!/bin/bash
mkdir -p gpurun_out/j14
HESIC_TRACE_SIMT=1 timeout 300 python tools/dsic_time.py 1 256 256 1 2> gpurun_out/j14/trace.txt | tail -1
sort gpurun_out/j14/trace.txt | uniq -c | sort -rn | head -20
