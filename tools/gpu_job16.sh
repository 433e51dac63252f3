#!/bin/bash
mkdir -p gpurun_out/j16
timeout 300 python tools/dbg_dsic_oplevel.py 2>&1 | tail -20
