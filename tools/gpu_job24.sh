#!/bin/bash
for f in 0 1 2; do echo "== force=$f"; HESIC_TC_TWO_STAGING=$f HESIC_ONE_STREAM=1 timeout 300 python tools/layer_times.py 16 hesic 3 2>&1 | head -12; done
