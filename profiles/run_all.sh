#!/bin/bash
# full GPU validation + stand-alone stage timings in one call:  gpurun --timeout 2400 -- 'bash profiles/run_all.sh tag'
tag=${1:-run}
bash profiles/run_gpu.sh $tag
timeout 300 python profiles/kbench.py $tag 2>&1 | tail -20
