#!/bin/bash
# parity tests matching $2 (pytest -k) + stand-alone stage timings $3...   gpurun -- 'bash profiles/run_k.sh tag "rle or iou" rle pack'
tag=${1:-k}; kexpr=${2:-rle}; shift; shift
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" 2>&1 | tail -15
timeout 300 python profiles/kbench.py $tag "$@" 2>&1 | tail -12
