#!/bin/bash
# ncu --set full of ONE kernel (regex $2) from the stand-alone stage runner.   gpurun -- 'bash profiles/run_ncu_one.sh tag regex stage'
tag=${1:-one}; kre=${2:-rle_to_bits}; stage=${3:-rle}
out=gpurun_out; mkdir -p $out
timeout 300 python profiles/kbench.py $tag 2>&1 | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s 3 -c 2 -o $out/${tag}_one -f \
    python profiles/kbench.py ${tag}_ncu $stage > $out/${tag}_ncu_one.log 2>&1; echo "ncu rc=$?"
ls -la $out | tail -4
