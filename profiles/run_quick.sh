#!/bin/bash
# Quick GPU check: parity tests + bench (overlapped and single-stream), no ncu.   gpurun --timeout 900 -- 'bash profiles/run_quick.sh tag'
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 300 python bench.py ${BENCH_ARGS} > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
