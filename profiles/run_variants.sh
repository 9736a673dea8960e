#!/bin/bash
# stand-alone stage timings for every library variant under profiles/_variants/ (tuning only):
#   gpurun -- 'bash profiles/run_variants.sh tag "pytest -k expr" stage ...'
tag=${1:-v}; kexpr=${2:-grid}; shift; shift
timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" 2>&1 | tail -4
for lib in profiles/_variants/*.so; do
  echo "== $lib"
  HGL_LIB=$PWD/$lib timeout 300 python profiles/kbench.py ${tag}_$(basename $lib .so) "$@" 2>&1 | grep -v "^shape"
done
