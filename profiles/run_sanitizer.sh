#!/bin/bash
# compute-sanitizer over the kernels added / reworked in round 2 (small parity cases only: the tools slow kernels down 10-100x)
out=gpurun_out; mkdir -p $out
K1="ellipse or circle or fuse_ln or mask_geometry or cls_attention_kernel or cls_head_kernel or pool_score_select_fused or empty_image or gem_token_pool_vs or packed_masks or crop or max_n_smaller"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "$K1" > $out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 $out/r2_memcheck.log
K2="pool_score_select_fused or ellipse or fuse_ln or packed_masks"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "$K2" > $out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 $out/r2_racecheck.log
