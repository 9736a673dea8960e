out=gpurun_out
timeout 400 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "pool_score_select_fused or two_halves or circle or gem_token_pool_vs or packed_masks or score_select_vs or geometry" > $out/r2_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -2 $out/r2_initcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "two_halves or cls_ or mask_geometry or fuse_ln or blur" > $out/r2_racecheck2.log 2>&1; echo "racecheck2 rc=$?"; tail -2 $out/r2_racecheck2.log
