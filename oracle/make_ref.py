#!/usr/bin/env python
"""Recipe for oracle/_ref: the REAL reference path as a CPU baseline that can travel to the GPU box.

    python oracle/make_ref.py            # in the build container (needs /root/reference); __graft_entry__.build() runs it

The reference (fhgyuanshen/HybridGL) is pure Python: nothing to compile.  What the hot path executes lives in four places --
Hybridgl_main.py (the inline prep / scoring / guidance blocks, :92-125 and :153-230), utils.py (relation_boxes, gen_dir_mask,
Compute_IoU), model/backbone.py (TF.resize of the masks, calculate_score) and third_party/modified_CLIP/clip (imported by
model/backbone.py) -- and those files are copied VERBATIM into oracle/_ref/, which is git-ignored (reference sources never enter
the repository or its history) but not gpurun-ignored, so bench.py's `--impl reference` and `cpu_baseline` can time the
reference's own code on the GPU box's host cores (kind: "reference") instead of the numpy port (kind: "port").
oracle/ref_runner.py drives it; tests/ use it, when present, as one more pin of the oracle.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HGL_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["Hybridgl_main.py", "utils.py", "model/backbone.py"]
DIRS = ["third_party/modified_CLIP/clip"]


def main() -> int:
    if not os.path.isdir(REF):
        print(f"make_ref: {REF} not found (only the build container has the reference); nothing to do")
        return 0
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy2(os.path.join(REF, rel), dst)
    for rel in DIRS:
        dst = os.path.join(DST, rel)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, rel), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Verbatim copies of reference files made by oracle/make_ref.py (git-ignored; CPU baseline only).\n")
    print("make_ref: copied", ", ".join(FILES + DIRS), "->", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
