"""CPU-side checks of the drop-in boundary: libhgl.so builds, loads, and exports exactly what include/hgl.h
declares, with the argument counts the ctypes binding uses.  No compute call (there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hybridgl_b200 import build
    build.build()
    from hybridgl_b200 import _lib
    return _lib


def _header_decls():
    text = open(os.path.join(ROOT, "include", "hgl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"HGL_API\s+([\w\s\*]+?)\b(hgl_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        decls[m.group(2)] = n
    return decls


def test_header_matches_binding_and_exports(lib):
    decls = _header_decls()
    assert set(decls) == set(lib.SIGNATURES), (set(decls) ^ set(lib.SIGNATURES))
    for name, n in decls.items():
        assert len(lib.SIGNATURES[name][1]) == n, name
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert {s for s in exported if s.startswith("hgl_")} == set(decls)


def test_library_loads_and_reports_errors(lib):
    h = lib.load()
    assert h.hgl_version() >= 100
    # argument validation happens before any CUDA call, so it is testable without a device
    rc = h.hgl_prep(None, None, None, None, 1, 1, 1, 8, 8, 8, 0, 0, None, None, None, None)
    assert rc == -1 and b"null pointer" in h.hgl_last_error()
    rc = h.hgl_mask_grid(1, 1, 8, 8, 99, 1, 1, None, None, None)
    assert rc == -1 and b"bad shape" in h.hgl_last_error()
    assert h.hgl_heat_pool_workspace_bytes(2, 10, 3, 48, 64, 8) > 0


def test_sass_is_sm100(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout or "SM100a" in out.stdout or "sm_100" in out.stdout
    assert "STG.E.NA" in out.stdout        # streaming (no-allocate) stores of the prep kernel
    # the Blackwell-only machinery the design relies on is really in the binary (B200_PROFILING.md, SASS mnemonics):
    assert "UTCHMMA" in out.stdout         # tcgen05.mma (mask pooling on the 5th-generation tensor cores)
    assert "LDTM" in out.stdout            # tcgen05.ld (TMEM accumulator read-back in the epilogue)
    assert "UBLKCP" in out.stdout          # cp.async.bulk (TMA bulk copies: bit-row stages of prep_main)
    assert "UTMALDG.3D" in out.stdout      # cp.async.bulk.tensor.3d through a CUtensorMap (token tiles of the pooling kernel)
    assert "UTCBAR" in out.stdout          # tcgen05.commit (MMA completion onto an mbarrier)
    assert "SYNCS.ARRIVE.TRANS64" in out.stdout   # mbarrier expect_tx / arrive (full / empty barriers of the stage ring)


def test_ops_fail_loudly_without_cuda(lib):
    import torch
    from hybridgl_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(TypeError):
        ops.masks_to_grid(torch.zeros((1, 8, 8), dtype=torch.bool), 2)     # CPU tensors are rejected: no CPU fallback
