"""CPU tier: the C-ABI library builds/loads and exports every symbol include/dgb200.h declares
(no compute calls here -- there is no GPU in this tier)."""
import ctypes
import os
import re

import pytest

from helpers import REPO


def _declared_symbols():
    src = open(os.path.join(REPO, "include", "dgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dgb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = _declared_symbols()
    for must in ("dgb_bsr_apply", "dgb_bsr_residual", "dgb_block_gs_pass", "dgb_block_relax_sweep",
                 "dgb_block_diag_inverse", "dgb_restrict", "dgb_prolong_add", "dgb_vcycle",
                 "dgb_metrics", "dgb_assemble_poisson", "dgb_assemble_rhs"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from dg_multigrid_solver_b200 import build
    lib = build.build_library()
    L = ctypes.CDLL(lib)
    for name in _declared_symbols():
        assert hasattr(L, name), f"{name} declared in include/dgb200.h but not exported"
    assert L.dgb_abi_version() == 6
    assert L.dgb_partials_len() >= 1024


def test_ctypes_binding_covers_the_header():
    from dg_multigrid_solver_b200 import _lib
    assert set(_lib.SIGNATURES) == set(_declared_symbols())
    assert ctypes.sizeof(_lib.SmootherCtl) == 32
    _lib.load()


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct that crosses the ABI, C compiler vs ctypes."""
    import subprocess
    from dg_multigrid_solver_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dgb200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dgb_operator), sizeof(dgb_level),'
                   ' sizeof(dgb_vcycle_opts), sizeof(dgb_smoother_ctl), sizeof(dgb_tables_desc),'
                   ' offsetof(dgb_level, post_omega), offsetof(dgb_vcycle_opts, coarse_inverse),'
                   ' offsetof(dgb_level, R));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.Operator), ctypes.sizeof(_lib.Level), ctypes.sizeof(_lib.VcycleOpts),
            ctypes.sizeof(_lib.SmootherCtl), ctypes.sizeof(_lib.TablesDesc), _lib.Level.post_omega.offset,
            _lib.VcycleOpts.coarse_inverse.offset, _lib.Level.R.offset]
    assert got == want


def test_ctypes_constants_match_the_header():
    """Every #define the Python side mirrors (flags, sweep modes, transfer kinds, smoother ids, return codes)."""
    from dg_multigrid_solver_b200 import _lib
    src = open(os.path.join(REPO, "include", "dgb200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"^#define (DGB_[A-Z0-9_]+)\s+(-?\d+)", src, flags=re.M)}
    want = {"DGB_FLAG_PERIODIC_I": _lib.FLAG_PERIODIC_I, "DGB_FLAG_PERIODIC_J": _lib.FLAG_PERIODIC_J,
            "DGB_FLAG_MINV": _lib.FLAG_MINV, "DGB_FLAG_GHOST_LO": _lib.FLAG_GHOST_LO, "DGB_FLAG_GHOST_HI": _lib.FLAG_GHOST_HI,
            "DGB_GS_LEXICOGRAPHIC": _lib.GS_LEXICOGRAPHIC, "DGB_GS_REDBLACK": _lib.GS_REDBLACK,
            "DGB_GS_SLAB_LEXICOGRAPHIC": _lib.GS_SLAB_LEXICOGRAPHIC, "DGB_TRANSFER_P": _lib.TRANSFER_P,
            "DGB_TRANSFER_H": _lib.TRANSFER_H, "DGB_UNSUPPORTED": _lib.UNSUPPORTED,
            "DGB_COARSE_SMOOTHER": _lib.COARSE_SMOOTHER, "DGB_COARSE_DIRECT": _lib.COARSE_DIRECT,
            "DGB_VCYCLE_ENTRY_PRIMED": _lib.VCYCLE_ENTRY_PRIMED,
            "DGB_SMOOTHER_BLOCK_GS_PYAMG": _lib.SMOOTHER_IDS["block_gauss_seidel_pyamg"],
            "DGB_SMOOTHER_BLOCK_JACOBI": _lib.SMOOTHER_IDS["block_jacobi"],
            "DGB_SMOOTHER_BLOCK_GS": _lib.SMOOTHER_IDS["block_gauss_seidel"]}
    for name, value in want.items():
        assert defs[name] == value, name


def test_product_has_no_cpu_path():
    """Without a GPU the product refuses to run (no oracle / CPU fallback on the product path)."""
    import torch
    from dg_multigrid_solver_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.DgbError):
        _lib.require_cuda()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "dg_multigrid_solver_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "dgoracle" not in txt and "oracle/" not in txt, f
