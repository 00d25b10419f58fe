"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol of include/icsb200.h, the
oracle mirrors the shared signatures, selectors behave like the reference's run-time selection, and the product refuses
to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from icsfoam_b200 import capi, cases, context
from oracle import pyoracle
from tests.conftest import ROOT, has_gpu


def test_library_exports_every_declared_symbol():
    lib = context.lib()
    names = context.exported_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libicsb200.so does not export {n}"


def test_header_cites_reference_for_every_entry_point():
    hdr = open(os.path.join(ROOT, "include", "icsb200.h")).read()
    # every block comment that introduces a hot-path function cites a reference file:line
    for fn in ("icsb200_calc_flux", "icsb200_residual", "icsb200_pseudo_dt", "icsb200_assemble", "icsb200_matrix_mul",
               "icsb200_precondition", "icsb200_solve_delta", "icsb200_update_fields", "icsb200_iterate_dev"):
        i = hdr.index(fn + "(")
        comment = hdr[hdr.rfind("/*", 0, i):i]
        assert re.search(r"\.[CH]:\d+", comment), fn


def test_oracle_mirrors_shared_signatures():
    lib = pyoracle.lib()
    for name in capi.SHARED_SIGNATURES:
        assert hasattr(lib, "orc_" + name), name


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors of the header structs have the size the C compiler gives them."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "icsb200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(icsb200_patch), '
                   'sizeof(icsb200_residuals), sizeof(icsb200_solver_controls), sizeof(icsb200_schemes));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(capi.Patch), C.sizeof(capi.Residuals), C.sizeof(capi.SolverControls), C.sizeof(capi.Schemes)]


@pytest.mark.skipif(has_gpu(), reason="CPU-only check")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        context.Context()


def test_selector_names_follow_reference_dictionaries():
    # src/Make/files:47-51: the reference has HLLC, ROE, AUSMPlusUp; "Rusanov" is this library's own addition (named by the brief)
    assert set(capi.FLUX_NAMES) == {"HLLC", "ROE", "AUSMPlusUp", "Rusanov"}
    assert set(capi.PRECOND_NAMES) == {"LUSGS", "Jacobi"}
    assert capi.DDT_NAMES["steadyState"] == 0
    o = pyoracle.Oracle()
    s = capi.default_schemes()
    s.flux_scheme = 7
    with pytest.raises(capi.ApiError, match="Unknown convectiveFluxScheme"):
        o.schemes_set(s)


def test_call_order_is_enforced():
    o = pyoracle.Oracle()
    with pytest.raises(capi.ApiError):
        o._call("state_set", None, None, None)


def test_case_library_matches_tutorial_controls():
    b = cases.bump()
    assert b.mesh.n_cells == 3 * 66 * 54 and b.controls.n_directions == 5 and b.controls.max_iter == 10
    assert abs(b.controls.rel_tol - 1e-2) < 1e-18 and b.mesh.solutionD == [1, 1, -1]
    s = cases.shock_tube()
    assert s.mesh.n_cells == 500 and s.schemes.flux_scheme == capi.FLUX_ROE and s.schemes.ddt_scheme == capi.DDT_BACKWARD
    o = cases.onera_box(6)
    assert o.controls.rel_tol == 0.1 and o.schemes.pseudo_co_num == 100.0
