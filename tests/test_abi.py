"""CPU tests: the C-ABI library loads and exports exactly what include/proxsdp_b200.h declares."""
import ctypes
import os
import re

import pytest

from proxsdp_b200 import _abi, build, options

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    path = build.build_extension()          # nvcc cross-compiles without a GPU
    return ctypes.CDLL(path)


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "proxsdp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(proxsdp_b200_\w+)\s*\(", text)))


def test_header_declares_the_seams():
    names = _declared_functions()
    for required in ("proxsdp_b200_solve", "proxsdp_b200_psd_project", "proxsdp_b200_soc_project",
                     "proxsdp_b200_lanczos", "proxsdp_b200_eigh", "proxsdp_b200_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/proxsdp_b200.h but not exported"


def test_pod_sizes_match_python_images(lib):
    for nm, cls in (("problem", _abi.ProblemPOD), ("options", options.OptionsPOD), ("result", _abi.ResultPOD)):
        fn = getattr(lib, f"proxsdp_b200_sizeof_{nm}")
        fn.restype = ctypes.c_int64
        assert fn() == ctypes.sizeof(cls), nm


def test_options_field_order_matches_header():
    text = open(os.path.join(ROOT, "include", "proxsdp_b200_types.h")).read()
    body = text[text.index("typedef struct proxsdp_options {"):text.index("} proxsdp_options_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(int64_t|double)\s+(\w+);", body)
    assert [f[1] for f in fields] == [n for n, _, _ in options.OPTION_FIELDS]
    for (ctype, name), (_, kind, _) in zip(fields, options.OPTION_FIELDS):
        assert (ctype == "double") == (kind == "f"), name


def test_options_mirror_reference_defaults():
    """Same names/defaults as reference src/options.jl (80 fields)."""
    assert len(options.OPTION_FIELDS) >= options.N_REFERENCE_FIELDS
    o = options.Options()
    assert (o.tol_gap, o.tol_psd, o.max_target_rank_krylov_eigs, o.min_size_krylov_eigs) == (1e-4, 1e-7, 16, 100)
    assert (o.eigsolver, o.eigsolver_min_lanczos, o.krylovkit_tol, o.krylovkit_max_iter) == (2, 25, 1e-12, 100)
    assert (o.convergence_window, o.rank_slack, o.delta, o.linsearch_decay) == (200, 3, 0.9999, 0.75)


def test_product_path_has_no_cpu_fallback(monkeypatch, tmp_path):
    """The product binding must fail loudly when the CUDA extension is missing."""
    from proxsdp_b200 import solver
    monkeypatch.setattr(solver, "_lib", None)
    monkeypatch.setattr(solver, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(solver.ExtensionMissing):
        solver.lib()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "proxsdp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", "").replace("oracle's", "").replace(
                    "oracle/", "").lower() or "import oracle" not in src, f
                assert "from oracle" not in src and "import oracle" not in src and "liboracle" not in src, f
