"""The C-ABI library loads on a CPU-only box and exports every symbol include/esfm_match.h declares
(no compute calls here: there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "esfm_match.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(esfm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("esfm_init", "esfm_bank_create", "esfm_bank_set_frame", "esfm_bank_commit", "esfm_match_all_pairs",
                 "esfm_match_pairs", "esfm_match_pair", "esfm_match_descriptors", "esfm_results_pair", "esfm_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import easysfm_b200 as esfm
    path = esfm.library_path()
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/esfm_match.h but not exported by {path}"


def test_python_binding_table_matches_header():
    from easysfm_b200 import capi
    assert sorted(capi.SIGNATURES) == declared_symbols()


def test_abi_version_and_struct_layout():
    import easysfm_b200 as esfm
    lib = esfm.load_library()
    assert lib.esfm_abi_version() == 2
    assert esfm.DMATCH_DTYPE.itemsize == 16          # layout of cv::DMatch
    assert [esfm.DMATCH_DTYPE.fields[k][1] for k in ("queryIdx", "trainIdx", "imgIdx", "distance")] == [0, 4, 8, 12]


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product path must fail loudly, not fall back."""
    import easysfm_b200 as esfm
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(esfm.EsfmError) as e:
        esfm.Context(0)
    assert "no CPU fallback" in str(e.value)
    lib = esfm.load_library()
    assert lib.esfm_bank_create(None, 0, 1, None) != 0   # NULL arguments are rejected, never dereferenced


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "easysfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
