"""CPU: the C-ABI library loads, exports every symbol include/fba_pomdp_b200.h declares, and
refuses to run without a GPU (no fallback). No compute calls here."""
import ctypes as C
import os
import re

import pytest

import fba_pomdp_b200 as fba
from fba_pomdp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fba_pomdp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fba_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    for name in _declared_symbols():
        assert hasattr(L, name), name


def test_struct_layouts_match_header_sizes():
    # fba_model_desc: 5 ints + 32 feature ints + 2 ints + 32 ints + 8 doubles + 4 ptrs + 2 ints
    # + 4 ints + ptr + double + ptr, natural alignment
    assert C.sizeof(capi.ModelDesc) % 8 == 0
    assert C.sizeof(capi.Rng) == 48


def test_no_gpu_means_no_context():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fba.FbaError) as e:
        fba.Context(0)
    assert e.value.status == capi.ERR_NO_DEVICE


def test_constructor_errors_mirror_reference():
    # BAImportanceSampling.cpp:19-22, BARejectionSampling.cpp:13-16, Reinvigorating…cpp:43-48
    with pytest.raises(fba.FbaError, match="cannot initiate BAImportanceSampling with n 0"):
        fba.BAImportanceSampling(0)
    with pytest.raises(fba.FbaError, match="cannot initiate RejectionSampling with n = 0"):
        fba.BARejectionSampling(0)
    with pytest.raises(fba.FbaError, match="resample size of < 1"):
        fba.ReinvigoratingRejectionSampling(10, 0, 0)


def test_structure_belief_constructor_errors_mirror_reference():
    # CheatingReinvigoration.cpp:34-44, StructureIncubatorSampling.cpp:29-40, NestedBelief.cpp:19-26
    from fba_pomdp_b200.structure_beliefs import CheatingReinvigoration, NestedBelief, StructureIncubatorSampling
    with pytest.raises(fba.FbaError, match="CheatingReinvigoration::cannot initiate belief of size < 1"):
        CheatingReinvigoration(0, 4, -1.0)
    with pytest.raises(fba.FbaError, match="resample_threshold >= 0"):
        CheatingReinvigoration(8, 4, 0.5)
    with pytest.raises(fba.FbaError, match="Cannot initiate Incubator belief update with size < 1"):
        StructureIncubatorSampling(8, 0, 0.5, 0)
    with pytest.raises(fba.FbaError, match="must initiate with 1 < threshold <= 0"):
        StructureIncubatorSampling(8, 2, 1.5, 0)
    with pytest.raises(fba.FbaError, match="NestedBelief: cannot initiate with filter size < 1"):
        NestedBelief(3, 0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fba-pomdp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "pyref" not in text, f
