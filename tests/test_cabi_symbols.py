"""The product library loads on a machine without a GPU and exports every symbol that
include/block_aligner_b200.h declares; creating an aligner without a device fails loudly
(no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from block_aligner_b200 import api  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(api.DEFAULT_LIB):
        import __graft_entry__
        __graft_entry__.build()
    return api.Library()


def test_all_declared_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "block_aligner_b200.h")).read()
    declared = set(re.findall(r"\b(ba_[a-z0-9_]+)\s*\(", hdr))
    declared |= set(api.PART1_FUNCTIONS) | set(api.PART1_DATA)
    assert set(api.PART2_FUNCTIONS) <= declared
    assert len(api.PART1_FUNCTIONS) == 54 and len(api.PART1_DATA) == 12     # c/block_aligner.h:140-162, 169-562
    for name in sorted(declared):
        assert hasattr(lib.L, name), f"{name} is declared in the header but not exported"


def test_struct_layouts_match_reference_header():
    # c/block_aligner.h:90-134
    assert C.sizeof(api.AlignResult) == 24 and api.AlignResult.query_idx.offset == 8
    assert C.sizeof(api.OpLen) == 16 and api.OpLen.len.offset == 8
    assert C.sizeof(api.Gaps) == 2 and C.sizeof(api.SizeRange) == 16


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.BlockAlignerError, match="no CPU fallback|CUDA"):
        api.Aligner(lib)
