"""The C ABI as a C program sees it. CPU part: examples/c_example.c (and, when the reference checkout is
present, the reference's own c/example.c against the reference's own header) compile and link against the
product library. GPU part: the example runs and prints what the CPU oracle computes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from block_aligner_b200 import api  # noqa: E402

PKG = os.path.join(ROOT, "block_aligner_b200")
EXE = os.path.join(ROOT, "examples", "c_example")


def _ensure_lib():
    if not os.path.exists(api.DEFAULT_LIB):
        import __graft_entry__
        __graft_entry__.build()


def _build(src, inc, out):
    subprocess.check_call(["gcc", "-O1", "-I", inc, src, "-o", out, "-L", PKG, "-lblock_aligner_b200",
                           f"-Wl,-rpath,{PKG}"])


def test_example_links_against_product_library():
    _ensure_lib()
    _build(os.path.join(ROOT, "examples", "c_example.c"), os.path.join(ROOT, "include"), EXE)
    assert os.path.exists(EXE)


def test_batch_example_on_emulated_library(tmp_path):
    """examples/batch_example.c (Part 2: ba_align_batch_cigar) against the emulated library: the doc-test pair of the
    reference (src/lib.rs:8-35) and a golden pair of src/scan_block.rs:2085-2094 come out right."""
    import backend
    backend.emu_lib()
    exe = str(tmp_path / "batch_example")
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "batch_example.c"),
                           "-o", exe, "-L", emu, "-lba_emu", f"-Wl,-rpath,{emu}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    assert lines[0] == "pair 0: score=7 idx=(24,21) cigar=2=6I16=3D"
    assert lines[1] == "pair 1: score=8 idx=(16,13) cigar=9=2I4=1I"


def test_reference_example_links_unchanged(tmp_path):
    ref = "/root/reference/c"
    if not os.path.exists(os.path.join(ref, "example.c")):
        pytest.skip("reference checkout not present on this machine")
    _ensure_lib()
    _build(os.path.join(ref, "example.c"), ref, str(tmp_path / "ref_example"))


@pytest.mark.gpu
def test_example_output_matches_oracle():
    import ora
    _ensure_lib()
    if not os.path.exists(EXE):
        _build(os.path.join(ROOT, "examples", "c_example.c"), os.path.join(ROOT, "include"), EXE)
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    b62 = ora.builtin("BLOSUM62")
    ob = ora.Block(8, 7, 32, ora.TRACE)
    q, r = ora.Padded(ora.AA, b"AAAAAAAA", 32), ora.Padded(ora.AA, b"AARAAAA", 32)
    s, qi, rj = ob.align(q, r, ora.AA, b62, (-11, -1), (32, 32), 0)
    assert lines[0] == f"global score={s} idx=({qi},{rj})"
    assert lines[1] == f"trace score={s} idx=({qi},{rj}) cigar={ob.cigar(qi, rj)}"
    prof = ora.Profile.new(7, 32, -1)
    for i in range(1, 8):
        for c in range(65, 91):
            prof.set(i, c, 1 if c == ord("A") else -1)
    for i in range(7):
        prof.set_gap_open_C(i, -10); prof.set_gap_close_C(i, 0); prof.set_gap_open_R(i, -10)
    ob2 = ora.Block(8, 7, 32, 0)
    s, qi, rj = ob2.align_profile(q, prof, (32, 32), 0)
    assert lines[2] == f"profile score={s} idx=({qi},{rj})"
