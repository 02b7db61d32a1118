"""The Rust surface mirrored in C++ (include/block_aligner_b200.hpp): the reference's own unit tests, restated in
tests/cpp/test_cpp_api.cpp, run through it. CPU: against the emulated library (the device source compiled for the
host, test infrastructure). GPU: against the product library."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = os.path.join(ROOT, "tests", "cpp", "test_cpp_api.cpp")


def _build(out, libdir, libname):
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
                           "-L", libdir, f"-l{libname}", f"-Wl,-rpath,{libdir}"])


def test_reference_unit_tests_through_cpp_api_emulated(tmp_path):
    import backend
    backend.emu_lib()      # builds tests/emu/libba_emu.so when needed
    exe = str(tmp_path / "cpp_emu")
    _build(exe, os.path.join(ROOT, "tests", "emu"), "ba_emu")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("0 failed")


@pytest.mark.gpu
def test_reference_unit_tests_through_cpp_api_gpu(tmp_path):
    pkg = os.path.join(ROOT, "block_aligner_b200")
    if not os.path.exists(os.path.join(pkg, "libblock_aligner_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    exe = str(tmp_path / "cpp_gpu")
    _build(exe, pkg, "block_aligner_b200")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("0 failed")
