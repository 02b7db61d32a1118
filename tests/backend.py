"""Which copy of the C-ABI library a test talks to.

* emu_lib():  tests/emu/libba_emu.so -- the device source + host runtime compiled for the CPU with the
              fiber SIMT emulator. Test infrastructure; lets `-m "not gpu"` tests check parity.
* cuda_lib(): block_aligner_b200/libblock_aligner_b200.so -- the product, needs a GPU.
"""
import os
import subprocess

from block_aligner_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_EMU = None
_CUDA = None


def emu_lib():
    global _EMU
    if _EMU is None:
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.check_call(["make", "-s", "-C", d])
        _EMU = api.Library(os.path.join(d, "libba_emu.so"))
    return _EMU


def cuda_lib():
    global _CUDA
    if _CUDA is None:
        _CUDA = api.Library()
    return _CUDA
