"""The reference's known-answer tests (scan_block.rs:1908-2168, lib.rs:8-35) replayed through the C ABI of
the *emulated* build: same device source as the product (ba_kernel.cuh) executed by the fiber SIMT
emulator, same host runtime. Runs without a GPU. The CUDA build repeats these in test_gpu_parity.py."""
import ctypes as C

import pytest

import backend
import golden_cases as G
from block_aligner_b200 import api


@pytest.fixture(scope="module")
def env():
    lib = backend.emu_lib()
    return lib, api.Aligner(lib)


def test_golden_batch_api(env):
    G.run_batch_golden(*env)


def test_golden_legacy_api(env):
    G.run_legacy_golden(env[0])
