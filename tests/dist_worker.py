"""Worker for tests/test_distributed.py: world_size ranks over gloo, each aligning its shard through the
emulated (CPU) build of the library; rank 0 checks the gathered results against the oracle."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import backend  # noqa: E402
import parity  # noqa: E402
from block_aligner_b200 import api, distributed, workloads  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = backend.emu_lib()
    al = api.Aligner(lib)
    w = dict(workloads.WORKLOADS["C2_nanopore_xdrop_10k"])
    w["gen"] = workloads.params(alphabet=0, len_dist=0, len_min=200, len_max=1500, sub_rate=0.04, ins_rate=0.04,
                                del_rate=0.04, long_indel_mean=1.0, long_indel_len=40.0, suffix_len=100)
    n = 41   # not divisible by the world size on purpose
    qa, qo, ra, ro = workloads.generate(w["gen"], n, stream=w["stream"])
    flags = w["flags"] | api.TRACE
    cfg = al.config(w["scoring"], workloads.matrix_of(lib, w), w["gaps"], w["size"], w["x_drop"], flags, True)
    res, cigs = distributed.align_sharded(al, cfg, qa, qo, ra, ro, want_cigars=True)
    bounds = distributed.shard_bounds(qo, ro, world)
    assert bounds[0] == 0 and bounds[-1] == n and (np.diff(bounds) >= 0).all()
    if rank == 0:
        exp = parity.oracle_batch(w["scoring"], workloads.matrix_of(lib, w), w["gaps"], w["size"], w["x_drop"], flags, True,
                                  qa, qo, ra, ro)
        assert (res == exp[0]).all(), "gathered results differ from the oracle"
        assert cigs == [api.runs_to_string(c) for c in exp[2]]
        print(f"DIST_OK world={world} shards={np.diff(bounds).tolist()}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
