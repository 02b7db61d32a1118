"""N > 1 host path on CPU: two gloo ranks shard a batch, align their shards (emulated build), gather, and
rank 0 compares with the oracle. The data path has no collective; only the result gather does."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from block_aligner_b200 import distributed  # noqa: E402


def test_shard_bounds_balanced_and_ordered():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 5000, size=1000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 3, 8):
        b = distributed.shard_bounds(off, off, world)
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all() and len(b) == world + 1
        work = np.add.reduceat(2 * lens + 64, b[:-1][np.diff(b) > 0])
        assert work.max() <= 1.2 * work.mean() + 2 * 5000


def test_two_ranks_gloo():
    import backend
    backend.emu_lib()   # build once, before the ranks race for it
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "DIST_OK world=2" in out.stdout
