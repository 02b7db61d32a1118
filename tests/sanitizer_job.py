"""Small mixed job for compute-sanitizer (memcheck / racecheck / synccheck): every scoring kind and flag
combination on a handful of pairs, including the fast phase, grow/shrink and the chunked (>256) rectangles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from block_aligner_b200 import api, workloads  # noqa: E402

lib = api.Library()
al = api.Aligner(lib, 0)
P = workloads.params
noisy = dict(sub_rate=0.05, ins_rate=0.04, del_rate=0.04, long_indel_mean=1.5, long_indel_len=50.0)
bad = 0
for flags in (0, api.XDROP, api.TRACE, api.TRACE | api.XDROP):
    for size in ((32, 256), (16, 64), (128, 1024)):
        w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=size, x_drop=50, flags=flags, stream=11,
                 gen=P(alphabet=0, len_dist=0, len_min=300, len_max=1200, suffix_len=100, big_indel_prob=0.5, big_indel_min=60,
                       big_indel_max=200, **noisy))
        bad += parity.check_workload(lib, al, w, 24, seed=3 + flags)
w = dict(workloads.WORKLOADS["C3_uniclust_protein_global"]); bad += parity.check_workload(lib, al, w, 64)
w = dict(workloads.WORKLOADS["C4_seq_to_profile_xdrop"]); w["flags"] = api.TRACE | api.XDROP; bad += parity.check_workload(lib, al, w, 24)
# live borders in global memory (max block >= 1024 with a fast phase: FM 34 / 35), trace overflow pool, retry pass
for size in ((64, 2048), (32, 1024)):
    w = dict(scoring=api.SCORING_NUC, matrix=(2, -4), gaps=(-6, -2), size=size, x_drop=100, flags=api.TRACE | api.XDROP, stream=21,
             gen=P(alphabet=0, len_dist=0, len_min=300, len_max=2500, suffix_len=150, big_indel_prob=0.5, big_indel_min=80,
                   big_indel_max=300, **noisy))
    bad += parity.check_workload(lib, al, w, 16, seed=11)
for env in ({}, {"BA_NO_TRACE_POOL": "1"}, {"BA_TRACE_POOL_BYTES": "65536"}):
    os.environ.update(env)
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=(32, 512), x_drop=0, flags=api.TRACE, stream=31,
             gen=P(alphabet=0, len_dist=0, len_min=1500, len_max=2500, sub_rate=0.75, ins_rate=0.0, del_rate=0.0))
    bad += parity.check_workload(lib, al, w, 40)
    for k in env:
        del os.environ[k]
# profiles built on the device, reversed inputs
bad += parity.check_pssm(lib, al, 40, 77, True, (0, 0), False, size=(32, 256))
bad += parity.check_reversed(lib, al, 24, api.REV_QUERY | api.REV_REFERENCE)
print("SANITIZER_JOB mismatches:", bad)
sys.exit(1 if bad else 0)
