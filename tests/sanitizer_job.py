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
print("SANITIZER_JOB mismatches:", bad)
sys.exit(1 if bad else 0)
