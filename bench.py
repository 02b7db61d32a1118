#!/usr/bin/env python3
"""bench.py -- the hot path on BASELINE.json's workloads.

One "step" = one pass of the alignment kernel over one batch of synthetic (query, reference) pairs.
Headline workload = BASELINE.json configs[1]: "Nanopore X-drop" -- 100 000 DNA pairs of ~10 kbp, NW1,
gaps -2/-1, x_drop 50, block 32..=256 (block_aligner_b200/workloads.py). Under torchrun
every rank aligns its own 100 k-pair shard (weak scaling, no data-path collective: pairs are independent).
At N = 1 the line also carries `configs`: the other four BASELINE.json workloads (C1, C3, C4, C5), each with
its own value / e2e / roofline / cpu_baseline / parity, measured with fewer steps.

metric   GCUPS = computed DP cells / second / 1e9, cells = sum of h*w over every rectangle the adaptive
         algorithm computes (path-determined: identical on GPU and CPU when results are bit-exact)
value    kernel time only, inputs resident in HBM (CUDA events on the kernel's stream, max over ranks)
e2e      through the C ABI (ba_align_batch / _cigar / _pssm) with host buffers: H2D of raw bytes +
         convert/pad + align + D2H, every step
parity   the results (and CIGARs) the LAST e2e step delivered, diffed against the CPU oracle's results for the
         pairs the cpu_baseline leg aligned; a mismatch makes the run exit non-zero
roofline integer-ALU bound (max-plus DP on i16 values; not HBM, not tensor): achieved = cells/s x algorithmic
         ops per cell vs. the add/max issue peak measured on this GPU
cpu_baseline  the CPU oracle (C++ restatement of the reference's AVX2 path) on all host cores, bounded sample

--impl reference times that CPU path alone (same metric/config) -- the reference is Rust and cannot be
built in this image, so the "reference arm" is its AVX2-intrinsic restatement in oracle/ba_oracle.cpp.
"""
import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL writes its "NCCL version ..." banner to stdout; rank 0 must print exactly one JSON line there
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

# Native libraries (the NCCL version banner, for one) write to file descriptor 1. The contract is ONE JSON line on
# stdout, so fd 1 is pointed at stderr for the whole run and the line goes to a private copy of the real stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

HEADLINE = "C2_nanopore_xdrop_10k"
OTHER_CONFIGS = ["C1_rand_scan_dna1k", "C3_uniclust_protein_global", "C4_seq_to_profile_xdrop", "C5_longread_trace_50k"]
OPS_PER_CELL = {0: 10, 2: 13, 1: 16, 3: 19}   # flags -> algorithmic i16 ops per cell
RES_DT = np.dtype([("score", np.int32), ("q", np.uint64), ("r", np.uint64)], align=True)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def log(msg):
    sys.stderr.write(f"[bench] {msg}\n")
    sys.stderr.flush()


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag = gpu_index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": reasons, "samples": len(self.rows)}


def load_workload(name):
    from block_aligner_b200 import workloads
    return workloads.WORKLOADS[name]


def pin(a):
    """numpy array -> (view of a pinned copy, owner)"""
    import torch
    t = torch.empty(a.shape, dtype={np.dtype("uint8"): torch.uint8, np.dtype("uint64"): torch.int64}[a.dtype]).pin_memory()
    v = t.numpy().view(a.dtype)
    v[...] = a
    return v, t


def gen_shard(w, n, first, pinned):
    """Synthetic shard -> (q_arena, q_off, r_arena, r_off, owners); arenas live in pinned host memory when asked."""
    from block_aligner_b200 import workloads
    arrs = workloads.generate(w["gen"], n, first=first, seed=1234, stream=w["stream"])
    if not pinned:
        return (*arrs, None)
    out, keep = [], []
    for a in arrs:
        v, t = pin(a)
        out.append(v)
        keep.append(t)
    return (*out, keep)


def oracle_matrix(w):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ora
    if w["matrix"] is None:
        return None
    if isinstance(w["matrix"], str):
        return ora.nw1() if w["matrix"] == "NW1" else ora.builtin(w["matrix"])
    return ora.nuc_matrix(*w["matrix"])


def cpu_sample(w, n_pairs_hint, target_s, first=0, pssm_seed=1234, keep_results=False):
    """Time the oracle on a bounded sample (the first pairs of the workload) with all host threads -> dict for
    cpu_baseline. TRACE workloads are timed with the trace and the CIGAR walk, like the GPU path. With keep_results
    the oracle's results (and CIGARs) of the sample ride along for the parity check."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ora
    import parity
    from block_aligner_b200 import api, workloads
    threads = ora.lib().ora_hw_threads()
    matrix = oracle_matrix(w)
    prof = w["scoring"] == api.SCORING_PROFILE
    eq = bool(w.get("cigar_eq")) and bool(w["flags"] & api.TRACE)

    def run(n, first_pair):
        qa, qo, ra, ro = workloads.generate(w["gen"], n, first=first_pair, seed=1234, stream=w["stream"])
        profs = parity.make_ora_profiles(ra, ro, w["size"][1], -10, -1, pssm_seed) if prof else None
        res, cells, cigs = parity.oracle_batch(w["scoring"], matrix, w["gaps"], w["size"], w["x_drop"], w["flags"], eq,
                                               qa, qo, ra, ro, profiles=profs, threads=threads, cigars_as_arrays=True)
        dt = ora.lib().ora_last_batch_seconds()   # wall clock around the align (+ cigar) calls only
        return dt, int(cells.sum()), n, (res, cigs)
    pilot = int(min(n_pairs_hint, max(threads * 4, 32)))
    dt, cells, n, kept = run(pilot, first)
    for _ in range(3):     # the pilot over-estimates the time per pair (per-thread Block set-up): grow until near the target
        if n >= n_pairs_hint or dt >= 0.6 * target_s:
            break
        n2 = int(min(n_pairs_hint, max(n + 1, n * target_s / max(dt, 1e-9))))
        dt, cells, n, kept = run(n2, first)
    out = {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
           "sample": f"first {n} pairs of the workload, oracle/libba_oracle.so (C++ restatement of the reference AVX2 path; "
                     f"the Rust original cannot be built here), {threads} threads, {dt:.2f} s wall"
                     + (", trace + CIGAR included" if w["flags"] & api.TRACE else ""),
           "alignments_per_s": n / dt, "seconds": dt, "pairs": n, "cells": cells}
    if keep_results:
        out["_results"] = kept
    return out


def parity_of(cb, out, cig, path):
    """Diff what the last e2e step delivered against the oracle's results of the cpu_baseline sample."""
    res_e, cigs_e = cb["_results"]
    n = len(res_e)
    got = np.stack([out["score"][:n].astype(np.int64), out["q"][:n].astype(np.int64), out["r"][:n].astype(np.int64)], axis=1)
    bad = (got != res_e).any(axis=1)
    cig_checked = 0
    if cig is not None and cigs_e is not None:
        runs = cig["runs"].numpy().view(np.uint32)
        for k in range(n):
            o, ln = int(cig["off"][k]), int(cig["len"][k])
            e = cigs_e[k]
            if ln != len(e) or not (runs[o:o + ln].astype(np.uint64) == e).all():
                bad[k] = True
        cig_checked = n
    return {"pairs_checked": int(n), "mismatches": int(bad.sum()), "cigars_checked": cig_checked, "path": path,
            "against": "CPU oracle results of the cpu_baseline sample (score, query_idx, reference_idx"
                       + (", CIGAR runs)" if cig_checked else ")")}


def newest_traffic(tag, n):
    """DRAM bytes per launch from the newest committed `ncu --set full` capture of this workload (per pair x pairs)."""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_*kernel_summary.json"))):
        try:
            for c in json.load(open(path))["captures"]:
                if tag in c["capture"].split() and "traceback" not in c["capture"] and c.get("dram_bytes_per_pair"):
                    best = (c["dram_bytes_per_pair"] * n, os.path.relpath(path, ROOT))
        except Exception:
            pass
    return best or (None, None)


def measure(name, n, steps, warmup, ctx, cpu_seconds, want_cpu, pageable_once=False):
    """One workload on this rank's GPU -> result dict (rank 0; None elsewhere). Collective: every rank calls it."""
    import torch
    from block_aligner_b200 import api, workloads
    lib, al, rank, world, barrier, dist = ctx["lib"], ctx["al"], ctx["rank"], ctx["world"], ctx["barrier"], ctx["dist"]
    w = load_workload(name)
    n = n or w["n"]
    t_gen = time.perf_counter()
    qa, qo, ra, ro, keep = gen_shard(w, n, first=rank * n, pinned=True)
    matrix = workloads.matrix_of(lib, w)
    cfg = al.config(w["scoring"], matrix, w["gaps"], w["size"], w["x_drop"], w["flags"], bool(w.get("cigar_eq")))
    profiles = None
    pssm_seed = 1234 + rank
    if w["scoring"] == api.SCORING_PROFILE:
        # raw PSSM rows in pinned host memory; the padded profiles are built on the device (ba_batch_upload_pssm)
        profiles = workloads.make_pssm_batch(lib, ra, ro, seed=pssm_seed, pinned=True)
    log(f"{name}: {n} pairs generated in {time.perf_counter() - t_gen:.1f} s")

    # ---- value: kernel only, inputs resident ----
    batch = al.upload(cfg, qa, qo, ra, ro, profiles)
    for _ in range(warmup):
        batch.run()
    sampler = ClockSampler(ctx["local_rank"])
    sampler.start()
    barrier()
    kernel_ms, value_launches = [], 0
    t0 = time.perf_counter()
    for _ in range(steps):
        st = batch.run()                       # CUDA events around the kernel on its own stream, synchronised
        kernel_ms.append(st.kernel_ms)
        value_launches += int(st.kernel_launches)
    barrier()
    wall = time.perf_counter() - t0
    batch.download()
    tot = batch.total_stats()
    cells_step, n_failed = int(tot.cells), int(tot.n_failed)
    batch.free()
    dev_s = sum(kernel_ms) / 1e3

    # ---- e2e: host buffers -> results, copies inside the timed region ----
    out = np.zeros(n, dtype=RES_DT)
    st = api.BaStats()
    trace = bool(w["flags"] & api.TRACE)
    cig = None
    if trace:     # the CIGARs come back too, straight into a pinned caller buffer (ba_align_batch_cigar)
        cap = int(0.3 * (qa.nbytes + ra.nbytes)) + 16 * n
        cig = dict(cap=cap, runs=torch.empty(cap, dtype=torch.int32).pin_memory(), off=np.zeros(n, dtype=np.uint64),
                   len=np.zeros(n, dtype=np.uint32), used=C.c_size_t())
    path = "ba_align_batch_cigar" if trace else ("ba_align_batch_pssm" if profiles is not None else "ba_align_batch")
    # Nucleotide workloads: the end-to-end leg hands the library BAM-style 4-bit codes (BA_INPUT_NUC4, two bases per byte;
    # packed once, outside the timed region: it is the caller's input format) -- half the bytes cross PCIe. The same call
    # with one ASCII byte per base is timed next to it (e2e.ascii_input).
    nuc4 = w["scoring"] == api.SCORING_NUC and not os.environ.get("BA_BENCH_NO_NUC4")
    cfg_main = cfg
    if nuc4:
        pq, _ = api.pack_nuc4(lib, qa, qo)
        pr, _ = api.pack_nuc4(lib, ra, ro)
        (pq, kq), (pr, kr) = pin(pq), pin(pr)
        keep += [kq, kr]
        cfg_main = al.config(w["scoring"], matrix, w["gaps"], w["size"], w["x_drop"], w["flags"] | api.INPUT_NUC4, bool(w.get("cigar_eq")))

    # One result buffer per timed call, zeroed (and thereby touched) before the clock starts: every call provably writes its
    # own results, and the timed region holds the library call only (clearing a 24 MB result array of C3 inside it cost 4 ms)
    obufs = [np.zeros(n, dtype=RES_DT) for _ in range(steps)]
    for b_ in obufs:
        b_.fill(0)

    def e2e_once(q_arena, r_arena, c=None, out=out):
        c = c or cfg
        if trace:
            lib.check(lib.L.ba_align_batch_cigar(al.h, C.byref(c), n, q_arena.ctypes.data, qo.ctypes.data, r_arena.ctypes.data,
                                                 ro.ctypes.data, out.ctypes.data, cig["runs"].data_ptr(), cig["cap"],
                                                 cig["off"].ctypes.data, cig["len"].ctypes.data, C.byref(cig["used"]), C.byref(st)))
        elif profiles is not None:
            lib.check(lib.L.ba_align_batch_pssm(al.h, C.byref(c), n, q_arena.ctypes.data, qo.ctypes.data, C.byref(profiles.c),
                                                out.ctypes.data, C.byref(st)))
        else:
            lib.check(lib.L.ba_align_batch(al.h, C.byref(c), n, q_arena.ctypes.data, qo.ctypes.data, r_arena.ctypes.data,
                                           ro.ctypes.data, out.ctypes.data, C.byref(st)))
    e2e_error, e2e_launches, pageable_s, ascii_s = None, 0, None, None
    try:
        if nuc4:      # the ASCII-input variant first (same entry point, one byte per base)
            e2e_once(qa, ra)
            barrier()
            t1 = time.perf_counter()
            for i_ in range(steps):
                e2e_once(qa, ra, None, obufs[i_])
            barrier()
            ascii_s = time.perf_counter() - t1
            ascii_out = obufs[steps - 1].copy()
            for b_ in obufs:
                b_.fill(0)
        mq, mr = (pq, pr) if nuc4 else (qa, ra)
        for _ in range(min(warmup, 2)):
            e2e_once(mq, mr, cfg_main)
        barrier()
        t1 = time.perf_counter()
        for i_ in range(steps):
            e2e_once(mq, mr, cfg_main, obufs[i_])
            e2e_launches += int(st.kernel_launches)      # alignment + convert/pad (+ profile build) launches of this call
        barrier()
        e2e_s = time.perf_counter() - t1
        e2e_out = obufs[steps - 1].copy()
        out[:] = e2e_out
        if nuc4 and not (ascii_out == e2e_out).all():
            e2e_error = "4-bit and ASCII inputs returned different results"
        if pageable_once and profiles is None:
            # the same call from ordinary (pageable) caller memory, once: what a drop-in user who does not pin gets
            qp, rp = np.array(qa, copy=True), np.array(ra, copy=True)
            e2e_once(qp, rp)
            out[:] = 0
            t2 = time.perf_counter()
            e2e_once(qp, rp)
            e2e_once(qp, rp)
            pageable_s = (time.perf_counter() - t2) / 2
            if not (out == e2e_out).all():
                e2e_error = "pageable-input run returned different results"
            del qp, rp
    except api.BlockAlignerError as e:      # reported, never hidden: the kernel-only number is still valid
        e2e_error = str(e)
        barrier()
        e2e_s = float("inf")
        e2e_out = out
    sampler.stop_flag = True
    seq_bytes = int(pq.nbytes + pr.nbytes) if nuc4 else int(qa.nbytes + (profiles.nbytes() if profiles is not None else ra.nbytes))
    h2d = int(seq_bytes + 2 * qo.nbytes + n * (8 + 8 + 4 + 4 + 4))
    d2h = int(n * 48) + (int(cig["used"].value) * 4 if trace else 0)

    # ---- max over ranks ----
    vals = torch.tensor([dev_s, e2e_s, wall, ascii_s or 0.0], dtype=torch.float64, device="cuda")
    cells_t = torch.tensor([cells_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells_t, op=dist.ReduceOp.SUM)
    dev_s, e2e_s, wall, ascii_s = [float(x) for x in vals.tolist()]
    cells_all = float(cells_t.item())
    if rank != 0:
        return None
    gcups = cells_all * steps / dev_s / 1e9
    e2e_gcups = cells_all * steps / e2e_s / 1e9

    # the packed 2 x i16 path serves every batch except the extended modes; its ceiling is the s16x2 issue rate
    packed = not (w["flags"] & (api.LOCAL_START | api.FREE_QUERY_START_GAPS | api.FREE_QUERY_END_GAPS)) \
        and not os.environ.get("BA_NO_PACKED") and (profiles is None or ctx["packed_profiles"])
    peak_gops = ctx["peak_s16x2"] if packed else ctx["peak_s32"]
    ops = OPS_PER_CELL[w["flags"] & 3]
    achieved = cells_step * steps / (sum(kernel_ms) / 1e3) * ops / 1e9     # this rank's kernel
    peaks = ctx["peaks"]
    alg_bytes = float(qa.nbytes + (profiles.nbytes() if profiles is not None else ra.nbytes) + n * 48)
    traffic, traffic_src = newest_traffic(name.split("_")[0], n)
    res = {
        "metric": "GCUPS (computed DP cells / s / 1e9)", "value": gcups, "unit": "GCUPS", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": dev_s / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "i16", "data": "synthetic",
        "config": {"workload": f"{name}: {n} pairs/GPU, {w['matrix']} gaps {w['gaps']}, block {w['size'][0]}..={w['size'][1]}, "
                               f"x_drop {w['x_drop']}, flags {w['flags']} (1=TRACE, 2=X_DROP)",
                   "pairs_per_gpu": n, "l2": "inputs (%.2f GB/GPU) are larger than L2; no flush needed" % (alg_bytes / 1e9)
                   if alg_bytes > 200e6 else "inputs (%.1f MB) fit in L2: launch-bound plumbing case, L2 not flushed" % (alg_bytes / 1e6),
                   "parallelism": f"pairs sharded over {world} GPU(s), no collective on the data path"},
        "alignments_per_s": n * world * steps / dev_s,
        "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / steps * 1e3, "alignments_per_s": n * world * steps / e2e_s, "entry_point": path,
                "caller_buffers": "pinned host memory",
                "input_format": "BAM 4-bit base codes, two per byte (BA_INPUT_NUC4)" if nuc4 else "one byte per residue"},
        # launches of this library's kernels inside the two timed regions (counted by the library, BaStats)
        "gpu_launches": value_launches + e2e_launches,
        "gpu_launches_detail": {"value_region": value_launches, "e2e_region": e2e_launches},
        "roofline": {"bound": "int_alu", "achieved": achieved / 1e3, "peak": peak_gops / 1e3, "unit": "Tiop/s",
                     "frac": achieved / peak_gops if peak_gops else None, "traffic": traffic, "traffic_source": traffic_src,
                     "ops_per_cell": ops, "arith": "s16x2 (two cells per DPX instruction)" if packed else "s32 (one cell per DPX instruction)",
                     "peak_s32": ctx["peak_s32"] / 1e3, "peak_s16x2": ctx["peak_s16x2"] / 1e3,
                     "peak_source": "ba_measure_int_peak[_packed] (DPX add-max / max3 issue rate measured on this GPU, "
                                    "for the arithmetic the kernel uses)",
                     "hbm": {"achieved": alg_bytes / (sum(kernel_ms) / steps / 1e3) / 1e9, "peak": peaks.get("hbm_gbs", 6650.0),
                             "unit": "GB/s", "algorithmic_bytes_per_step": alg_bytes,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}},
        "clocks": sampler.summary(), "n_failed_pairs": n_failed, "wall_s_timed_region": wall,
    }
    if nuc4 and ascii_s:
        res["e2e"]["ascii_input"] = {"value": cells_all * steps / ascii_s / 1e9, "unit": "GCUPS", "ms_per_step": ascii_s / steps * 1e3,
                                     "h2d_bytes_per_step": int(qa.nbytes + ra.nbytes + 2 * qo.nbytes + n * 28),
                                     "note": "same entry point, one ASCII byte per base (the reference's own input format)"}
    if pageable_s is not None:
        res["e2e"]["pageable_caller_buffers"] = {"value": cells_step / pageable_s / 1e9, "unit": "GCUPS", "ms_per_step": pageable_s * 1e3,
                                                 "steps": 2, "note": "same call, inputs in ordinary malloc'ed memory (this rank only; the library stages them through pinned buffers with four copy threads)"}
    if e2e_error:
        res["e2e"] = {"value": None, "unit": "GCUPS", "error": e2e_error}
    res["_e2e_out"], res["_cig"] = e2e_out, cig
    if want_cpu:
        cb = cpu_sample(w, n, cpu_seconds, first=0, pssm_seed=pssm_seed, keep_results=True)
        res["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "alignments_per_s")}
        if not e2e_error:
            res["parity"] = parity_of(cb, e2e_out, cig, path)
    del keep
    return res


def strong_scaling(lib, al, name, n, n_gpus, steps, cb):
    """ONE batch of the workload's full size over n_gpus GPUs in ONE call (ba_align_batch_multi): host buffers in, results
    in the caller's order out, wall clock around the call. Run by rank 0 alone after the per-rank (weak-scaling) part."""
    from block_aligner_b200 import api, workloads
    w = load_workload(name)
    n = n or w["n"]
    lib.check(lib.L.ba_trim(al.h))
    qa, qo, ra, ro, keep = gen_shard(w, n, first=0, pinned=True)
    pq, _ = api.pack_nuc4(lib, qa, qo)
    pr, _ = api.pack_nuc4(lib, ra, ro)
    (pq, kq), (pr, kr) = pin(pq), pin(pr)
    matrix = workloads.matrix_of(lib, w)
    cfg = al.config(w["scoring"], matrix, w["gaps"], w["size"], w["x_drop"], w["flags"] | api.INPUT_NUC4, False)
    out = np.zeros(n, dtype=RES_DT)
    st = api.BaStats()

    def once():
        lib.check(lib.L.ba_align_batch_multi(None, n_gpus, C.byref(cfg), n, pq.ctypes.data, qo.ctypes.data, pr.ctypes.data, ro.ctypes.data,
                                             out.ctypes.data, C.byref(st)))
    once()
    once()
    t0 = time.perf_counter()
    for _ in range(steps):
        out[:] = 0
        once()
    dt = (time.perf_counter() - t0) / steps
    res = {"n_gpus": n_gpus, "pairs": n, "entry_point": "ba_align_batch_multi", "value": int(st.cells) / dt / 1e9, "unit": "GCUPS",
           "ms_per_step": dt * 1e3, "steps": steps, "scaling": "strong",
           "note": "one batch, one call, all GPUs: contiguous shards of equal sum(|q|+|r|), one host thread per GPU, "
                   "results written in the caller's order; wall clock incl. H2D (4-bit input) and D2H"}
    if cb is not None:
        res["parity"] = parity_of(cb, out, None, "ba_align_batch_multi")
    lib.L.ba_multi_release()
    del keep, kq, kr
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the workload's full size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json workloads (N = 1 adds them by default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        w = load_workload(args.workload)
        n = args.pairs or w["n"]
        # the reference's CPU implementation of the path (restated), all host threads, bounded sample per step
        vals, cb = [], None
        for s in range(args.warmup + args.steps):
            cb = cpu_sample(w, n, max(2.0, args.cpu_seconds / 2), first=0)
            if s >= args.warmup:
                vals.append(cb)
        tot_cells = sum(v["cells"] for v in vals)
        tot_s = sum(v["seconds"] for v in vals)
        val = tot_cells / tot_s / 1e9
        line = {"impl": "reference", "metric": "GCUPS (computed DP cells / s / 1e9)", "value": val, "unit": "GCUPS",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_s / max(len(vals), 1) * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {n} pairs/GPU, {w['matrix']} gaps {w['gaps']}, block {w['size'][0]}..={w['size'][1]}, "
                                       f"x_drop {w['x_drop']}, flags {w['flags']} (1=TRACE, 2=X_DROP)", "pairs_per_gpu": n},
                "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cb["cores"], "kind": "port", "sample": cb["sample"]},
                "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "alignments_per_s": sum(v["pairs"] for v in vals) / tot_s}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from block_aligner_b200 import api
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = api.Library()
    al = api.Aligner(lib, local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ctx = dict(lib=lib, al=al, rank=rank, world=world, local_rank=local_rank, barrier=barrier, dist=dist, peaks=peaks,
               peak_s32=al.int_peak_gops(False), peak_s16x2=al.int_peak_gops(True),
               packed_profiles=not os.environ.get("BA_NO_FAST"))
    want_cpu = not args.no_cpu_baseline
    line = measure(args.workload, args.pairs, args.steps, args.warmup, ctx, args.cpu_seconds, False, pageable_once=(world == 1))
    # every rank is past its last collective: ranks > 0 are done, rank 0 goes on alone (CPU leg, other configs)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    bad = 0
    if want_cpu:
        # the CPU figure "in the same run" (north_star) at every N: rank 0's host cores, after the GPU timing
        w = load_workload(args.workload)
        cb = cpu_sample(w, args.pairs or w["n"], args.cpu_seconds, first=0, pssm_seed=1234, keep_results=True)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "alignments_per_s")}
        if line["e2e"].get("value") is not None:
            line["parity"] = parity_of(cb, line.pop("_e2e_out"), line.pop("_cig"), line["e2e"]["entry_point"])
            bad += line["parity"]["mismatches"]
    line.pop("_e2e_out", None)
    line.pop("_cig", None)
    if args.workload == HEADLINE and not os.environ.get("BA_BENCH_NO_STRONG"):
        try:
            line["strong_scaling"] = strong_scaling(lib, al, args.workload, args.pairs, max(1, args.gpus), max(2, args.steps // 2),
                                                    cb if want_cpu else None)
            bad += line["strong_scaling"].get("parity", {}).get("mismatches", 0)
        except Exception as e:
            line["strong_scaling"] = {"error": repr(e)}
    if world == 1 and not args.no_configs and args.workload == HEADLINE and not args.pairs:
        ctx1 = dict(ctx, world=1)
        line["configs"] = {}
        for name in OTHER_CONFIGS:
            try:
                r = measure(name, 0, max(2, args.steps // 2), 3, ctx1, max(4.0, args.cpu_seconds / 2), want_cpu)
                for k in ("_e2e_out", "_cig"):
                    r.pop(k, None)
                keep = ("value", "unit", "ms_per_step", "steps", "warmup", "alignments_per_s", "e2e", "roofline", "cpu_baseline",
                        "parity", "gpu_launches", "n_failed_pairs", "config", "clocks")
                line["configs"][name] = {k: r[k] for k in keep if k in r}
                bad += r.get("parity", {}).get("mismatches", 0)
                al.lib.check(lib.L.ba_trim(al.h))
            except Exception as e:       # a failing side config is reported in the line, the headline stays valid
                line["configs"][name] = {"error": repr(e)}
                bad += 1
    emit(line)
    if bad:
        log(f"PARITY FAILURE: {bad} mismatching pair(s) / failed config(s)")
        sys.exit(1)


if __name__ == "__main__":
    main()
