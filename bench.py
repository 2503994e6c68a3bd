#!/usr/bin/env python3
"""bench.py — Foldseek search hot path (k-mer prefilter + structurealign score phase) on B200.

One "step" = one pass of the hot path over one batch of synthetic queries against the staged target DB:
    fsb200_set_queries -> fsb200_prefilter -> fsb200_align      (libfsb200.so, include/fsb200.h)
metric  : query-structures/s (BASELINE.json), whole job, all ranks
value   : queries already resident (set_queries done before the timed region), no result download
e2e     : the same step through the C ABI with HOST buffers: query upload, search, download of hits + alignments
workload: BASELINE.json configs[1] — 1k synthetic queries (len 350) vs 100k-structure DB, prefilter+align
Multi-GPU (torchrun): rank 0 stages the target DB and broadcasts the staged arrays over NCCL; every rank searches its
own shard of 1k queries (weak scaling, no further collectives).
--impl reference: the reference's own CPU modules (oracle/_ref/refseek = unmodified reference sources) on all host
threads, same workload.
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NQ, NT, LEN, MAXSEQS = 1000, 100000, 350, 1000
PLANT = 0.1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def make_workload(rank, nq=NQ, nt=NT):
    """Target DB is identical on every rank (seed fixed); queries differ per rank."""
    from foldseek_b200 import synth
    rng = np.random.default_rng(20240922)
    taa = synth.random_codes(rng, nt, LEN, synth.BG_AA)
    tss = synth.random_codes(rng, nt, LEN, synth.BG_3DI)
    qrng = np.random.default_rng(777 + rank)
    qaa = synth.random_codes(qrng, nq, LEN, synth.BG_AA)
    qss = synth.random_codes(qrng, nq, LEN, synth.BG_3DI)
    # planted homologs: mutated copies of targets (substitutions only keep the length at 350)
    n_plant = int(nq * PLANT)
    src = qrng.integers(0, nt, size=n_plant)
    for k in range(n_plant):
        a, s = taa[src[k]].copy(), tss[src[k]].copy()
        m = qrng.random(LEN) < 0.3
        a[m] = qaa[k][m]
        m = qrng.random(LEN) < 0.3
        s[m] = qss[k][m]
        qaa[k], qss[k] = a, s
    return taa, tss, qaa, qss


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def _last_progress_seconds(text):
    """Elapsed time printed by the reference's progress bar for its last (query) loop, e.g. '] 1.00K 0s 337ms'."""
    best = None
    for m in re.finditer(r"\]\s+\S+\s+(?:(\d+)h\s+)?(?:(\d+)m\s+)?(\d+)s\s+(\d+)ms", text):
        best = int(m.group(1) or 0) * 3600 + int(m.group(2) or 0) * 60 + int(m.group(3)) + int(m.group(4)) / 1000.0
    return best


def run_reference(nq, steps, warmup, tmp=None, threads=None):
    """Times the reference CPU modules (prefilter + structurealign) on `nq` queries of the workload. Returns dict."""
    from foldseek_b200 import synth
    refseek = os.path.join(ROOT, "oracle", "_ref", "refseek")
    if not os.path.exists(refseek):
        return None
    threads = threads or os.cpu_count()
    taa, tss, qaa, qss = make_workload(0, nq=nq)
    d = tempfile.mkdtemp(prefix="fsb200_ref_", dir=tmp)
    try:
        synth.write_seq_db_fixed(os.path.join(d, "t"), taa)
        synth.write_seq_db_fixed(os.path.join(d, "t_ss"), tss)
        synth.write_seq_db_fixed(os.path.join(d, "q"), qaa)
        synth.write_seq_db_fixed(os.path.join(d, "q_ss"), qss)
        times = []
        for it in range(warmup + steps):
            for f in os.listdir(d):
                if f.startswith("pref") or f.startswith("aln"):
                    os.remove(os.path.join(d, f))
            t0 = time.time()
            out1 = subprocess.run([refseek, "prefilter", "q_ss", "t_ss", "pref", "-s", "9.5", "--max-seqs", str(MAXSEQS),
                                   "--comp-bias-corr-scale", "0.15", "--threads", str(threads)], cwd=d, capture_output=True, text=True)
            t1 = time.time()
            out2 = subprocess.run([refseek, "structurealign", "q", "t", "pref", "aln", "-e", "10", "--sort-by-structure-bits", "0",
                                   "--threads", str(threads)], cwd=d, capture_output=True, text=True)
            t2 = time.time()
            if out1.returncode != 0 or out2.returncode != 0:
                raise RuntimeError("refseek failed: " + out1.stdout[-500:] + out2.stdout[-500:])
            m = re.search(r"Time for index table init: (\d+)h (\d+)m (\d+)s (\d+)ms", out1.stdout)
            t_index = (int(m.group(1)) * 3600 + int(m.group(2)) * 60 + int(m.group(3)) + int(m.group(4)) / 1000.0) if m else 0.0
            # the query loops themselves (Debug::Progress prints their elapsed time when they finish): this is what a run
            # against a precomputed index (createindex stores the index table and the 3-mer matrices) spends per batch
            lp1, lp2 = _last_progress_seconds(out1.stdout), _last_progress_seconds(out2.stdout)
            times.append((lp1 if lp1 is not None else (t1 - t0) - t_index, lp2 if lp2 is not None else t2 - t1, t_index, t1 - t0, t2 - t1))
        times = times[warmup:]
        t_pref = float(np.mean([t[0] for t in times])); t_aln = float(np.mean([t[1] for t in times]))
        return {"t_prefilter_s": t_pref, "t_align_s": t_aln, "t_index_s": float(np.mean([t[2] for t in times])),
                "wall_prefilter_s": float(np.mean([t[3] for t in times])), "wall_align_s": float(np.mean([t[4] for t in times])),
                "qps": nq / (t_pref + t_aln), "threads": threads, "nq": nq}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def ncu_traffic_per_query():
    """dram__bytes_read.sum + dram__bytes_write.sum of match_kernel per query, from the committed ncu --set full capture
    (profiles/r01_pf_match5.txt: one launch over 148 queries of this workload). Returns (bytes, provenance) or (None, None)."""
    path = os.path.join(ROOT, "profiles", "r01_pf_match5.txt")
    try:
        rd = wr = None
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] == "dram__bytes_read.sum":
                rd = float(f[2]) * scale[f[1]]
            elif len(f) >= 3 and f[0] == "dram__bytes_write.sum":
                wr = float(f[2]) * scale[f[1]]
        if rd is None or wr is None:
            return None, None
        return (rd + wr) / 148.0, "profiles/r01_pf_match5.txt (ncu --set full, 148-query launch) scaled to the queries of one bench launch"
    except (OSError, ValueError, KeyError):
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "configs[1]: %d synthetic queries (len %d, random 3Di+AA, %d%% planted) vs %d-structure DB, prefilter+align+traceback, "
                          "-s 9.5 -k 6 --max-seqs %d -e 10 --alignment-type 2" % (NQ, LEN, int(PLANT * 100), NT, MAXSEQS),
              "queries_per_gpu": NQ, "targets": NT, "parallelism": "query-sharded x%d, DB replicated (NCCL broadcast)" % world,
              "l2": "inputs > L2: k-mer index + 3-mer tables + DB = 1.1 GB per pass vs 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        warm = min(args.warmup, 1)
        r = run_reference(NQ, max(1, min(args.steps, 3)), warm)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/refseek not built"}))
            return
        ms = 1000.0 * (r["t_prefilter_s"] + r["t_align_s"])
        print(json.dumps({
            "impl": "reference", "metric": "query-structures/s", "value": r["qps"], "unit": "queries/s", "n_gpus": args.gpus,
            "steps": max(1, min(args.steps, 3)), "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["qps"], "unit": "queries/s", "cores": r["threads"], "kind": "reference",
                             "sample": "all %d queries; query loops of the reference prefilter (%.2fs) + structurealign (%.2fs) modules; module "
                                       "start-up incl. index-table build excluded like DB staging on the GPU arm (module wall times %.2fs + %.2fs); "
                                       "traceback through the restated block-aligner engine (no Rust toolchain for the crate itself)" % (NQ, r["t_prefilter_s"], r["t_align_s"], r["wall_prefilter_s"], r["wall_align_s"])},
            "e2e": {"value": r["qps"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ranks share the host: split the cores between their OpenMP teams (host-side query preparation, task building).
    # torch.distributed.run exports OMP_NUM_THREADS=1 for every worker, which would serialise the host side of each rank
    # (measured at 2 GPUs: 234 ms/step instead of 183): the bench sets its own split unless FSB200_OMP_THREADS says otherwise.
    os.environ["OMP_NUM_THREADS"] = os.environ.get("FSB200_OMP_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world))))
    import torch
    from foldseek_b200 import api
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    taa, tss, qaa, qss = make_workload(rank)
    ctx = api.Context(local_rank, max_seqs=MAXSEQS)
    nt, nq = NT, NQ
    off = np.arange(nt, dtype=np.uint64) * LEN
    ln = np.full(nt, LEN, np.uint32)
    t_stage0 = time.time()
    if world == 1:
        ctx.load_target_flat(np.ascontiguousarray(taa.reshape(-1)), np.ascontiguousarray(tss.reshape(-1)), off, ln, nt * LEN)
    else:
        # rank 0 stages the padded DB; the staged arrays travel over NVLink (one NCCL broadcast), no further collectives
        meta = torch.zeros(1, dtype=torch.int64, device="cuda")
        if rank == 0:
            ctx.load_target_flat(np.ascontiguousarray(taa.reshape(-1)), np.ascontiguousarray(tss.reshape(-1)), off, ln, nt * LEN)
            pa, pt, nbytes, soff, sln = ctx.staged_target()
            meta[0] = nbytes
        dist.broadcast(meta, 0)
        nbytes = int(meta.item())
        buf = torch.empty(2 * nbytes, dtype=torch.uint8, device="cuda")
        offs_t = torch.empty(nt, dtype=torch.int64, device="cuda")
        if rank == 0:
            import ctypes
            cudart = ctypes.CDLL("libcudart.so")
            cudart.cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(pa), ctypes.c_size_t(nbytes), ctypes.c_int(3))
            cudart.cudaMemcpy(ctypes.c_void_p(buf.data_ptr() + nbytes), ctypes.c_void_p(pt), ctypes.c_size_t(nbytes), ctypes.c_int(3))
            offs_t.copy_(torch.from_numpy(soff.astype(np.int64)))
        dist.broadcast(buf, 0)
        dist.broadcast(offs_t, 0)
        torch.cuda.synchronize()
        if rank != 0:
            ctx.load_target_device(buf.data_ptr(), buf.data_ptr() + nbytes, nbytes, offs_t.cpu().numpy().astype(np.uint64), ln, nt * LEN)
    t_stage = time.time() - t_stage0
    qflat_aa = np.ascontiguousarray(qaa.reshape(-1)); qflat_ss = np.ascontiguousarray(qss.reshape(-1))
    qoff = np.arange(nq, dtype=np.uint64) * LEN
    qln = np.full(nq, LEN, np.uint32)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()
        barrier()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- resident: queries uploaded once, timed region = prefilter + align on resident data ----
    ctx.set_queries_flat(qflat_aa, qflat_ss, qoff, qln)
    def res_step():
        ctx.search_run()
        ctx.backtrace_run()
    for _ in range(args.warmup):
        res_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res = timed(res_step, args.steps)
    pst, ast, bst = ctx.prefilter_stats(), ctx.align_stats(), ctx.backtrace_stats()
    # ---- end to end: host buffers in, host buffers out, every step ----
    ctx.reuse_buffers = True     # host result arrays are allocated once, like a C++ host module would
    bt_bytes = [0]
    def e2e_step():
        ctx.set_queries_flat(qflat_aa, qflat_ss, qoff, qln)
        r = ctx.search()
        meta, bt = ctx.backtrace(int(r[1].sum()))
        bt_bytes[0] = meta.nbytes + bt.nbytes
        return r
    for _ in range(min(args.warmup, 1)):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    sampler.stop_flag = True
    hits, counts, aln = ctx.search_download()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pk, pk_kind = peaks()
    per_step_res = ms_res / args.steps
    per_step_e2e = ms_e2e / args.steps
    value = world * nq / (per_step_res / 1000.0)
    e2e_value = world * nq / (per_step_e2e / 1000.0)
    # roofline of the dominant kernel (prefilter match kernel, HBM-bound gather/scan; SURVEY.md 8d byte model)
    match_s = float(pst[6]) / 1e9 / max(1, int(pst[7]))
    alg_bytes = 16.0 * float(pst[0]) + 6.0 * float(pst[1]) + 14.0 * float(pst[1]) + float(pst[3]) + 7.0 * float(pst[2])
    ach = alg_bytes / match_s / 1e9
    traffic, traffic_src = ncu_traffic_per_query()
    if traffic is not None:
        traffic *= float(nq) / max(1, int(pst[7]))
    sw_s = float(ast[6]) / 1e9
    cells = float(ast[1] + ast[2] + ast[3])
    sm_clock = 1.965e9
    alu_peak = 148 * 64 * sm_clock   # ALU-pipe lane-instructions/s (4 SMSP x 16 lanes), nominal max clock
    out = {
        "metric": "query-structures/s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": config,
        "e2e": {"value": e2e_value, "unit": "queries/s", "ms_per_step": per_step_e2e,
                "h2d_bytes_per_step": int(2 * nq * LEN + 12 * nq), "d2h_bytes_per_step": int(hits.nbytes + counts.nbytes + aln.nbytes + bt_bytes[0])},
        "gpu_launches": int(pst[4] + ast[4] + bst[1]) * args.steps,
        "traceback": {"pairs": int(bst[0]), "kernel_ms": float(bst[2]) / 1e6, "launches": int(bst[1]), "score_mismatch": int(bst[4])},
        "aligned_pairs_per_s": float(ast[0]) * world / (per_step_res / 1000.0),
        "dp_cells_per_s_kernel": cells / sw_s if sw_s > 0 else None,
        "roofline": {"kernel": "prefilter::match_kernel", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src, "peak_kind": pk_kind,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": match_s * 1000.0,
                     "counters": {"similar_kmers": int(pst[0]), "index_hits": int(pst[1]), "double_hits": int(pst[2]), "ungapped_cells": int(pst[3])}},
        "roofline_sw": {"kernel": "sw_packed_kernel<22>", "bound": "alu", "achieved": 4.0 * cells / sw_s / 1e12 if sw_s > 0 else None,
                        "peak": alu_peak / 1e12, "unit": "T lane-instr/s (4 per cell, SURVEY.md 8d)",
                        "frac": (4.0 * cells / sw_s) / alu_peak if sw_s > 0 else None, "cells": cells, "kernel_ms": sw_s * 1000.0},
        "stage_db_s": t_stage, "clocks": sampler.summary(),
        "results": {"hits_per_query": float(counts.mean()), "alignments_passing": int((aln["status"] == 0).sum() - (aln["fwd_score"] == 0).sum() * 0)},
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            ncores = os.cpu_count()
            sample = NQ if ncores >= 32 else 128
            r = run_reference(sample, 1, 0)
            if r:
                out["cpu_baseline"] = {"value": r["qps"], "unit": "queries/s", "cores": r["threads"], "kind": "reference",
                                       "sample": "%d of the %d queries vs the full %d-target DB; query loops of the reference prefilter %.2fs + "
                                                 "structurealign %.2fs (module start-up incl. index build excluded, module wall %.2fs + %.2fs; "
                                                 "traceback through the restated block-aligner engine)" % (sample, NQ, NT, r["t_prefilter_s"], r["t_align_s"],
                                                                         r["wall_prefilter_s"], r["wall_align_s"])}
        except Exception as e:  # the baseline must not take the bench line down
            out["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
