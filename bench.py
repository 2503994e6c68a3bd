#!/usr/bin/env python3
"""bench.py — Foldseek search hot path (3Di k-mer prefilter + structurealign) on B200.

One "step" = one pass of the hot path over the query batch against the staged 1M-structure target DB:
    prefilter -> 3Di+AA Smith-Waterman (fwd + reversed query, E-value gates) -> block-aligner traceback of the accepted pairs
    -> TM-score / LDDT structure bits                                   (libfsb200.so, include/fsb200.h)
metric  : query-structures/s (BASELINE.json), whole job, all ranks
workload: BASELINE.json configs[2] — 1k synthetic queries (len 350, 10 % planted) vs a 1M-structure DB, default `search` flags
          (-s 9.5 -k 6 --max-seqs 1000 -e 10 --alignment-type 2 --sort-by-structure-bits 1)
value   : queries already resident (set_queries done before the timed region), results stay in the context
e2e     : the same step through the C ABI with HOST buffers: query upload, search, download of hits + alignments + backtraces
          + structure scores
Multi-GPU (torchrun): rank 0 stages the padded target DB and broadcasts the staged arrays over NCCL; the SAME 1k queries are
split over the ranks by residues (strong scaling, no further collectives in the timed region).  After the timed region the
per-rank results are gathered and compared with rank 0's own 1-GPU run over all queries.
Lanes: a GPU's shard runs on one or two contexts that share the staged DB and index (fsb200_create_lane), one host thread each, started
out of phase so that one lane's kernels run during the other's host-side work; `lanes` in the JSON line says how many
each pass used, `single_lane` is the same shard on one context (the per-kernel roofline times come from that pass).
--impl reference: the reference's own CPU modules (oracle/_ref/refseek = unmodified reference sources) on the host threads this
process may use (affinity / cgroup quota recorded), same workload, bounded query sample.
"""
import argparse
import json
import os
import re
import shutil
import struct
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NQ, NT, LEN, MAXSEQS = 1000, 1000000, 350, 1000
PLANT = 0.1
CA_POOL = 1024          # distinct synthetic C-alpha traces; target i uses trace i % CA_POOL


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def make_workload(nq=NQ, nt=NT):
    """Seeded: every rank (and the reference arm) builds the same target DB and the same queries."""
    from foldseek_b200 import synth
    rng = np.random.default_rng(20240922)
    taa = synth.random_codes(rng, nt, LEN, synth.BG_AA)
    tss = synth.random_codes(rng, nt, LEN, synth.BG_3DI)
    pool = synth.random_walk_ca(np.random.default_rng(99), CA_POOL, LEN)          # [CA_POOL, 3, LEN]
    qrng = np.random.default_rng(777)
    qaa = synth.random_codes(qrng, NQ, LEN, synth.BG_AA)
    qss = synth.random_codes(qrng, NQ, LEN, synth.BG_3DI)
    qca = synth.random_walk_ca(qrng, NQ, LEN)
    # planted homologs: mutated copies of targets (substitutions only keep the length at 350), C-alpha = target trace + noise
    n_plant = int(NQ * PLANT)
    src = qrng.integers(0, nt, size=n_plant)
    for k in range(n_plant):
        a, s = taa[src[k]].copy(), tss[src[k]].copy()
        m = qrng.random(LEN) < 0.3
        a[m] = qaa[k][m]
        m = qrng.random(LEN) < 0.3
        s[m] = qss[k][m]
        qaa[k], qss[k] = a, s
        qca[k] = pool[src[k] % CA_POOL] + qrng.normal(scale=0.5, size=(3, LEN)).astype(np.float32)
    return taa, tss, pool, qaa[:nq], qss[:nq], qca[:nq]


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


def host_cpus():
    """What this process may actually use: affinity mask, cgroup CPU quota, CPU model, load."""
    info = {"os_cpu_count": os.cpu_count()}
    try:
        info["affinity"] = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        info["affinity"] = os.cpu_count()
    quota = None
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(p).read().split()
            if p.endswith("cpu.max"):
                if txt[0] != "max":
                    quota = float(txt[0]) / float(txt[1])
            else:
                q = float(txt[0])
                per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    quota = q / per
            break
        except (OSError, ValueError, IndexError):
            continue
    info["cgroup_cpu_quota"] = quota
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                info["model"] = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    try:
        info["loadavg_1m"] = os.getloadavg()[0]
    except OSError:
        pass
    usable = info["affinity"] or 1
    if quota:
        usable = max(1, min(usable, int(quota + 0.5)))
    info["usable"] = usable
    return info


def _last_progress_seconds(text):
    """Elapsed time printed by the reference's progress bar for its last (query) loop, e.g. '] 1.00K 0s 337ms'."""
    best = None
    for m in re.finditer(r"\]\s+\S+\s+(?:(\d+)h\s+)?(?:(\d+)m\s+)?(\d+)s\s+(\d+)ms", text):
        best = int(m.group(1) or 0) * 3600 + int(m.group(2) or 0) * 60 + int(m.group(3)) + int(m.group(4)) / 1000.0
    return best


def write_ca_db_pooled(path, pool, n):
    """<db>_ca with n entries that alias CA_POOL raw-float32 payloads (Coordinate16::read returns raw floats as-is,
    src/commons/Coordinate16.h:15-18); entry i -> payload i % CA_POOL."""
    L = pool.shape[2]
    payload = 3 * L * 4 + 1
    buf = np.zeros((len(pool), payload), np.uint8)
    buf[:, :3 * L * 4] = np.ascontiguousarray(pool, np.float32).reshape(len(pool), -1).view(np.uint8)
    buf.tofile(path)
    with open(path + ".index", "w") as f:
        f.write("".join("%d\t%d\t%d\n" % (i, (i % len(pool)) * payload, payload) for i in range(n)))
    with open(path + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", 12))


def write_ca_db(path, ca):
    n, _, L = ca.shape
    payload = 3 * L * 4 + 1
    buf = np.zeros((n, payload), np.uint8)
    buf[:, :3 * L * 4] = np.ascontiguousarray(ca, np.float32).reshape(n, -1).view(np.uint8)
    buf.tofile(path)
    with open(path + ".index", "w") as f:
        f.write("".join("%d\t%d\t%d\n" % (i, i * payload, payload) for i in range(n)))
    with open(path + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", 12))


class ReferenceRun:
    """The reference's prefilter + structurealign modules on the workload; the DB files are written once."""

    def __init__(self, sample, tmp=None):
        from foldseek_b200 import synth
        self.refseek = os.path.join(ROOT, "oracle", "_ref", "refseek")
        self.ok = os.path.exists(self.refseek)
        self.sample = sample
        if not self.ok:
            return
        taa, tss, pool, qaa, qss, qca = make_workload(nq=sample)
        self.d = tempfile.mkdtemp(prefix="fsb200_ref_", dir=tmp)
        synth.write_seq_db_fixed(os.path.join(self.d, "t"), taa)
        synth.write_seq_db_fixed(os.path.join(self.d, "t_ss"), tss)
        write_ca_db_pooled(os.path.join(self.d, "t_ca"), pool, NT)
        synth.write_seq_db_fixed(os.path.join(self.d, "q"), qaa)
        synth.write_seq_db_fixed(os.path.join(self.d, "q_ss"), qss)
        write_ca_db(os.path.join(self.d, "q_ca"), qca)

    def close(self):
        if self.ok:
            shutil.rmtree(self.d, ignore_errors=True)

    def run(self, threads):
        d = self.d
        for f in os.listdir(d):
            if f.startswith("pref") or f.startswith("aln"):
                os.remove(os.path.join(d, f))
        t0 = time.time()
        out1 = subprocess.run([self.refseek, "prefilter", "q_ss", "t_ss", "pref", "-s", "9.5", "--max-seqs", str(MAXSEQS),
                               "--comp-bias-corr-scale", "0.15", "--threads", str(threads)], cwd=d, capture_output=True, text=True)
        t1 = time.time()
        out2 = subprocess.run([self.refseek, "structurealign", "q", "t", "pref", "aln", "-e", "10", "--sort-by-structure-bits", "1",
                               "--threads", str(threads)], cwd=d, capture_output=True, text=True)
        t2 = time.time()
        if out1.returncode != 0 or out2.returncode != 0:
            raise RuntimeError("refseek failed: " + out1.stdout[-500:] + out1.stderr[-300:] + out2.stdout[-500:] + out2.stderr[-300:])
        m = re.search(r"Time for index table init: (\d+)h (\d+)m (\d+)s (\d+)ms", out1.stdout)
        t_index = (int(m.group(1)) * 3600 + int(m.group(2)) * 60 + int(m.group(3)) + int(m.group(4)) / 1000.0) if m else 0.0
        # the query loops themselves (Debug::Progress prints their elapsed time when they finish): this is what a run against a
        # precomputed index (createindex stores the index table and the 3-mer matrices) spends per batch
        lp1, lp2 = _last_progress_seconds(out1.stdout), _last_progress_seconds(out2.stdout)
        t_pref = lp1 if lp1 is not None else (t1 - t0) - t_index
        t_aln = lp2 if lp2 is not None else t2 - t1
        return {"t_prefilter_s": t_pref, "t_align_s": t_aln, "t_index_s": t_index, "wall_prefilter_s": t1 - t0, "wall_align_s": t2 - t1,
                "qps": self.sample / (t_pref + t_aln), "threads": threads, "nq": self.sample}


def reference_sample_size(usable):
    """bounded sample: ~10-30 s of query-loop work on the host (the reference needs ~0.3 core-seconds per query at 1M targets)"""
    return int(min(NQ, max(32, 32 * usable)))


def ncu_traffic(kernel):
    """dram bytes (read + write) per launch of `kernel` from a committed ncu --set full summary (profiles/r02_<kernel>.txt)."""
    path = os.path.join(ROOT, "profiles", "r02_%s.txt" % kernel)
    try:
        rd = wr = None
        launches = 1.0
        scale = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] == "dram__bytes_read.sum":
                rd = float(f[2]) * scale[f[1]]
            elif len(f) >= 3 and f[0] == "dram__bytes_write.sum":
                wr = float(f[2]) * scale[f[1]]
            elif len(f) >= 2 and f[0] == "queries_in_capture":
                launches = float(f[1])
        if rd is None or wr is None:
            return None, None
        return (rd + wr) / launches, "profiles/r02_%s.txt (ncu --set full), per query of the capture" % kernel
    except (OSError, ValueError, KeyError):
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the multi-rank result check against rank 0's 1-GPU run")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("FSB200_LANES", "0")),
                    help="contexts per GPU that share the staged DB + index, one host thread each (hides the host-side parts of a step); "
                         "0 = 2 lanes when the GPU's shard has at least 800 queries, else 1 (measured: 1000 queries/GPU +3.3 %%, 500 queries/GPU -1.7 %%: "
                         "half-size batches cost more in kernel tails and launches than the hidden host time is worth)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "configs[2]: %d synthetic queries (len %d, random 3Di+AA+C-alpha, %d%% planted) vs %d-structure DB, prefilter + "
                          "structurealign (SW fwd+rev, traceback, TM/LDDT), -s 9.5 -k 6 --max-seqs %d -e 10 --alignment-type 2 "
                          "--sort-by-structure-bits 1" % (NQ, LEN, int(PLANT * 100), NT, MAXSEQS),
              "queries": NQ, "targets": NT, "parallelism": "the %d queries split over %d GPU(s) by residues, DB replicated (one NCCL broadcast)" % (NQ, world),
              "l2": "inputs > L2: k-mer index 3.5 GB + 3-mer tables 0.4 GB + DB 0.7 GB + per-step hit stream ~100 GB vs 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        cpus = host_cpus()
        threads = cpus["usable"]
        sample = reference_sample_size(threads)
        steps = max(1, min(args.steps, 2))
        warm = 0
        ref = ReferenceRun(sample)
        if not ref.ok:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/refseek not built"}))
            return
        try:
            runs = [ref.run(threads) for _ in range(warm + steps)][warm:]
        finally:
            ref.close()
        t_pref = float(np.mean([r["t_prefilter_s"] for r in runs])); t_aln = float(np.mean([r["t_align_s"] for r in runs]))
        qps = sample / (t_pref + t_aln)
        print(json.dumps({
            "impl": "reference", "metric": "query-structures/s", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * (t_pref + t_aln) * NQ / sample, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "reference", "host": cpus,
                             "sample": "%d of the %d queries vs the full %d-target DB (per-query work is independent: queries/s = sample / time; ms_per_step "
                                       "is scaled to %d queries); query loops of the reference prefilter (%.2fs) + structurealign (%.2fs) modules; module "
                                       "start-up incl. index-table build (%.1fs) excluded like DB staging + index build on the GPU arm (module wall times "
                                       "%.1fs + %.1fs); traceback through the restated block-aligner engine (no Rust toolchain for the crate itself)"
                                       % (sample, NQ, NT, NQ, t_pref, t_aln, runs[-1]["t_index_s"], runs[-1]["wall_prefilter_s"], runs[-1]["wall_align_s"])},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    cpus = host_cpus()
    # ranks share the host: split the usable cores between their OpenMP teams (host-side query preparation, task building).
    # torch.distributed.run exports OMP_NUM_THREADS=1 for every worker, which would serialise the host side of each rank
    os.environ["OMP_NUM_THREADS"] = os.environ.get("FSB200_OMP_THREADS", str(max(1, cpus["usable"] // max(1, world))))
    import torch
    from foldseek_b200 import api
    from foldseek_b200.dist import shard_by_residues
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    taa, tss, pool, qaa, qss, qca = make_workload()
    ctx = api.Context(local_rank, max_seqs=MAXSEQS)
    nt = NT
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
        del buf
        torch.cuda.empty_cache()
    # C-alpha: target i aliases pool trace i % CA_POOL (the library stages one copy per target)
    px = np.ascontiguousarray(pool[:, 0, :].reshape(-1)); py = np.ascontiguousarray(pool[:, 1, :].reshape(-1)); pz = np.ascontiguousarray(pool[:, 2, :].reshape(-1))
    ctx.load_target_ca_flat(px, py, pz, (np.arange(nt, dtype=np.uint64) % CA_POOL) * LEN)
    t_stage = time.time() - t_stage0
    t_index0 = time.time()
    ctx.build_index()
    t_index = time.time() - t_index0

    # strong scaling: the same NQ queries, split by residues (DBReader::decomposeDomainByAminoAcid, Prefiltering.cpp:781-786)
    qb, qe = shard_by_residues(np.full(NQ, LEN), world, rank)
    nq = qe - qb
    # lanes: contexts on this GPU that share the staged DB, C-alpha and index; the rank's queries are split over them the same way.
    # Half-size batches cost kernel tails and launches, hidden host time pays for them only when there is enough of it (measured, 5
    # steps): resident 1000 queries/GPU +3.3 %, 500 queries/GPU -3.7 %; end to end (more host work per step) 1000: +5 %, 500: +5 %.
    if args.lanes > 0:
        lanes_res = lanes_e2e = max(1, min(args.lanes, nq))
    else:
        lanes_res = 2 if nq >= 800 else 1
        lanes_e2e = 2 if nq >= 400 else 1
    lanes = [ctx] + [ctx.create_lane() for _ in range(max(lanes_res, lanes_e2e) - 1)]
    cur = {"n": 1, "LQ": None, "threads": 1}

    def flat_queries(b, e):
        n = e - b
        return (np.ascontiguousarray(qaa[b:e].reshape(-1)), np.ascontiguousarray(qss[b:e].reshape(-1)), np.arange(n, dtype=np.uint64) * LEN,
                np.full(n, LEN, np.uint32), np.ascontiguousarray(qca[b:e, 0, :].reshape(-1)), np.ascontiguousarray(qca[b:e, 1, :].reshape(-1)),
                np.ascontiguousarray(qca[b:e, 2, :].reshape(-1)))

    def set_lanes(n):
        """the next passes run on the first n lanes: query ranges by residues, the rank's host threads split over them"""
        cur["n"] = n
        cur["LQ"] = [flat_queries(qb + b, qb + e) for b, e in (shard_by_residues(np.full(nq, LEN), n, k) for k in range(n))]
        cur["threads"] = max(1, cpus["usable"] // max(1, world) // n)
        for c in lanes[:n]:
            c.set_host_threads(cur["threads"])

    def upload(k):
        LQ = cur["LQ"]
        lanes[k].set_queries_flat(LQ[k][0], LQ[k][1], LQ[k][2], LQ[k][3])
        lanes[k].set_queries_ca_flat(LQ[k][4], LQ[k][5], LQ[k][6], LQ[k][2])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_lanes(fn, reps, timers=None, stagger_s=0.0):
        """every lane runs fn(k) reps times on its own host thread (ctypes releases the GIL inside the library).  Lane k starts
        k * stagger_s later (inside its timed region): lanes that start together stay in lockstep - both in their host-side parts at the
        same time - while lanes out of phase keep the GPU busy with one lane's kernels during the other's host work."""
        errs = []

        def work(k):
            try:
                if timers is not None:
                    lanes[k].timer_start()
                if stagger_s > 0 and k > 0:
                    time.sleep(k * stagger_s)
                for _ in range(reps):
                    fn(k)
                if timers is not None:
                    timers[k] = lanes[k].timer_stop()
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        if cur["n"] == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(k,)) for k in range(cur["n"])]
            for t in th:
                t.start()
            for t in th:
                t.join()
        if errs:
            raise errs[0]

    def timed(fn, steps, stagger_s=0.0):
        """device time of `steps` steps of every lane: CUDA events on each lane's stream (the start events are recorded together, after a
        synchronise and before any stagger delay), the slowest lane counts; max over ranks"""
        barrier()
        tms = [0.0] * cur["n"]
        t0 = time.perf_counter()
        run_lanes(fn, steps, tms, stagger_s)
        wall = 1e3 * (time.perf_counter() - t0)
        ms = max(tms)
        barrier()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    # ---- resident: queries uploaded once, timed region = prefilter + align + traceback + structure scores on resident data ----
    set_lanes(lanes_res)
    run_lanes(lambda k: upload(k), 1)

    def res_step(k):
        lanes[k].search_run()
        lanes[k].backtrace_run()
        lanes[k].structure_scores_run()
    t_w = time.perf_counter()
    run_lanes(res_step, args.warmup)
    # stagger between lanes: 0.35 of a lane's step (measured on the warm-up steps), split over the lanes
    stagger_res = 0.7 * (time.perf_counter() - t_w) / max(1, args.warmup) / lanes_res if lanes_res > 1 else 0.0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, wall_res = timed(res_step, args.steps, stagger_res)
    launches = sum(int(c.prefilter_stats()[4] + c.align_stats()[4] + c.backtrace_stats()[1] + c.structure_scores_stats()[1]) for c in lanes[:lanes_res])
    threads_res = cur["threads"]
    # ---- end to end: host buffers in, host buffers out, every step ----
    for c in lanes:
        c.reuse_buffers = True     # host result arrays are allocated once, like a C++ host module would
    set_lanes(lanes_e2e)
    d2h = [0] * lanes_e2e
    last = [None] * lanes_e2e

    def e2e_step(k):
        upload(k)
        r = lanes[k].search()
        npairs = int(r[1].sum())
        meta, bt = lanes[k].backtrace(npairs)
        ss = lanes[k].structure_scores(npairs)
        d2h[k] = r[0].nbytes + r[1].nbytes + r[2].nbytes + meta.nbytes + bt.nbytes + ss.nbytes
        last[k] = (r, meta, ss)
    t_w = time.perf_counter()
    run_lanes(e2e_step, max(1, args.warmup))
    stagger_e2e = 0.7 * (time.perf_counter() - t_w) / max(1, args.warmup) / lanes_e2e if lanes_e2e > 1 else 0.0
    ms_e2e, wall_e2e = timed(e2e_step, args.steps, stagger_e2e)
    sampler.stop_flag = True
    run_lanes(e2e_step, 1)
    hits = np.concatenate([last[k][0][0] for k in range(lanes_e2e)])
    counts = np.concatenate([last[k][0][1] for k in range(lanes_e2e)])
    aln = np.concatenate([last[k][0][2] for k in range(lanes_e2e)])
    threads_e2e = cur["threads"]
    mask = np.arange(MAXSEQS)[None, :] < counts[:, None]
    passing = int((aln["status"][mask] == 0).sum())

    # ---- one lane alone over the rank's queries: clean per-kernel device times for the roofline (with several lanes the events of one
    # lane also see the other lanes' kernels), and the host-clock anatomy of an end-to-end step ----
    Q = flat_queries(qb, qe)

    def upload_all():
        ctx.set_queries_flat(Q[0], Q[1], Q[2], Q[3])
        ctx.set_queries_ca_flat(Q[4], Q[5], Q[6], Q[2])

    def res_step_all():
        ctx.search_run()
        ctx.backtrace_run()
        ctx.structure_scores_run()
    ctx.set_host_threads(max(1, cpus["usable"] // max(1, world)))   # alone on the GPU: the whole OpenMP team of this rank
    upload_all()
    res_step_all()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        res_step_all()
    ms_single = ctx.timer_stop()
    barrier()
    pst, ast, bst, sst, stage = ctx.prefilter_stats(), ctx.align_stats(), ctx.backtrace_stats(), ctx.structure_scores_stats(), ctx.prefilter_stage_ms()
    e2e_parts = {}

    def clocked(name, fn, *a):
        t0 = time.perf_counter()
        out = fn(*a)
        e2e_parts[name] = round(1e3 * (time.perf_counter() - t0), 2)
        return out
    clocked("set_queries", ctx.set_queries_flat, Q[0], Q[1], Q[2], Q[3])
    clocked("set_queries_ca", ctx.set_queries_ca_flat, Q[4], Q[5], Q[6], Q[2])
    clocked("search_run", ctx.search_run)
    clocked("search_download", ctx.search_download)
    clocked("backtrace_run", ctx.backtrace_run)
    clocked("backtrace_download", ctx.backtrace_download, int(counts.sum()))
    clocked("structure_scores_run+download", ctx.structure_scores, int(counts.sum()))

    # ---- multi-rank check: the sharded results, gathered, equal rank 0's own run over all queries ----
    verify = None
    if dist is not None and not args.no_verify:
        per = (NQ + world - 1) // world + 1
        pack = torch.zeros((per, MAXSEQS, 4), dtype=torch.int32, device="cuda")
        cnt = torch.zeros(per + 1, dtype=torch.int32, device="cuda")
        h = np.stack([hits["target"].astype(np.int64), hits["score"], hits["diagonal"], aln["score"]], axis=2).astype(np.int32)
        pack[:nq] = torch.from_numpy(h).cuda()
        cnt[:nq] = torch.from_numpy(counts.astype(np.int32)).cuda()
        cnt[per] = nq
        allp = [torch.zeros_like(pack) for _ in range(world)]
        allc = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(allp, pack)
        dist.all_gather(allc, cnt)
        if rank == 0:
            FQ = flat_queries(0, NQ)
            ctx.set_queries_flat(FQ[0], FQ[1], FQ[2], FQ[3])
            ctx.set_queries_ca_flat(FQ[4], FQ[5], FQ[6], FQ[2])
            fh, fc, fa = ctx.search()
            full = np.stack([fh["target"].astype(np.int64), fh["score"], fh["diagonal"], fa["score"]], axis=2).astype(np.int32)
            got, gotc = [], []
            for r in range(world):
                k = int(allc[r][per].item())
                got.append(allp[r][:k].cpu().numpy()); gotc.append(allc[r][:k].cpu().numpy())
            got = np.concatenate(got); gotc = np.concatenate(gotc)
            fm = np.arange(MAXSEQS)[None, :] < fc[:, None]
            same = len(got) == NQ and (gotc == fc.astype(np.int32)).all() and (got[fm] == full[fm]).all()
            verify = {"queries": int(len(got)), "identical_to_1gpu": bool(same)}
            if not same:
                raise SystemExit("multi-rank results differ from the 1-GPU run")
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pk, pk_kind = peaks()
    per_step_res = ms_res / args.steps
    per_step_e2e = ms_e2e / args.steps
    value = NQ / (per_step_res / 1000.0)
    e2e_value = NQ / (per_step_e2e / 1000.0)
    # ---- roofline per prefilter stage (SURVEY.md 8d byte model, split by the kernel that moves the bytes); rank 0's shard ----
    kmers, hitsN, dbl, cells = float(pst[0]), float(pst[1]), float(pst[2]), float(pst[3])
    stage_bytes = {"kmer_lists": 16.0 * kmers,                  # index offsets/list heads per similar k-mer
                   "expand": 6.0 * hitsN + 7.0 * hitsN,         # index entries read + hit stream written
                   "diag": 7.0 * hitsN + 7.0 * dbl,             # hit stream read + double hits written
                   "score": cells + 7.0 * dbl}                  # target residues on the diagonals + candidates
    kern = {"kmer_lists": "prefilter::kmer_lists_kernel", "expand": "prefilter::expand_kernel", "diag": "prefilter::diag_kernel", "score": "prefilter::score_kernel"}
    stages = []
    for k in ("kmer_lists", "expand", "diag", "score"):
        ms = stage[k]
        ach = stage_bytes[k] / (ms / 1000.0) / 1e9 if ms > 0 else 0.0
        stages.append({"kernel": kern[k], "ms_per_step": ms, "algorithmic_bytes": stage_bytes[k], "achieved": ach, "frac": ach / pk["hbm_gbs"]})
    dom = max(stages, key=lambda s: s["ms_per_step"])
    traffic, traffic_src = ncu_traffic(dom["kernel"].split("::")[1])
    if traffic is not None:
        traffic *= nq
    match_ms = stage["kmer_lists"] + stage["expand"] + stage["diag"]
    match_bytes = stage_bytes["kmer_lists"] + stage_bytes["expand"] + stage_bytes["diag"]
    sw_s = float(ast[6]) / 1e9
    cells_sw = float(ast[1] + ast[2] + ast[3])
    sm_clock = 1.965e9
    alu_peak = 148 * 64 * sm_clock   # ALU-pipe lane-instructions/s (4 SMSP x 16 lanes), nominal max clock
    out = {
        "metric": "query-structures/s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step_res, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": config,
        "e2e_host_clock_ms": e2e_parts,
        "e2e": {"value": e2e_value, "unit": "queries/s", "ms_per_step": per_step_e2e,
                "h2d_bytes_per_step": int(2 * nq * LEN + 12 * nq + 3 * 4 * nq * LEN), "d2h_bytes_per_step": int(sum(d2h)), "note": "bytes of rank 0's shard",
                "host_clock_ms_per_step": wall_e2e / args.steps},
        "gpu_launches": launches * args.steps,
        "lanes": {"per_gpu": lanes_res, "per_gpu_e2e": lanes_e2e, "host_threads_per_lane": threads_res, "host_threads_per_lane_e2e": threads_e2e,
                  "stagger_ms": 1e3 * stagger_res, "stagger_ms_e2e": 1e3 * stagger_e2e,
                  "note": "contexts on one GPU sharing the staged DB, C-alpha and k-mer index (fsb200_create_lane), one host thread and stream each; "
                          "value / e2e = all lanes together (CUDA events per lane stream, slowest lane); the roofline stage times come from one "
                          "lane alone over the same queries (single_lane), where the events of a stage see only that stage's kernels",
                  "host_clock_ms_per_step_resident": wall_res / args.steps},
        "single_lane": {"ms_per_step": ms_single / args.steps, "value": nq * world / (ms_single / args.steps / 1000.0) if ms_single > 0 else None,
                        "note": "rank 0's shard on one lane, resident; scaled by the number of ranks"},
        "traceback": {"pairs": int(bst[0]), "kernel_ms": float(bst[2]) / 1e6, "launches": int(bst[1]), "score_mismatch": int(bst[4])},
        "structure_scores": {"pairs": int(sst[0]), "kernel_ms": float(sst[2]) / 1e6, "degenerate": int(sst[3])},
        "aligned_pairs_per_s": float(ast[0]) * (NQ / max(1, nq)) / (per_step_res / 1000.0),
        "dp_cells_per_s_kernel": cells_sw / sw_s if sw_s > 0 else None,
        "roofline": {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_kind": pk_kind,
                     "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "launch_ms": dom["ms_per_step"],
                     "note": "rank 0's shard (%d queries); launch_ms = sum of this kernel's launches in one step (CUDA events on the launching stream)" % nq,
                     "counters": {"similar_kmers": int(pst[0]), "index_hits": int(pst[1]), "double_hits": int(pst[2]), "ungapped_cells": int(pst[3]),
                                  "overflow_flushes": int(pst[5])}},
        "roofline_stages": stages,
        "roofline_match_pipeline": {"kernels": "kmer_lists + expand + diag", "ms_per_step": match_ms, "algorithmic_bytes": match_bytes,
                                    "achieved": match_bytes / (match_ms / 1000.0) / 1e9 if match_ms > 0 else None,
                                    "frac": (match_bytes / (match_ms / 1000.0) / 1e9) / pk["hbm_gbs"] if match_ms > 0 else None},
        "prefilter_stage_ms": stage,
        "roofline_sw": {"kernel": "sw_packed_kernel<22>", "bound": "alu", "achieved": 4.0 * cells_sw / sw_s / 1e12 if sw_s > 0 else None,
                        "peak": alu_peak / 1e12, "unit": "T lane-instr/s (4 per cell, SURVEY.md 8d)",
                        "frac": (4.0 * cells_sw / sw_s) / alu_peak if sw_s > 0 else None, "cells": cells_sw, "kernel_ms": sw_s * 1000.0,
                        "generic_kernel_pairs": int(ast[5])},
        "stage_db_s": t_stage, "index_build_s": t_index, "clocks": sampler.summary(), "host": cpus,
        "results": {"hits_per_query": float(counts.mean()), "alignments_passing": passing, "multi_rank_check": verify},
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            threads = cpus["usable"]
            sample = reference_sample_size(threads)
            ref = ReferenceRun(sample)
            if ref.ok:
                try:
                    r = ref.run(threads)
                    sweep = None
                    if threads >= 4 and r["t_prefilter_s"] + r["t_align_s"] < 40:
                        r2 = ref.run(threads // 2)     # 2-point thread sweep: does the reference scale with the cores it is given?
                        sweep = {"threads": threads // 2, "qps": r2["qps"]}
                finally:
                    ref.close()
                out["cpu_baseline"] = {"value": r["qps"], "unit": "queries/s", "cores": r["threads"], "kind": "reference", "host": cpus, "half_threads": sweep,
                                       "sample": "%d of the %d queries vs the full %d-target DB; query loops of the reference prefilter %.2fs + "
                                                 "structurealign %.2fs (module start-up incl. index build %.1fs excluded, module wall %.1fs + %.1fs; "
                                                 "traceback through the restated block-aligner engine)" % (sample, NQ, NT, r["t_prefilter_s"], r["t_align_s"],
                                                                                                       r["t_index_s"], r["wall_prefilter_s"], r["wall_align_s"])}
        except Exception as e:  # the baseline must not take the bench line down
            out["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
