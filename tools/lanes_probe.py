"""Probe: does a second context ("lane") on the SAME GPU hide the host-side parts of a step?
    python tools/lanes_probe.py [queries] [steps] [lanes]
Each lane is a context with its own stream and host thread; the query batch is split over the lanes by residues.
Prints ms per step (host clock around all lanes, device events per lane), resident and end to end."""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from foldseek_b200 import api  # noqa: E402
from foldseek_b200.dist import shard_by_residues  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
nt = int(os.environ.get("PROF_NT", bench.NT))
LEN = bench.LEN
taa, tss, pool, qaa, qss, qca = bench.make_workload(nq=nq, nt=nt)
off = np.arange(nt, dtype=np.uint64) * LEN
ln = np.full(nt, LEN, np.uint32)
px = np.ascontiguousarray(pool[:, 0, :].reshape(-1)); py = np.ascontiguousarray(pool[:, 1, :].reshape(-1)); pz = np.ascontiguousarray(pool[:, 2, :].reshape(-1))
caoff = (np.arange(nt, dtype=np.uint64) % bench.CA_POOL) * LEN


def flat_queries(b, e):
    n = e - b
    return (np.ascontiguousarray(qaa[b:e].reshape(-1)), np.ascontiguousarray(qss[b:e].reshape(-1)), np.arange(n, dtype=np.uint64) * LEN,
            np.full(n, LEN, np.uint32), np.ascontiguousarray(qca[b:e, 0, :].reshape(-1)), np.ascontiguousarray(qca[b:e, 1, :].reshape(-1)),
            np.ascontiguousarray(qca[b:e, 2, :].reshape(-1)))


def make_lanes(n):
    c = api.Context(0, max_seqs=bench.MAXSEQS)
    c.load_target_flat(np.ascontiguousarray(taa.reshape(-1)), np.ascontiguousarray(tss.reshape(-1)), off, ln, nt * LEN)
    c.load_target_ca_flat(px, py, pz, caoff)
    c.build_index()
    ctxs = [c] + [c.create_lane() for _ in range(n - 1)]
    threads = max(1, bench.host_cpus()["usable"] // n)
    for x in ctxs:
        x.reuse_buffers = True
        x.set_host_threads(threads)
    return ctxs


def run(ctxs, e2e):
    n = len(ctxs)
    Q = []
    for k in range(n):
        b, e = shard_by_residues(np.full(nq, LEN), n, k)
        Q.append(flat_queries(b, e))

    def upload(c, Qt):
        c.set_queries_flat(Qt[0], Qt[1], Qt[2], Qt[3])
        c.set_queries_ca_flat(Qt[4], Qt[5], Qt[6], Qt[2])

    def step(k):
        c = ctxs[k]
        if e2e:
            upload(c, Q[k])
            r = c.search()
            npairs = int(r[1].sum())
            c.backtrace(npairs)
            c.structure_scores(npairs)
        else:
            c.search_run(); c.backtrace_run(); c.structure_scores_run()

    for k in range(n):
        upload(ctxs[k], Q[k])
    dev = [0.0] * n

    stagger = [0.0]

    def worker(k, reps, timed):
        if timed:
            ctxs[k].timer_start()
        if stagger[0] > 0 and k > 0:
            time.sleep(k * stagger[0] / n)      # lanes out of phase: one lane's host-side parts fall into the other's kernels
        for _ in range(reps):
            step(k)
        if timed:
            dev[k] = ctxs[k].timer_stop()

    def all_lanes(reps, timed):
        th = [threading.Thread(target=worker, args=(k, reps, timed)) for k in range(n)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return 1e3 * (time.perf_counter() - t0)
    all_lanes(2, False)
    wall = all_lanes(steps, True)
    print("%d lane(s), %s: %.1f ms/step host clock, device events per lane %s ms/step" % (
        n, "e2e" if e2e else "resident", wall / steps, [round(d / steps, 1) for d in dev]), flush=True)
    if n > 1:
        for frac in (0.5, 0.35):
            stagger[0] = frac * 2 * (wall / steps) / 1e3          # seconds; lane k starts k/n of this later
            w2 = all_lanes(steps, True)
            print("%d lane(s), %s, staggered by %.0f ms: %.1f ms/step host clock, device events per lane %s ms/step" % (
                n, "e2e" if e2e else "resident", 1e3 * stagger[0] / n, w2 / steps, [round(d / steps, 1) for d in dev]), flush=True)
        stagger[0] = 0.0


ctxs = make_lanes(lanes)
run(ctxs, False)
run(ctxs, True)
