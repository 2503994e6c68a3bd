"""Profiling driver for the headline workload (bench.py make_workload): prefilter of the first N queries against the 1M DB.
    python tools/prof_1m.py [N] [reps]        # prints the per-stage device times (CUDA events)
Wrapped by ncu for the committed captures under profiles/ (see profiles/README.md)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from foldseek_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nt = int(os.environ.get("PROF_NT", bench.NT))
taa, tss, pool, qaa, qss, qca = bench.make_workload(nq=n, nt=nt)
ctx = api.Context(0)
off = np.arange(nt, dtype=np.uint64) * bench.LEN
ln = np.full(nt, bench.LEN, np.uint32)
ctx.load_target_flat(np.ascontiguousarray(taa.reshape(-1)), np.ascontiguousarray(tss.reshape(-1)), off, ln, nt * bench.LEN)
ctx.set_queries([qaa[i] for i in range(n)], [qss[i] for i in range(n)])
for r in range(reps):
    t = time.time()
    ctx.prefilter()
    dt = time.time() - t
    st = ctx.prefilter_stats()
    print("rep %d: %.1f ms wall; stages ms: %s; kmers %d hits %d double %d cells %d" % (
        r, dt * 1e3, {k: round(v, 2) for k, v in ctx.prefilter_stage_ms().items()}, st[0], st[1], st[2], st[3]), flush=True)
