import sys, numpy as np
sys.path.insert(0, '.')
from foldseek_b200 import api, synth
db = synth.gen_db(3000, 350, 51, min_len=40)
q = synth.gen_queries(db, 4, 350, 52, planted_frac=0.3)
ctx = api.Context(0)
ctx.load_target(db.aa, db.tdi); ctx.set_queries(q.aa, q.tdi)
h = ctx.prefilter(); print([len(x) for x in h])
