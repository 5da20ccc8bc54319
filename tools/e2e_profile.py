#!/usr/bin/env python
"""cProfile of the public-API path (Recommender.bulk_fit + recommend_batch) with host buffers. Development tool."""
import cProfile, io, os, pstats, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import bench
from rtrec_b200.models import SLIM
from rtrec_b200.recommender import Recommender

wl = sys.argv[1] if len(sys.argv) > 1 else "ml20m"
shape, kwargs, desc = (bench.WORKLOADS[wl][k] for k in ("shape", "kwargs", "desc"))
u, i, ts, r = bench.load_events(shape)
U = int(u.max()) + 1
df = pd.DataFrame({"user": u, "item": i, "tstamp": ts, "rating": r})
users = list(range(U))
for rep in range(3):
    rec = Recommender(SLIM(**kwargs))
    pr = cProfile.Profile() if rep == 2 else None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if pr: pr.enable()
    with contextlib.redirect_stdout(io.StringIO()):
        rec.bulk_fit(df, parallel=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = rec.recommend_batch(users, top_k=10, filter_interacted=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if pr: pr.disable()
    print(f"rep {rep}: fit {t1 - t0:.4f}s recommend {t2 - t1:.4f}s")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
