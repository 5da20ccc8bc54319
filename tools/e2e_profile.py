"""Host-side profile of the end-to-end API step (bench.py's e2e leg): cProfile over Recommender.bulk_fit(DataFrame) +
recommend_batch(all users) at a workload's full size.  Prints the top functions by cumulative time.  GPU box only.

    python tools/e2e_profile.py [--workload ml20m] [--reps 3]
"""
import argparse
import contextlib
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ml20m")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    import numpy as np
    import pandas as pd
    import torch
    import bench
    from rtrec_b200.models import SLIM
    from rtrec_b200.recommender import Recommender
    wl = bench.WORKLOADS[a.workload]
    u, i, ts, r = bench.load_events(wl["shape"])
    U = int(u.max()) + 1
    df = pd.DataFrame({"user": u, "item": i, "tstamp": ts, "rating": r})
    users = list(range(U))
    pr = cProfile.Profile()
    for rep in range(a.reps + 1):
        rec = Recommender(SLIM(**wl["kwargs"]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rep:
            pr.enable()
        with contextlib.redirect_stdout(io.StringIO()):
            rec.bulk_fit(df, parallel=True)
        t1 = time.perf_counter()
        out = rec.recommend_batch(users, top_k=10, filter_interacted=True)
        torch.cuda.synchronize()
        if rep:
            pr.disable()
        t2 = time.perf_counter()
        print(f"rep {rep}: fit {1e3 * (t1 - t0):.1f} ms, recommend {1e3 * (t2 - t1):.1f} ms", flush=True)
        del rec, out
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(a.top)
    print(s.getvalue())


if __name__ == "__main__":
    main()
