#!/usr/bin/env python
"""Calibration of the CPU port against the REAL reference (build container only: /root/reference is not on the GPU box).

Runs BASELINE configs[0] (synthetic ML-1M shape, nn_feature_selection=50) through
  (a) the unmodified reference: Recommender.bulk_fit(df, parallel=True) + recommend_batch in batches of 100 users
      (as evaluate does, recommender.py:185-192), and
  (b) the oracle port exactly as bench.py's cpu_baseline / --impl reference leg times it (full, not sampled),
on the same host cores, and writes the phase times and the ratios to profiles/r3_ref_calibration.json.  bench.py copies
the ``summary`` of that file into ``cpu_baseline.calibration_vs_real_reference``.
"""
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import pandas as pd

import bench
from make_golden import _import_reference


def main():
    _import_reference()
    from rtrec.models import SLIM
    from rtrec.recommender import Recommender
    wl = bench.WORKLOADS["ml1m"]
    u, i, ts, r = bench.load_events(wl["shape"])
    U, I = int(u.max()) + 1, int(i.max()) + 1
    df = pd.DataFrame({"user": u, "item": i, "tstamp": ts, "rating": r})
    out = {"workload": wl["desc"], "cores": os.cpu_count(), "reference_workers": int(0.7 * (os.cpu_count() or 1))}
    # ---- (a) the real reference
    rec = Recommender(SLIM(**wl["kwargs"]))
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        rec.bulk_fit(df, parallel=True)
    t_fit = time.perf_counter() - t0
    users = list(range(U))
    t0 = time.perf_counter()
    lists = []
    for a in range(0, U, 100):
        lists.extend(rec.recommend_batch(users[a:a + 100], top_k=10, filter_interacted=True))
    t_rec = time.perf_counter() - t0
    out["reference"] = {"bulk_fit_sec": round(t_fit, 3), "recommend_sec": round(t_rec, 3), "users_per_s": round(U / (t_fit + t_rec), 2),
                        "recommend_users_per_s": round(U / t_rec, 1)}
    # ---- (b) the port, whole job (every column, every user)
    port = bench.CpuPort(wl["kwargs"], u, i, ts, r)
    v, detail = port.sample(n_cols=I, n_users_rec=U, n_events_ingest=len(u))
    out["port"] = {"users_per_s": round(v, 2), **detail}
    same = sum(int(list(a) == list(b)) for a, b in zip(lists, port._rec_cache["lists"]))
    t_ing, t_fitp = detail["ingest_sec_est"], detail["fit_sec_est"]
    out["summary"] = {
        "config": "BASELINE configs[0] (ML-1M shape, nn=50), whole job, same host cores, build container",
        "cores": os.cpu_count(), "real_reference_users_per_s": out["reference"]["users_per_s"], "port_users_per_s": out["port"]["users_per_s"],
        "port_over_reference": round(out["port"]["users_per_s"] / out["reference"]["users_per_s"], 2),
        "real_reference_bulk_fit_sec": out["reference"]["bulk_fit_sec"], "port_ingest_plus_fit_sec": round(t_ing + t_fitp, 3),
        "real_reference_recommend_users_per_s": out["reference"]["recommend_users_per_s"],
        "port_recommend_users_per_s": detail["recommend_users_per_s"],
        "top10_lists_identical": f"{same}/{U}",
        "note": "the port is FASTER than the real reference (vectorised ingest, C solver without process-pool overhead), so "
                "GPU/port ratios understate GPU/reference ratios by about this factor",
    }
    path = os.path.join(ROOT, "profiles", "r3_ref_calibration.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out["summary"]))


if __name__ == "__main__":
    main()
