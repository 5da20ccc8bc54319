#!/usr/bin/env python
"""Kernel A/B micro-benchmark on one GPU: times alternative implementations of the hot kernels on a
named synthetic shape and checks that they agree.  Development tool (not the bench contract)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import WORKLOADS, load_events
from rtrec_b200 import device as D, pipeline as P
from rtrec_b200._lib import RT_TOPK_SPARSE, RT_TOPK_DENSE
from rtrec_b200.models.internal.slim_elastic import SLIMElastic


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ml20m")
    ap.add_argument("--what", default="gram,score")
    args = ap.parse_args()
    shape, kwargs, desc = (WORKLOADS[args.workload][k] for k in ("shape", "kwargs", "desc"))
    u, i, ts, r = load_events(shape)
    decay = kwargs.get("decay_in_days")
    rate = None if decay is None else 1.0 - (np.log(2) / decay)
    du, di = D.to_dev(u.astype(np.int32)), D.to_dev(i.astype(np.int32))
    dts, dd = D.to_dev(ts), D.to_dev(r)
    st = P.fold_events(P.empty_store(), du, di, dts, dd, upsert=False, min_value=-5, max_value=10, decay_rate=rate)
    X = P.build_matrix(st, decay_rate=rate)
    I, U = X.n_items, X.n_users
    res = {"workload": args.workload, "U": U, "I": I, "nnz": X.nnz}
    what = args.what.split(",")
    G = torch.zeros((I, I), dtype=torch.float32, device="cuda")
    if "gram" in what:
        def g1():
            G.zero_(); D.gram(X, out=G); return None
        ms1, _ = timeit(g1, reps=1, warm=0)
        G1 = G.clone()
        res["gram_v1_ms"] = ms1
        box = {}
        def g3a():
            box["L"] = D.gram_lower(X); return None
        ms3a, _ = timeit(g3a)
        def g3b():
            D.gram_finish(box["L"], out=G); return None
        ms3b, _ = timeit(g3b)
        res["gram_v3_lower_ms"] = ms3a; res["gram_v3_finish_ms"] = ms3b
        for sl, rg in ((1728, 3), (1728, 4), (1152, 3), (1152, 4), (2304, 2), (2304, 3)):
            D.set_option("gram_slice", sl); D.set_option("gram_ranges", rg)
            msv, _ = timeit(g3a)
            D.gram_finish(box["L"], out=G)
            res[f"gram_v3_lower_slice{sl}_r{rg}"] = {"ms": msv, "equal": bool(torch.equal(G1, G))}
        D.set_option("gram_slice", 0); D.set_option("gram_ranges", 0)
        # experimental: segment-length guards in gram_lower_kernel
        for mode in (0, 1, 2):
            D.set_option("gram_adapt", mode)
            msa, _ = timeit(g3a)
            D.gram_finish(box["L"], out=G)
            res[f"gram_v3_lower_adapt{mode}"] = {"ms": msa, "equal": bool(torch.equal(G1, G))}
        D.set_option("gram_adapt", 2)
        D.gram_finish(D.gram_lower(X), out=G)
        res["gram_v3_equal"] = bool(torch.equal(G1, G))
        res["gram_v3_maxdiff"] = float((G1 - G).abs().max())
        box.clear()
        del G1
    else:
        D.gram(X, out=G)
    op = SLIMElastic(kwargs)
    cfg = op._config(X, into_empty_w=False)   # warp and CTA solver are compared pair by pair below
    tg = torch.arange(0, I, dtype=torch.int32, device="cuda")
    D.set_option("solve_impl", 1)
    ms1, sol1 = timeit(lambda: D.solve(G, I, tg, cfg), reps=2)
    D.set_option("solve_impl", 2)
    ms, sol = timeit(lambda: D.solve(G, I, tg, cfg), reps=2)
    res["solve_block_ms"] = ms1
    res["solve_ms"] = ms
    res["solve_rows_equal"] = bool(torch.equal(sol1.rows, sol.rows))
    res["solve_vals_maxdiff"] = float((sol1.vals - sol.vals).abs().max())
    res["solve_cols_differ"] = int(((sol1.vals.view(-1, cfg.nn) != sol.vals.view(-1, cfg.nn)).any(dim=1)).sum()) if cfg.nn > 0 else None
    res["solve_iters_differ"] = int((sol1.stats[:, 0] != sol.stats[:, 0]).sum())
    del G
    W = D.w_merge(None, I, sol)
    res["nnz_W"] = W.nnz
    users = torch.arange(U, dtype=torch.int32, device="cuda")
    if "score" in what:
        for mode, name in ((RT_TOPK_SPARSE, "sparse"), (RT_TOPK_DENSE, "dense")):
            D.set_option("score_impl", 1)
            ms1, o1 = timeit(lambda: D.recommend(X, users, W, 10, True, mode))
            D.set_option("score_impl", 2)
            ms2, o2 = timeit(lambda: D.recommend(X, users, W, 10, True, mode))
            res[f"score_{name}_v1_ms"] = ms1; res[f"score_{name}_v2_ms"] = ms2
            D.set_option("score_impl", 3)
            for min_row in (768, 1536):
                D.PACK_MIN_ROW = min_row
                W.packs = None
                ms3, o3 = timeit(lambda: D.recommend(X, users, W, 10, True, mode))
                pk = D.score_pack(W, 0, I)
                res[f"score_{name}_v3_min{min_row}"] = {"ms": ms3, "n_heavy": pk.n_heavy, "n_groups": pk.n_groups,
                                                         "equal": bool(torch.equal(o3[0], o2[0]) and torch.equal(o3[1], o2[1]) and torch.equal(o3[2], o2[2]))}
            D.PACK_MIN_ROW = 768
            res[f"score_{name}_equal"] = bool(torch.equal(o1[0], o2[0]) and torch.equal(o1[1], o2[1]) and torch.equal(o1[2], o2[2]))
            if not res[f"score_{name}_equal"]:
                bad = (o1[0] != o2[0]).any(dim=1).nonzero().flatten()
                res[f"score_{name}_nbad"] = int(bad.numel())
                if bad.numel():
                    b = int(bad[0])
                    res[f"score_{name}_example"] = {"user": b, "v1": o1[0][b].tolist(), "v2": o2[0][b].tolist(),
                                                    "s1": o1[1][b].tolist(), "s2": o2[1][b].tolist()}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
