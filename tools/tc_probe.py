#!/usr/bin/env python
"""Scoring paths at one bench shape on one GPU: exact dense-tile kernel vs tensor-core path, timings with CUDA events,
agreement of the lists, fallback share.  Run under `ncu --metrics gpu__time_duration.sum` for per-kernel times."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import WORKLOADS, load_events
from rtrec_b200 import device as D, pipeline as P
from rtrec_b200._lib import RT_TOPK_SPARSE
from rtrec_b200.models.internal.slim_elastic import SLIMElastic


def main():
    wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "ml20m"]
    u, i, ts, r = load_events(wl["shape"])
    decay = wl["kwargs"].get("decay_in_days")
    rate = None if decay is None else 1.0 - (np.log(2) / decay)
    st = P.fold_events(P.empty_store(), D.to_dev(u.astype(np.int32)), D.to_dev(i.astype(np.int32)), D.to_dev(ts), D.to_dev(r),
                       upsert=False, min_value=-5, max_value=10, decay_rate=rate)
    X = P.build_matrix(st, decay_rate=rate)
    op = SLIMElastic(wl["kwargs"])
    G = D.gram_full(X)
    res = D.solve(G, X.n_items, torch.arange(X.n_items, dtype=torch.int32, device="cuda"), op._config(X))
    del G
    W = D.w_merge(None, X.n_items, res)
    users = torch.arange(X.n_users, dtype=torch.int32, device="cuda")
    out = {"n_users": X.n_users, "nnz_W": W.nnz}

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        ts_ = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = fn(); e1.record(); torch.cuda.synchronize()
            ts_.append(e0.elapsed_time(e1))
        return float(np.median(ts_)), o

    D.set_option("score_tc", 0)
    out["exact_ms"], ex = timed(lambda: D.recommend(X, users, W, 10, True, RT_TOPK_SPARSE))
    D.set_option("score_tc", 1)
    pk = D.tc_pack(W)
    out["tc_pack"] = None if pk is None else {"n_heavy": pk.n_heavy, "w_nonneg": pk.w_nonneg, "x": D.values_bf16_exact(X)}
    if pk is not None:
        out["tc_ms"], tc = timed(lambda: D.recommend_tc(X, users, W, 10, True, RT_TOPK_SPARSE))
        full = D.recommend_tc(X, users, W, 10, True, RT_TOPK_SPARSE, debug_scores=False)
        # fallback share: run the C entry once more by hand to read the flags
        t = torch
        Q = X.n_users
        bufs = [D.empty(Q * 10, t.int32), D.empty(Q * 10, t.float32), D.empty(Q, t.int32), D.empty(Q * 32, t.int32),
                D.empty(Q * 32, t.float32), D.empty(Q, t.int32), D.empty(Q, t.int32)]
        from rtrec_b200 import _lib
        x_nonneg, x_exact = D.values_bf16_exact(X)

        def raw():
            _lib.check(_lib.load().rt_slim_recommend_tc(D.ptr(X.rptr), D.ptr(X.ridx), D.ptr(X.rval), D.ptr(users), Q, D.ptr(W.wrptr),
                                                        D.ptr(W.wridx), D.ptr(W.wrval), D.ptr(pk.heavy_of), pk.n_heavy, D.ptr(pk.bt), D.ptr(pk.wd),
                                                        W.n_items, 10, 1, RT_TOPK_SPARSE, 1 if x_exact else 3, D.ptr(bufs[3]),
                                                        D.ptr(bufs[4]), D.ptr(bufs[5]), D.ptr(bufs[0]), D.ptr(bufs[1]), D.ptr(bufs[2]),
                                                        D.ptr(bufs[6]), None, D.stream_ptr()), "rt_slim_recommend_tc")
        out["tc_kernels_only_ms"], _ = timed(raw)
        out["fallback_users"] = int((bufs[6] != 0).sum())
        out["tc_list_len_mean"] = float((bufs[3].view(Q, 32) >= 0).sum(dim=1).float().mean())
        same = (tc[0] == ex[0]).all(dim=1)
        out["lists_identical"] = f"{int(same.sum())}/{Q}"
        out["cnt_equal"] = bool(torch.equal(tc[2], ex[2]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
