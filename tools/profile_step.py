"""Runs a few device-resident steps of the hot path for ncu (see profiles/README.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from rtrec_b200 import device as D, pipeline as P
from rtrec_b200.models.internal.slim_elastic import SLIMElastic
from rtrec_b200._lib import RT_TOPK_SPARSE

wl = sys.argv[1] if len(sys.argv) > 1 else "ml20m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if len(sys.argv) > 3:
    D.set_option("score_impl", int(sys.argv[3]))
shape, kwargs, desc = (bench.WORKLOADS[wl][k] for k in ("shape", "kwargs", "desc"))
u, i, ts, r = bench.load_events(shape)
U = int(u.max()) + 1
op = SLIMElastic(kwargs)
decay = kwargs.get("decay_in_days")
rate = None if decay is None else 1.0 - (np.log(2) / decay)
du, di, dts, dd = D.to_dev(u.astype(np.int32)), D.to_dev(i.astype(np.int32)), D.to_dev(ts), D.to_dev(r)
users = torch.arange(U, dtype=torch.int32, device="cuda")
for s in range(steps):
    st = P.fold_events(P.empty_store(), du, di, dts, dd, upsert=False, min_value=-5, max_value=10, decay_rate=rate)
    X = P.build_matrix(st, decay_rate=rate)
    res, jr = P.fit_sharded(X, op._config(X))
    W = D.w_merge(None, X.n_items, res)
    ids, sc, cnt = D.recommend(X, users, W, 10, True, RT_TOPK_SPARSE)
    torch.cuda.synchronize()
print("done", wl, steps, int(cnt.sum().item()))
