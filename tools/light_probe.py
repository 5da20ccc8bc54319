"""Distribution of the light-row work per user behind recommend_tcfix_kernel (ML-20M shape): staged light items and light
entries per user.  GPU box only:  python tools/light_probe.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import bench
    from rtrec_b200 import device as D, pipeline as P
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    name = sys.argv[1] if len(sys.argv) > 1 else "ml20m"
    wl = bench.WORKLOADS[name]
    u, i, ts, r = bench.load_events(wl["shape"])
    decay = wl["kwargs"].get("decay_in_days")
    rate = None if decay is None else 1.0 - (np.log(2) / decay)
    st = P.fold_events(P.empty_store(), D.to_dev(u.astype(np.int32)), D.to_dev(i.astype(np.int32)), D.to_dev(ts), D.to_dev(r),
                       upsert=False, min_value=-5, max_value=10, decay_rate=rate)
    X = P.build_matrix(st, decay_rate=rate)
    op = SLIMElastic(wl["kwargs"])
    op.fit(X)
    W = op._W
    pk = D.tc_pack(W)
    print("W nnz", W.nnz, "heavy rows", None if pk is None else pk.n_heavy)
    if pk is None:
        return
    length = (W.wrptr[1:] - W.wrptr[:-1]).to(torch.int64)
    light = torch.where(pk.heavy_of[:W.n_items] < 0, length, torch.zeros_like(length))
    per_entry = light[X.ridx[:X.nnz].long()]
    rows = torch.repeat_interleave(torch.arange(X.n_users, device="cuda"), (X.rptr[1:] - X.rptr[:-1]).long())
    n_light = torch.zeros(X.n_users, dtype=torch.int64, device="cuda").index_add_(0, rows, per_entry)
    n_rows = torch.zeros(X.n_users, dtype=torch.int64, device="cuda").index_add_(0, rows, (per_entry > 0).long())
    for nm, v in (("light entries", n_light), ("light items", n_rows)):
        q = torch.quantile(v.double(), torch.tensor([0.5, 0.75, 0.9, 0.95, 0.99, 1.0], dtype=torch.float64, device="cuda"))
        print(nm, "p50/75/90/95/99/max", [int(x) for x in q.tolist()], "mean", float(v.double().mean()))
    for cap, rows_cap in ((512, 192), (640, 256), (1024, 384), (1280, 512)):
        ok = ((n_light <= cap) & (n_rows <= rows_cap)).double().mean()
        print(f"fit (entries <= {cap}, items <= {rows_cap}): {float(ok):.4f} of the users")


if __name__ == "__main__":
    main()
