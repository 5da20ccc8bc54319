"""Device time of the chunked, queued scoring used by SLIMElastic.recommend_lists, by chunk size (ML-20M shape).
GPU box only:  python tools/chunk_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import bench
    from rtrec_b200 import device as D, pipeline as P
    from rtrec_b200._lib import RT_TOPK_SPARSE
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    wl = bench.WORKLOADS["ml20m"]
    u, i, ts, r = bench.load_events(wl["shape"])
    U = int(u.max()) + 1
    st = P.fold_events(P.empty_store(), D.to_dev(u.astype(np.int32)), D.to_dev(i.astype(np.int32)), D.to_dev(ts), D.to_dev(r),
                       upsert=False, min_value=-5, max_value=10, decay_rate=None)
    X = P.build_matrix(st, decay_rate=None)
    op = SLIMElastic(wl["kwargs"])
    op.fit(X)
    W = op._W
    users = torch.arange(U, dtype=torch.int32, device="cuda")
    only = int(os.environ.get("CHUNK_ONLY", "0"))
    for chunk in ([only] if only else [U, 65536, 32768, 16384, 8192]):
        for rep in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for a in range(0, U, chunk):
                D.recommend(X, users[a:a + chunk], W, 10, True, RT_TOPK_SPARSE, no_sync=True)
            e1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
        print(f"chunk {chunk}: device {e0.elapsed_time(e1):.2f} ms, host launch {1e3 * (t1 - t0):.2f} ms", flush=True)


if __name__ == "__main__":
    main()
