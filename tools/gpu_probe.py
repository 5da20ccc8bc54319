"""Timing probe of the individual kernels on synthetic BASELINE shapes (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rtrec_b200.utils.synth import synth_shape
from rtrec_b200 import device as D, _lib
from rtrec_b200.models import SLIM


def ev(fn, n=1):
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); out = None
    for _ in range(n):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out


def run(shape, nn, scale=1.0):
    t0 = time.time()
    u, i, ts, r = synth_shape(shape, scale)
    print(f"[{shape} x{scale}] generated {len(u)} events in {time.time()-t0:.1f}s", flush=True)
    m = SLIM(nn_feature_selection=nn, keep_fit_details=True)
    t0 = time.time(); m.add_interaction_arrays(u, i, ts, r); t1 = time.time()
    ms, X = ev(lambda: m.interactions.device_matrix())
    print(f"  host ingest {t1-t0:.2f}s; fold+build {ms:.1f} ms; shape {X.shape} nnz {X.nnz} nonneg {X.nonneg}", flush=True)
    ms, G = ev(lambda: D.gram(X))
    rl = np.diff(X.rptr.cpu().numpy()).astype(np.float64)
    macs = float((rl * rl).sum())
    print(f"  gram {ms:.1f} ms; {macs:.3e} MACs -> {macs/ms/1e6:.1f} G atomics/s; alg bytes {8*(macs+X.nnz)/1e9:.1f} GB -> {8*(macs+X.nnz)/ms/1e6:.0f} GB/s", flush=True)
    cfg = m.model._config(X)
    tg = torch.arange(X.n_items, dtype=torch.int32, device="cuda")
    ms, res = ev(lambda: D.solve(G, X.n_items, tg, cfg))
    st = res.stats.cpu().numpy()
    print(f"  solve nn={nn}: {ms:.1f} ms; mean iters {st[:,0].mean():.1f} draws {st[:,1].mean():.0f} gaps {st[:,2].mean():.1f} live {st[:,3].mean():.1f}", flush=True)
    ms, W = ev(lambda: D.w_merge(None, X.n_items, res))
    print(f"  w_merge+transpose {ms:.1f} ms; nnz(W) {W.nnz}", flush=True)
    del G
    users = torch.arange(X.n_users, dtype=torch.int32, device="cuda")
    for mode in (1, 0):
        ms, out = ev(lambda: D.recommend(X, users, W, 10, True, mode))
        print(f"  recommend mode={mode}: {ms:.1f} ms -> {X.n_users/ms*1e3:.0f} users/s", flush=True)
    m.model._W = W
    t0 = time.time(); m.bulk_fit(); torch.cuda.synchronize(); print(f"  bulk_fit wall {time.time()-t0:.3f}s", flush=True)
    t0 = time.time(); ids, sc, cnt = m.model.recommend_batch_device(np.arange(X.n_users), X, None, 10, True, False); print(f"  recommend_batch_device wall {time.time()-t0:.3f}s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ml1m"]
    for w in which:
        if w == "ml1m":
            run("ml1m", 50); run("ml1m", None)
        elif w == "ml20m":
            run("ml20m", 50)
        elif w == "hm":
            run("hm", 50)
    print("launches", _lib.launch_count())
