#!/usr/bin/env python
"""CPU experiment behind the parity bar at the full ML-20M shape (tests/helpers.py::assert_w_parity_at_scale).

Runs the exact port of the reference path (oracle/slim_oracle.c, float32 residual form, bit-exact with sklearn) and the
CPU model of the device algorithm (oracle/gram_model.c, Gram form, float64 solver state) on the same sampled target
columns of the synthetic ML-20M-shaped matrix, with the same candidates, and prints per column: stored entries, sweep
counts of both, relative error against the column maximum.  Test infrastructure (executes oracle/); ~1 minute, 6 GB.
Output of the round-1 run: profiles/r2k_c2_parity_cpu.log.
"""
import sys, time, numpy as np, scipy.sparse as sp
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import slim_oracle as so
from rtrec_b200.utils.synth import synth_shape
t0=time.time()
u,i,ts,r = synth_shape("ml20m")
U,I = int(u.max())+1, int(i.max())+1
X = sp.csc_matrix((r.astype(np.float32),(u,i)),shape=(U,I)); X.sort_indices()
print("X built", time.time()-t0, flush=True)
rng=np.random.default_rng(7)
cols=np.unique(np.concatenate([rng.choice(I,40,replace=False), rng.choice(2000,24,replace=False)])).astype(np.int32)
res, sel, st = so.fit_columns(X, cols, 50, n_threads=8)
print("exact port done", time.time()-t0, "iters", st[:,0].tolist(), flush=True)
G = so.gram_model_gram(X)
print("gram done", time.time()-t0, flush=True)
res2, sel2, st2 = so.gram_model_fit_columns(X, cols, 50, sel_in=sel, G=G)
print("gram model done", time.time()-t0, flush=True)
np.savez(os.environ.get("C2_PARITY_OUT", "/tmp/c2_parity.npz"), cols=cols, sel=sel, st=st, st2=st2, v1=np.array([np.pad(v,(0,50-len(v))) for _,v in res]), v2=np.array([np.pad(v,(0,50-len(v))) for _,v in res2]),
         r1=np.array([np.pad(rr,(0,50-len(rr)),constant_values=-1) for rr,_ in res]), r2=np.array([np.pad(rr,(0,50-len(rr)),constant_values=-1) for rr,_ in res2]))
for k,j in enumerate(cols):
    (ra,va),(rb,vb)=res[k],res2[k]
    same_rows = np.array_equal(ra,rb)
    m = max(np.abs(va).max() if len(va) else 0, 1e-30)
    err = np.abs(va-vb).max()/m if same_rows and len(va) else float('nan')
    print(f"col {j:6d} nnz_col {X.indptr[j+1]-X.indptr[j]:6d} iters {st[k,0]:3d}/{st2[k,0]:3d} rows_equal {same_rows} relerr {err:.3e} maxw {m:.3e}")
