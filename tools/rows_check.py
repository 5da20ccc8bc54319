#!/usr/bin/env python
"""Multi-GPU check of the owner-rows fit (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/rows_check.py

Every rank builds the same synthetic X.  W is fitted once on a single GPU (dense Gram matrix), once with the
owner-rows pipeline (block-cyclic Gram rows completed out of peer memory, solver gathering foreign rows over NVLink)
and, for comparison, once with the whole-triangle exchange; after the all-gather of the solver outputs the assembled
W must be identical bit for bit in all three.  Repeats the owner-rows fit to exercise buffer reuse across fits.
Prints one line per stage (flushed) so that a hang can be located.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
from rtrec_b200.utils.synth import synth_events
from rtrec_b200 import device as D, pipeline as P
from rtrec_b200.models.internal.slim_elastic import SLIMElastic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
T0 = time.time()


def say(msg):
    print(f"[{time.time() - T0:6.2f}s rank {rank}] {msg}", file=sys.stderr, flush=True)


def same_w(a, b):
    return a.nnz == b.nnz and torch.equal(a.wptr, b.wptr) and torch.equal(a.widx[:a.nnz], b.widx[:b.nnz]) \
        and torch.equal(a.wval[:a.nnz], b.wval[:b.nnz])


torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
say("process group up")
U, I, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (6040, 3706, 1000000)))
u, i, ts, r = synth_events(U, I, N, seed=21, rating="int")
i = np.random.default_rng(5).permutation(I)[i]        # item ids uncorrelated with popularity
X = D.DeviceMatrix.from_scipy(sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I)))
for nn in (50, None):
    cfg = SLIMElastic({"nn_feature_selection": nn} if nn else {})._config(X)
    tg_all = torch.arange(I, dtype=torch.int32, device="cuda")
    if nn is None:
        tg_all = tg_all[::37].contiguous()             # all-features solves are slow: a sample of the targets
    G1 = D.gram_full(X)
    W1 = D.w_merge(None, I, D.solve(G1, I, tg_all, cfg))
    del G1
    torch.cuda.synchronize()
    say(f"nn={nn}: single-GPU fit done, nnz(W) = {W1.nnz}")
    for rep in range(3):
        if nn is None:
            # sampled targets: owner-rows Gram, explicit target list restricted to the own blocks
            slabs = P.PeerSlabs.get(I, rank, world)
            if slabs is None:
                say("peer slabs unavailable on this node: nothing to check")
                break
            slabs.barrier()
            rank_of, orig_of = D.gram_lower_blocks(X, rank, world, slabs.own)
            slabs.barrier()
            D.gram_pull_cols(slabs.ptrs, rank, I)
            rows = slabs.rows_ptrs()
            D.gram_unpermute_rows(slabs.own, rows[rank], slabs.rows_alloc, I, rank_of)
            G = D.GramRows(rows, D.gram_row_slots(rank_of, world), D.slab_ld(I))
            mine = D.block_targets(orig_of, I, rank, world)
            keep = torch.isin(mine, tg_all)
            slabs.barrier()
            res = D.solve(G, I, mine[keep].contiguous(), cfg)
        else:
            res = P.fit_owner_rows(X, cfg, rank=rank, world=world, marks=lambda name: say(f"  rows: {name} queued") if rep == 0 else None)
            if res is None:
                say("peer slabs unavailable on this node: nothing to check")
                break
        W = D.w_merge(None, I, P.gather_solve_results(res, world))
        torch.cuda.synchronize()
        ok = same_w(W, W1)
        say(f"nn={nn} rep {rep}: owner-rows W equal to the single-GPU W = {ok}")
        assert ok
    if nn is not None:
        res, _ = P.fit_sharded(X, cfg, rank=rank, world=world, strided=True)
        W = D.w_merge(None, I, P.gather_solve_results(res, world))
        torch.cuda.synchronize()
        ok = same_w(W, W1)
        say(f"nn={nn}: whole-triangle exchange W equal to the single-GPU W = {ok}")
        assert ok
# ---- the public API in SPMD mode: SLIM(distributed=True) on every rank against an ordinary single-GPU model
from rtrec_b200.models import SLIM
users = list(range(U)) + [U + 5]          # + a cold user
out = {}
for flag in (False, True):
    m = SLIM(nn_feature_selection=50, distributed=flag)
    m.add_interaction_arrays(u, i, ts, r.astype(np.float64))
    m.bulk_fit()
    out[flag] = (m.recommend_batch(users, top_k=10), m.model.item_similarity)
    say(f"API distributed={flag}: fitted, nnz(W) = {out[flag][1].nnz}")
dW = (out[False][1] != out[True][1]).nnz
same = sum(int(a == b) for a, b in zip(out[False][0], out[True][0]))
say(f"API: W entries differing = {dW}, identical top-10 lists = {same}/{len(users)}")
assert dW == 0 and same == len(users)
dist.barrier()
say("done")
dist.destroy_process_group()
