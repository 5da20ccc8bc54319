"""Host-side logic of the N>1 path on CPU: world_size-2 ``gloo`` process group.

The collectives the multi-GPU pipeline issues (``exchange_slabs``, ``all_gather_ragged``,
``gather_solve_results``) and the partition helpers are device-agnostic torch code; here they run on
CPU tensors over gloo.  The kernels they feed are covered by the ``-m gpu`` tests.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtrec_b200 import device as D
from rtrec_b200 import pipeline as P


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- slabs of unequal height: every rank ends with the whole matrix
        I = 37
        cuts = [0, 9, I]
        full = torch.arange(I * I, dtype=torch.float32).view(I, I)
        M = torch.zeros(I, I)
        M[cuts[rank]:cuts[rank + 1]] = full[cuts[rank]:cuts[rank + 1]]
        P.exchange_slabs(M, cuts)
        assert torch.equal(M, full)
        # ---- ragged all-gather
        counts = [5, 0] if world == 2 else [3] * world
        x = torch.arange(counts[rank], dtype=torch.int32) + 100 * rank
        parts = P.all_gather_ragged(x, counts)
        assert [p.tolist() for p in parts] == [list(range(100 * r, 100 * r + counts[r])) for r in range(world)]
        # ---- solver outputs of two ranks -> one result, offsets rebased
        j0, j1 = P.item_shard(11, rank, world)
        T = j1 - j0
        rng = np.random.default_rng(rank)
        cnt = rng.integers(0, 4, T).astype(np.int32)
        # rank-local append order is arbitrary (atomic cursor in all-features mode): use a shuffled layout
        order = rng.permutation(T)
        off = np.zeros(T, dtype=np.int64)
        pos = 0
        for tt in order:
            off[tt] = pos
            pos += int(cnt[tt])
        rows = np.zeros(pos, dtype=np.int32)
        vals = np.zeros(pos, dtype=np.float32)
        for tt in range(T):
            for e in range(cnt[tt]):
                rows[off[tt] + e] = 1000 * (j0 + tt) + e
                vals[off[tt] + e] = float(j0 + tt) + 0.25 * e
        res = D.SolveResult(torch.arange(j0, j1, dtype=torch.int32), torch.from_numpy(off), torch.from_numpy(cnt),
                            torch.from_numpy(rows), torch.from_numpy(vals), None,
                            torch.full((T, 4), rank, dtype=torch.int32), True, pos)
        g = P.gather_solve_results(res, world)
        assert g.targets.tolist() == list(range(11))
        assert int(g.n_pairs) == int(g.rows.numel())
        for tt, j in enumerate(g.targets.tolist()):
            o, c = int(g.off[tt]), int(g.cnt[tt])
            assert g.rows[o:o + c].tolist() == [1000 * j + e for e in range(c)]
            assert g.vals[o:o + c].tolist() == [float(j) + 0.25 * e for e in range(c)]
        assert g.stats.shape == (11, 4)
        # ---- query cuts: identical on every rank, contiguous, balanced by row length
        rptr = torch.tensor([0, 100, 100, 101, 150, 150, 400, 401], dtype=torch.int32)
        users = torch.tensor([6, 0, 2, 5, 3, 1], dtype=torch.int32)
        qc = P.query_cuts(rptr, users, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, qc)
        assert all(c == qc for c in gathered)
        assert qc[0] == 0 and qc[-1] == 6 and all(a <= b for a, b in zip(qc[:-1], qc[1:]))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_item_shard_partitions_the_range():
    for n_items, world in [(11, 2), (26744, 8), (5, 8), (1, 2)]:
        cuts = [P.item_shard(n_items, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n_items
        assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1
    assert P.item_shard(10, 0, 1) == (0, 10)


def test_world2_collectives_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
