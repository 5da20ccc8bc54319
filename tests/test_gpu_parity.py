"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle and the golden vectors.

Bars (BASELINE.json north_star 6): indices / bookkeeping / store matrices bit-exact; W within 1e-4
of the column's largest coefficient (columns whose stop decision flips by one sweep: 1e-3, at most
2 % of columns); top-k lists equal up to score ties.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import slim_oracle as so
from rtrec_b200.utils.synth import synth_events
from tests.helpers import assert_w_parity, csc_from, topk_consistent, w_from

pytestmark = pytest.mark.gpu

FIT_CASES = [
    ("slim_all_int", {}),
    ("slim_nn20_int", {"nn_feature_selection": 20}),
    ("slim_nn20_cont", {"nn_feature_selection": 20}),
    ("slim_nn20_decay", {"nn_feature_selection": 20}),
    ("slim_all_decay_partial", {}),
    ("slim_nn20_partial", {"nn_feature_selection": 20}),
    ("slim_all_strids_fit", {}),
]


def _exact(A, B):
    A = sp.csc_matrix(A); B = sp.csc_matrix(B)
    A.sort_indices(); B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data, B.data))


# ------------------------------------------------------------------------------------------ store
@pytest.mark.parametrize("k", range(4))
def test_store_matches_reference_bit_exact(golden, k):
    from rtrec_b200.utils.interactions import UserItemInteractions
    z = golden(f"store_{k}")
    ev = z["events"]
    decay = None if z["decay"] < 0 else int(z["decay"])
    ups = bool(z["upsert"])
    st = UserItemInteractions(min_value=-5, max_value=10, decay_in_days=decay)
    u, i = ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64)
    for a, b in ((0, 1500), (1500, 1501), (1501, 2600)):
        st.add_interactions_batch(u[a:b], i[a:b], ev[a:b, 2], ev[a:b, 3], upsert=ups)
        st.device_matrix()  # force a fold per chunk: state must carry over
    for r in range(2600, len(u)):  # the scalar API interleaves with the batched one
        st.add_interaction(int(u[r]), int(i[r]), float(ev[r, 2]), float(ev[r, 3]), upsert=ups)
    X = st.to_csc()
    ref = sp.csc_matrix((z["X_data"], z["X_indices"], z["X_indptr"]), shape=tuple(z["X_shape"]))
    assert _exact(X, ref)
    assert st.max_timestamp == float(z["max_timestamp"])
    assert (st.max_user_id, st.max_item_id) == (int(z["max_user_id"]), int(z["max_item_id"]))
    Xs = st.to_csc(z["sel_items"].tolist())
    refs = sp.csc_matrix((z["Xs_data"], z["Xs_indices"], z["Xs_indptr"]), shape=tuple(z["X_shape"]))
    assert _exact(Xs, refs)
    R = sp.csr_matrix(st.to_csr(z["sel_users"].tolist()))
    refr = sp.csr_matrix((z["R_data"], z["R_indices"], z["R_indptr"]), shape=tuple(z["X_shape"]))
    assert _exact(R.tocsc(), refr.tocsc())
    assert list(st.hot_items.get_freq_items(20)) == z["hot_items"].tolist()
    # CSR and CSC on the device describe the same matrix
    dm = st.device_matrix()
    assert _exact(dm.to_scipy_csr().tocsc(), dm.to_scipy_csc())


def test_store_edge_cases():
    from rtrec_b200.utils.interactions import UserItemInteractions
    st = UserItemInteractions()
    assert st.to_csc().shape == (1, 1) and st.to_csc().nnz == 0
    st.add_interaction(1, 10, 1000.0, 5.0)
    st.add_interaction(1, 10, 1000.0, -5.0)
    assert st.get_user_item_rating(1, 10) == 0.0           # tests/utils/test_interactions.py:31-36
    assert 10 in st.get_user_items(1)
    assert st.to_csc().nnz == 1                              # explicit zero stays
    st.add_interaction(1, 10, 1001.0, 30.0)
    assert st.get_user_item_rating(1, 10) == 10.0           # clipped at max_value
    st.add_interaction(2, 3, 1002.0, 7.0, upsert=True)
    st.add_interaction(2, 3, 1003.0, 70.0, upsert=True)
    assert st.get_user_item_rating(2, 3) == 70.0            # upsert is not clipped
    assert set(st.get_users_by_items([10])) == {1}
    assert st.get_users_by_items([99]) == []


# ------------------------------------------------------------------------------------------ gram
def test_gram_rows_exact_on_integer_ratings():
    from rtrec_b200 import device as D
    U, I, N = 3000, 700, 150000
    u, i, ts, r = synth_events(U, I, N, seed=3, rating="int")
    X = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))
    G_ref = so.gram_model_gram(X)
    G = D.gram(D.DeviceMatrix.from_scipy(X)).cpu().numpy()
    assert np.array_equal(G, G_ref)
    assert np.array_equal(G, G.T)


@pytest.mark.parametrize("shuffle", [False, True])
def test_gram_v3_equals_model_and_shards(shuffle):
    """Popularity-ranked lower-triangle Gram (rt_gram_lower/finish): exact on integer ratings for any item id
    order, symmetric, and the row slabs of a 3-way split assemble to the same matrix."""
    from rtrec_b200 import device as D
    U, I, N = 3000, 2100, 150000
    u, i, ts, r = synth_events(U, I, N, seed=4, rating="int")
    if shuffle:
        i = np.random.default_rng(0).permutation(I)[i]
    X = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))
    G_ref = so.gram_model_gram(X)
    dX = D.DeviceMatrix.from_scipy(X)
    G = D.gram_full(dX).cpu().numpy()
    assert np.array_equal(G, G_ref)
    assert np.array_equal(G, G.T)
    t = D.torch()
    parts = [D.gram_lower(dX, part=p, n_parts=3) for p in range(3)]
    cuts = parts[0].cuts
    assert cuts[0] == 0 and cuts[-1] == I and all(parts[p].cuts == cuts for p in range(3))
    for p in (1, 2):
        parts[0].Gp[cuts[p]:cuts[p + 1]].copy_(parts[p].Gp[cuts[p]:cuts[p + 1]])
        assert t.equal(parts[p].rank_of, parts[0].rank_of)
    G3 = D.gram_finish(parts[0]).cpu().numpy()
    assert np.array_equal(G3, G_ref)
    # continuous ratings: agreement to fp32 rounding of the differently ordered sums
    Xc = sp.csc_matrix((np.random.default_rng(1).uniform(0.5, 5.0, len(u)).astype(np.float32), (u, i)), shape=(U, I))
    Gc = D.gram_full(D.DeviceMatrix.from_scipy(Xc)).cpu().numpy()
    ref = (Xc.T.astype(np.float64) @ Xc.astype(np.float64)).toarray()
    assert np.abs(Gc - ref).max() <= 2e-6 * np.abs(ref).max()


def test_gram_fused_pull_mirror_equals_single_pass():
    """rt_gram_finish_p2p on one GPU: three slabs in three separate buffers (standing in for the IPC mappings
    of peer ranks) are pulled, mirrored and un-permuted in one call; from every rank's point of view the
    result is the single-pass Gram matrix, exactly.  Also covers the raw-buffer form of rt_gram_lower."""
    import ctypes as C
    from rtrec_b200 import _lib, device as D
    lib = _lib.load()
    U, I, N = 2500, 1777, 120000
    u, i, ts, r = synth_events(U, I, N, seed=9, rating="int")
    i = np.random.default_rng(2).permutation(I)[i]
    X = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))
    dX = D.DeviceMatrix.from_scipy(X)
    G_ref = D.gram_full(dX).cpu().numpy()
    n_parts = 3
    bufs, handles = [], []
    try:
        for p in range(n_parts):
            ptr_, h = C.c_void_p(0), (C.c_uint8 * 64)()
            _lib.check(lib.rt_ipc_alloc(4 * I * D.slab_ld(I), C.byref(ptr_), h), "rt_ipc_alloc")
            bufs.append(int(ptr_.value))
        parts = [D.gram_lower(dX, part=p, n_parts=n_parts, raw_ptr=bufs[p]) for p in range(n_parts)]
        assert all(q.cuts == parts[0].cuts for q in parts)
        # the protocol of the real ranks, run in lock step: everybody phase 0, (barrier), everybody phase 1, then
        # each rank un-permutes its own copy (phase 2)
        t = D.torch()
        arr = (C.c_void_p * n_parts)(*[C.c_void_p(b) for b in bufs])
        cuts = (C.c_int32 * (n_parts + 1))(*parts[0].cuts)
        outs = [t.empty((I, I), dtype=t.float32, device="cuda") for _ in range(n_parts)]
        for phase in (0, 1, 2):
            for me in range(n_parts):
                _lib.check(lib.rt_gram_finish_p2p(I, arr, n_parts, me, cuts, D.slab_ld(I), D.ptr(parts[me].rank_of),
                                                  D.ptr(parts[me].orig_of), D.ptr(outs[me]), I, phase, D.stream_ptr()))
        for me in range(n_parts):
            assert np.array_equal(outs[me].cpu().numpy(), G_ref), f"rank {me}"
    finally:
        D.torch().cuda.synchronize()
        for b in bufs:
            lib.rt_ipc_free(C.c_void_p(b))


@pytest.mark.parametrize("n_parts", [1, 2, 3, 8])
def test_gram_owner_rows_equals_single_pass(n_parts):
    """The owner-rows layout of the multi-GPU fit on one GPU: n_parts slabs in separate buffers (standing in for the
    IPC mappings of peer ranks), run in lock step like the real ranks -- everybody rt_gram_lower_blocks, (barrier),
    everybody rt_gram_pull_cols + rt_gram_unpermute_rows, (barrier), everybody rt_slim_solve_rows on its own targets.
    Every assembled row equals the row of the single-pass Gram matrix bit for bit, the union of the parts' targets is
    every item exactly once, and the solver output (candidates, coefficients, sweep counts) through the row slots
    equals the single-GPU solve on the dense matrix, for the warp kernel (nn = 20), the CTA kernel (all features on a
    sampled target list) and with items that nobody rated (all-zero rows)."""
    import ctypes as C
    import torch
    from rtrec_b200 import _lib, device as D
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    lib = _lib.load()
    U, I, N = 2500, 1777, 120000          # 28 blocks of 64 rows, the last one partial
    u, i, ts, r = synth_events(U, I, N, seed=9, rating="int")
    i = np.random.default_rng(2).permutation(I)[i]
    keep = (i % 89) != 3
    X = sp.csc_matrix((r[keep].astype(np.float32), (u[keep], i[keep])), shape=(U, I))
    dX = D.DeviceMatrix.from_scipy(X)
    G_ref = D.gram_full(dX)
    ld = D.slab_ld(I)
    rows_alloc, _ = D.gram_block_rows(I, n_parts, 0)
    slabs = []
    row_bufs = [torch.empty((rows_alloc, ld), dtype=torch.float32, device="cuda") for _ in range(n_parts)]
    rows = [int(b.data_ptr()) for b in row_bufs]
    try:
        for p in range(n_parts):
            ptr_, h = C.c_void_p(0), (C.c_uint8 * 64)()
            _lib.check(lib.rt_ipc_alloc(4 * rows_alloc * ld, C.byref(ptr_), h), "rt_ipc_alloc")
            slabs.append(int(ptr_.value))
        ro = [D.gram_lower_blocks(dX, p, n_parts, slabs[p]) for p in range(n_parts)]
        for p in range(n_parts):
            assert torch.equal(ro[p][0], ro[0][0]) and torch.equal(ro[p][1], ro[0][1])
        rank_of, orig_of = ro[0]
        for p in range(n_parts):
            D.gram_pull_cols(slabs, p, I)
        for p in range(n_parts):
            D.gram_unpermute_rows(slabs[p], rows[p], rows_alloc, I, rank_of)
        slots = D.gram_row_slots(rank_of, n_parts)
        torch.cuda.synchronize()
        # rows: G_ref[i] == buffer[slot >> 24][slot & 0xffffff]
        sl = slots.cpu().numpy()
        Gr = G_ref.cpu().numpy()
        bufs = [b.cpu().numpy() for b in row_bufs]
        got = np.stack([bufs[s >> 24][s & 0xffffff, :I] for s in sl])
        assert np.array_equal(got, Gr)
        # targets: a partition of the items
        tgs = [D.block_targets(orig_of, I, p, n_parts) for p in range(n_parts)]
        allt = np.concatenate([x.cpu().numpy() for x in tgs])
        assert np.array_equal(np.sort(allt), np.arange(I))
        GR = D.GramRows(rows, slots, ld)
        for nn, sample in ((20, None), (None, 40)):
            cfg = SLIMElastic({"nn_feature_selection": nn} if nn else {})._config(dX)
            for p in range(n_parts):
                tg = tgs[p] if sample is None else tgs[p][:: max(1, int(tgs[p].numel()) // sample)].contiguous()
                a = D.solve(G_ref, I, tg, cfg, want_sel=nn is not None)
                b = D.solve(GR, I, tg, cfg, want_sel=nn is not None)
                assert torch.equal(a.cnt, b.cnt) and torch.equal(a.stats, b.stats)
                if nn is not None:
                    n = int(a.n_pairs)
                    assert torch.equal(a.off, b.off) and torch.equal(a.sel, b.sel)
                    assert torch.equal(a.rows[:n], b.rows[:n]) and torch.equal(a.vals[:n], b.vals[:n])
                else:
                    # all-features mode appends each column at an atomic cursor: compare column by column
                    ao, bo, cn = a.off.cpu().numpy(), b.off.cpu().numpy(), a.cnt.cpu().numpy()
                    ar, br = a.rows.cpu().numpy(), b.rows.cpu().numpy()
                    av, bv = a.vals.cpu().numpy(), b.vals.cpu().numpy()
                    for q in range(len(cn)):
                        assert np.array_equal(ar[ao[q]:ao[q] + cn[q]], br[bo[q]:bo[q] + cn[q]])
                        assert np.array_equal(av[ao[q]:ao[q] + cn[q]], bv[bo[q]:bo[q] + cn[q]])
    finally:
        D.torch().cuda.synchronize()
        for b in slabs:
            lib.rt_ipc_free(C.c_void_p(b))


def test_gram_adaptive_variant_equals_default():
    """rt_set_option("gram_adapt", 0 | 1 | 2): same Gram matrix bit for bit (0: four unconditional 32-entry batches per
    rater; 1: empty batches are skipped; 2 = default: also packed (relative index, value) entries), popular head and RED tail."""
    from rtrec_b200 import device as D
    U, I, N = 2500, 7500, 160000            # I > 4 * 1728: all four shared-memory ranges and the tail are populated
    u, i, ts, r = synth_events(U, I, N, seed=9, rating="cont")
    i = np.random.default_rng(3).permutation(I)[i]
    X = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))
    dX = D.DeviceMatrix.from_scipy(X)
    G0 = D.gram_full(dX).cpu().numpy()
    try:
        for mode in (0, 1):
            D.set_option("gram_adapt", mode)
            G1 = D.gram_full(dX).cpu().numpy()
            assert np.array_equal(G0, G1), mode
        D.set_option("gram_adapt", 2)
        # rank-sorted rows by the segmented sort (default: warp kernel for rows <= 64, CTA kernel above) vs the two global
        # radix sorts of the first version: the same rows, so the same matrix bit for bit
        D.set_option("gram_impl", 1)
        assert np.array_equal(G0, D.gram_full(dX).cpu().numpy())
    finally:
        D.set_option("gram_adapt", 2)
        D.set_option("gram_impl", 2)


def test_gram_row_sort_all_row_lengths():
    """Segmented row sort of the Gram preparation: rows of 1..64 entries (registers) and longer rows (counting sort over a
    shared-memory bitmap of ranks: one-warp CTAs up to 64k items, multi-warp CTAs above), against the radix path.  Integer
    ratings: every Gram entry is an exact integer sum, so the result must equal scipy's bit for bit whatever the summation
    order."""
    from rtrec_b200 import device as D
    rng = np.random.default_rng(5)
    for I, longest in ((10000, 3000), (10000, 9000), (70000, 20000)):
        U = 700
        lens = np.concatenate([rng.integers(1, 65, 300), rng.integers(65, 700, 390), [1, 64, 65, 128, 129, 2048, 2049, longest, 0, 0]])
        rows, cols = [], []
        for u, n in enumerate(lens):
            c = rng.choice(I, int(n), replace=False)
            rows.append(np.full(len(c), u)); cols.append(c)
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        X = sp.csc_matrix((rng.integers(1, 6, len(rows)).astype(np.float32), (rows, cols)), shape=(U, I))
        dX = D.DeviceMatrix.from_scipy(X)
        sel = np.unique(np.concatenate([rng.integers(0, I, 400), cols[-longest:][:50], [0, I - 1]]))
        ref = np.asarray((X[:, sel].T @ X).todense(), dtype=np.float32)
        for impl in (0, 1):
            D.set_option("gram_impl", impl)
            try:
                G = D.gram_full(dX)
                got = G[D.to_dev(sel, np.int64)][:, :I].cpu().numpy()
                del G
            finally:
                D.set_option("gram_impl", 0)
            assert np.array_equal(got, ref), (I, longest, impl)


@pytest.mark.parametrize("rating", ["int", "half", "cont", "big"])
def test_gram_head_on_tensor_cores_is_exact(rating):
    """gram_tc.cu: the 2,048 x 2,048 corner of the most popular items as a tcgen05 SYRK over a bf16 copy of X.  Taken only
    when the values are exact in bf16 and every partial sum is exact in fp32 (integer and half-integer ratings): the whole
    matrix is then bit-identical to the sparse kernel's and to scipy's.  Continuous ratings, or values whose sums could
    pass 2^24 units, keep the sparse kernel."""
    from rtrec_b200 import _lib, device as D
    U, I, N = 20000, 3000, 4_000_000
    u, i, ts, r = synth_events(U, I, N, seed=31, rating="int")
    X = sp.csc_matrix((np.ones(len(u), np.float32), (u, i)), shape=(U, I))
    X.sum_duplicates()
    rng = np.random.default_rng(7)
    if rating == "int":
        X.data[:] = rng.integers(1, 6, X.nnz)
    elif rating == "half":
        X.data[:] = rng.integers(-2, 11, X.nnz) * 0.5          # negative and explicit zero entries included
    elif rating == "big":
        X.data[:] = rng.integers(1, 200, X.nnz)                # exact in bf16, but the largest diagonal entry passes 2^24
    else:
        X.data[:] = rng.uniform(0.5, 5.0, X.nnz)
    dX = D.DeviceMatrix.from_scipy(X)
    lib = _lib.load()
    G1 = D.gram_full(dX)
    taken = int(lib.rt_gram_last_head())
    assert taken == (2048 if rating in ("int", "half") else 0), taken
    D.set_option("gram_head", 0)
    try:
        G0 = D.gram_full(dX)
        assert int(lib.rt_gram_last_head()) == 0
    finally:
        D.set_option("gram_head", 1)
    import torch
    if taken:
        assert torch.equal(G0, G1)
        sel = np.unique(np.concatenate([rng.integers(0, I, 200), np.argsort(-np.diff(X.indptr))[:100]]))
        ref = np.asarray((X[:, sel].T.astype(np.float64) @ X.astype(np.float64)).todense())
        got = G1[D.to_dev(sel, np.int64)][:, :I].cpu().numpy().astype(np.float64)
        assert np.array_equal(got, ref)
    else:
        assert float((G0 - G1).abs().max()) <= 2e-6 * float(G0.max())   # (float atomics: two runs differ in the last bits)


def test_predict_family_matches_scipy(golden):
    """predict / predict_selected / predict_all (slim_elastic.py:566-626) against scipy on the golden W: same float32
    sums (ascending source item), dense and sparse output forms, errors as in the reference."""
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    z = golden("slim_nn20_cont")
    X = csc_from(z, "X0").tocsr().astype(np.float32)
    W = w_from(z, "W0").astype(np.float32)
    m = SLIMElastic({"nn_feature_selection": 20})
    with pytest.raises(RuntimeError, match="Model must be fitted"):
        m.predict(0, X)
    m.item_similarity = W
    S = np.asarray((X @ W).todense(), dtype=np.float32)
    assert np.array_equal(m.predict_all(X), S)
    assert np.array_equal(m.predict(3, X), S[3:4])
    ids = [5, 0, 17, 5, 2]
    assert np.array_equal(m.predict_selected(3, ids, X), S[3:4][:, ids])
    sp_out = m.predict(3, X, dense_output=False)
    assert sp.issparse(sp_out) and np.array_equal(np.asarray(sp_out.todense()), S[3:4])
    with pytest.raises(IndexError):
        m.predict(X.shape[0], X)


@pytest.mark.parametrize("rating", ["half", "cont"])
def test_tensor_core_scoring_matches_exact_scores(rating):
    """score_tc.cu (tcgen05 + TMEM + TMA): heavy rows of W as a split-bf16 contraction, light rows and the merge on the CUDA
    cores.  Tolerance mode (north_star: top-k equal up to score ties within tolerance): (a) every heavy-only score the
    epilogue sees equals X_h . W_h to 5e-6 of the largest score (one bf16 plane for half-integer ratings, three for
    continuous ones), (b) every final list is a valid top-10 of the float64 scores within 2e-5, (c) almost all lists are
    identical to the exact kernel's."""
    import torch
    from rtrec_b200 import device as D
    from rtrec_b200._lib import RT_TOPK_DENSE, RT_TOPK_SPARSE
    rng = np.random.default_rng(23)
    n_items, n_users = 3000, 5000
    rows, cols, vals = [], [], []
    heavy = np.sort(rng.choice(n_items, 40, replace=False))
    for i in heavy:
        c = np.flatnonzero(rng.random(n_items) < 0.3)
        rows.append(np.full(len(c), i)); cols.append(c); vals.append((rng.random(len(c)) * 0.3).astype(np.float32))
    nb = 4 * n_items
    rows.append(rng.integers(0, n_items, nb)); cols.append(rng.integers(0, n_items, nb)); vals.append((rng.random(nb) * 0.2).astype(np.float32))
    W = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_items, n_items))
    W.sum_duplicates()
    xr, xc = [], []
    for u in range(n_users):
        its = np.unique(np.concatenate([rng.choice(n_items, int(rng.integers(1, 80)), replace=False),
                                        rng.choice(heavy, int(rng.integers(0, 12)), replace=False)]))
        xr.append(np.full(len(its), u)); xc.append(its)
    xr, xc = np.concatenate(xr), np.concatenate(xc)
    xv = (rng.integers(1, 11, len(xr)) * 0.5 if rating == "half" else rng.uniform(0.05, 5.0, len(xr))).astype(np.float32)
    X = sp.csr_matrix((xv, (xr, xc)), shape=(n_users, n_items))
    dX, dW = D.DeviceMatrix.from_scipy(X), D.DeviceW.from_scipy(W)
    users = torch.from_numpy(rng.permutation(n_users).astype(np.int32)).cuda()
    uh = users.cpu().numpy()
    old_min = D.PACK_MIN_ROW
    try:
        D.PACK_MIN_ROW = 256
        pk = D.tc_pack(dW)
        assert pk is not None and pk.n_heavy == 40 and pk.w_nonneg
        assert D.values_bf16_exact(dX) == (True, rating == "half")
        S64 = np.asarray((X.astype(np.float64) @ W.astype(np.float64)).todense())
        Wr = W.tocsr()
        hmask = np.zeros(n_items, bool); hmask[heavy] = True
        Xh = X.copy().tocsc(); Xh.data[~hmask[np.repeat(np.arange(n_items), np.diff(Xh.indptr))]] = 0
        Sh64 = np.asarray((Xh.tocsr().astype(np.float64) @ W.astype(np.float64)).todense())
        for mode in (RT_TOPK_SPARSE, RT_TOPK_DENSE):
            for filt in (True, False):
                ids, sc, cnt, dbg, tc, redo = D.recommend_tc(dX, users, dW, 10, filt, mode, debug_scores=True)
                dbg = dbg.cpu().numpy()[:, :n_items]
                scale = float(Sh64.max())
                err = np.abs(dbg - Sh64[uh]).max()
                assert err <= 5e-6 * scale, (rating, mode, filt, err, scale)
                ids, sc, cnt = ids.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()
                D.set_option("score_tc", 0)
                e_ids, e_sc, e_cnt = [x.cpu().numpy() for x in D.recommend(dX, users, dW, 10, filt, mode)]
                D.set_option("score_tc", 1)
                assert np.array_equal(cnt, e_cnt)
                same = 0
                for q in range(n_users):
                    u = int(uh[q])
                    inter = np.zeros(n_items, bool)
                    if filt:
                        inter[X[u].indices] = True
                    elig = ~inter if mode == RT_TOPK_DENSE else (~inter & (S64[u] != 0))
                    ok, why = topk_consistent(ids[q, :cnt[q]].tolist(), S64[u], 10, elig, tol=2e-5)
                    assert ok, (rating, mode, filt, q, why)
                    same += int(np.array_equal(ids[q], e_ids[q]))
                assert same >= 0.98 * n_users, (same, n_users)
                assert int(redo.numel()) < 0.2 * n_users
        # the public entry point takes this path for large batches by itself
        a = D.recommend(dX, users, dW, 10, True, RT_TOPK_SPARSE)
        assert a[0].shape == (n_users, 10)
        # queued form (no host wait for the hand-back list): same lists; too few slots are reported, and the list
        # builder then repeats the chunk the waiting way
        from rtrec_b200.models.internal.slim_elastic import SLIMElastic
        def same_lists(x, y):
            # (two runs of the tensor-core path agree up to the order of a few float atomics in the light-row table)
            assert len(x) == len(y)
            return sum(int(p == q) for p, q in zip(x, y)) >= 0.995 * len(x)

        def as_lists(r):
            return [row[:c].tolist() for row, c in zip(r[0].cpu().numpy(), r[2].cpu().numpy())]

        for mode in (RT_TOPK_SPARSE, RT_TOPK_DENSE):
            a = D.recommend(dX, users, dW, 10, True, mode)
            b = D.recommend(dX, users, dW, 10, True, mode, no_sync=True)
            n_back, slots = int(b[3][0]), b[3][1]
            if n_back > slots:      # more hand-backs than slots: the caller is told, and has to ask again the waiting way
                old_slots, D.TC_REDO_MIN_SLOTS = D.TC_REDO_MIN_SLOTS, n_back
                b = D.recommend(dX, users, dW, 10, True, mode, no_sync=True)
                D.TC_REDO_MIN_SLOTS = old_slots
                assert int(b[3][0]) == n_back and b[3][1] == n_back
            assert torch.equal(a[2], b[2]) and same_lists(as_lists(a), as_lists(b))
            valid = (torch.arange(10, device="cuda")[None, :] < a[2][:, None]).cpu().numpy()
            assert np.abs(a[1].cpu().numpy() - b[1].cpu().numpy())[valid].max() <= 1e-5 * float(a[1].cpu().numpy()[valid].max())
            op = SLIMElastic({})
            op._W = dW
            want = as_lists(a)
            assert same_lists(op.recommend_lists(uh, dX, 10, True, dense_output=mode == RT_TOPK_DENSE), want)
            if n_back > 1:
                old_slots, old_div, old_chunk = D.TC_REDO_MIN_SLOTS, D.TC_REDO_DIV, SLIMElastic._LIST_CHUNK
                D.TC_REDO_MIN_SLOTS, D.TC_REDO_DIV, SLIMElastic._LIST_CHUNK = 1, 1 << 30, 1 << 30
                try:
                    short = D.recommend(dX, users, dW, 10, True, mode, no_sync=True)[3]
                    assert int(short[0]) == n_back and short[1] == 1
                    assert same_lists(op.recommend_lists(uh, dX, 10, True, dense_output=mode == RT_TOPK_DENSE), want)
                finally:
                    D.TC_REDO_MIN_SLOTS, D.TC_REDO_DIV, SLIMElastic._LIST_CHUNK = old_slots, old_div, old_chunk
        # unused query slots (user id -1) of the exact kernels: empty answers, the other rows untouched
        D.set_option("score_tc", 0)
        try:
            for mode in (RT_TOPK_SPARSE, RT_TOPK_DENSE):
                for nq in (200, n_users):            # without / with the work-ordered queue
                    uu = users[:nq].clone()
                    uu[::3] = -1
                    r = D.recommend(dX, uu, dW, 10, True, mode)
                    f = D.recommend(dX, users[:nq], dW, 10, True, mode)
                    hole = (uu < 0)
                    assert int(r[2][hole].abs().sum()) == 0 and bool((r[0][hole] == -1).all())
                    assert torch.equal(r[0][~hole], f[0][~hole]) and torch.equal(r[1][~hole], f[1][~hole]) and torch.equal(r[2][~hole], f[2][~hole])
        finally:
            D.set_option("score_tc", 1)
    finally:
        D.PACK_MIN_ROW = old_min
        D.set_option("score_tc", 1)


def test_pruned_fit_equals_dense_path():
    """rt_slim_fit_pruned: all-features fit from the Gram rows of the Cauchy-Schwarz candidates only.  Same W as the dense
    path (Gram entries are accumulated by a different kernel: agreement to float32 rounding, same non-zero pattern), same
    stats for the zero columns, far fewer Gram rows; falls back (None) when the candidates are most of the catalogue."""
    import torch
    from rtrec_b200 import device as D
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    U, I, N = 60000, 2500, 300000
    u, i, ts, r = synth_events(U, I, N, seed=13, rating="cont")
    # small values (as after a time decay): a = 0.01 * 60000 = 600 is above most Gram entries, and the Cauchy-Schwarz
    # bound d_j * d_max > a^2 only holds for the few dozen most popular items
    X = sp.csc_matrix(((0.2 * r).astype(np.float32), (u, i)), shape=(U, I))     # 301 candidates, 7 non-trivial columns
    dX = D.DeviceMatrix.from_scipy(X)
    op = SLIMElastic({})
    cfg = op._config(dX)
    tg = torch.arange(I, dtype=torch.int32, device="cuda")
    res_p = D.fit_pruned(dX, tg, cfg)
    assert res_p is not None and 0 < D.last_pruned_rows <= I // 4, D.last_pruned_rows
    Wp = D.w_merge(None, I, res_p).to_scipy_csc()
    G = D.gram_full(dX)
    res_d = D.solve(G, I, tg, cfg)
    Wd = D.w_merge(None, I, res_d).to_scipy_csc()
    assert Wd.nnz > 0
    assert np.array_equal(Wp.indptr, Wd.indptr) and np.array_equal(Wp.indices, Wd.indices)
    assert np.abs(Wp.data - Wd.data).max() <= 1e-4 * np.abs(Wd.data).max()
    sp_, sd = res_p.stats.cpu().numpy(), res_d.stats.cpu().numpy()
    triv = sd[:, 3] == 0
    assert np.array_equal(sp_[triv], sd[triv])
    # every non-zero column lies inside the candidate set, and the operator takes this path by itself
    # (the selected Gram rows are accumulated with float atomics: two runs agree to float32 rounding, not bit for bit)
    op.fit(dX)
    Wo = sp.csc_matrix(op.item_similarity)
    assert np.array_equal(Wo.indptr, Wp.indptr) and np.array_equal(Wo.indices, Wp.indices)
    assert np.abs(Wo.data - Wp.data).max() <= 1e-5 * np.abs(Wp.data).max()
    # a small threshold (few users): most items are candidates -> the dense path is the better one
    u2, i2, ts2, r2 = synth_events(400, 300, 9000, seed=14, rating="cont")
    dX2 = D.DeviceMatrix.from_scipy(sp.csc_matrix((r2.astype(np.float32), (u2, i2)), shape=(400, 300)))
    assert D.fit_pruned(dX2, torch.arange(300, dtype=torch.int32, device="cuda"), op._config(dX2)) is None
    assert D.last_pruned_rows > 300 // 4
    # feature selection: a bulk fit takes the same path (columns without a live coordinate come back empty), a merge into
    # an existing W does not (a returned zero deletes a stale entry)
    opn = SLIMElastic({"nn_feature_selection": 20})
    assert D.fit_pruned(dX, tg, opn._config(dX, into_empty_w=False)) is None
    res_pn = D.fit_pruned(dX, tg, opn._config(dX))
    assert res_pn is not None
    Wpn = D.w_merge(None, I, res_pn).to_scipy_csc()
    Wdn = D.w_merge(None, I, D.solve(G, I, tg, opn._config(dX, into_empty_w=False))).to_scipy_csc()
    assert Wdn.nnz > 0 and np.array_equal(Wpn.indptr, Wdn.indptr) and np.array_equal(Wpn.indices, Wdn.indices)
    assert np.abs(Wpn.data - Wdn.data).max() <= 1e-4 * np.abs(Wdn.data).max()


def test_live_rows_only_gram_gives_the_same_w():
    """rt_gram_finish_live: for a bulk fit with feature selection the Gram rows of items without an entry above the L1
    threshold are never read by the solver (trivial targets are skipped, and by symmetry such an item is nobody's live
    coordinate), so only their diagonal is written.  The unwritten entries are NaN here: W must not notice, and must be
    bit-identical to the fit on the full matrix."""
    import torch
    from rtrec_b200 import device as D
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    U, I, N = 30000, 1500, 200000
    u, i, ts, r = synth_events(U, I, N, seed=17, rating="cont")
    dX = D.DeviceMatrix.from_scipy(sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I)))
    op = SLIMElastic({"nn_feature_selection": 20})
    cfg = op._config(dX)
    tg = torch.arange(I, dtype=torch.int32, device="cuda")
    L = D.gram_lower(dX)
    Gp_keep = L.Gp.clone()
    G_full = D.gram_finish(L)
    res_full = D.solve(G_full, I, tg, cfg)
    W_full = D.w_merge(None, I, res_full).to_scipy_csc()
    L.Gp.copy_(Gp_keep)                                     # (the finish pass mirrors Gp in place)
    out = torch.full((I, I), float("nan"), dtype=torch.float32, device="cuda")
    G_live = D.gram_finish(L, out=out, live_cfg=cfg)
    assert G_live._rt_live_only
    untouched = torch.isnan(G_live).any(dim=1)
    trivial = ~(G_live._rt_rowmax.double() > float(np.float32(0.1 * 0.1 * U)))
    assert int(untouched.sum()) > I // 10 and bool((untouched == trivial).all())
    assert torch.equal(torch.diagonal(G_live), torch.diagonal(G_full))
    assert torch.equal(G_live[~untouched], G_full[~untouched])
    for impl in (2, 1, 3):                                  # warp kernel, CTA kernel for every target, forced hand-overs
        D.set_option("solve_impl", impl)
        try:
            res = D.solve(G_live, I, tg, cfg)
        finally:
            D.set_option("solve_impl", 2)
        W = D.w_merge(None, I, res).to_scipy_csc()
        assert W.nnz == W_full.nnz and W.nnz > 0 and np.array_equal(W.indptr, W_full.indptr)
        assert np.array_equal(W.indices, W_full.indices) and np.array_equal(W.data, W_full.data), impl
        assert np.isfinite(W.data).all()
    # such a matrix cannot serve a merge fit or a call with candidate lists
    with pytest.raises(ValueError):
        D.solve(G_live, I, tg, op._config(dX, into_empty_w=False))
    with pytest.raises(ValueError):
        D.solve(G_live, I, tg, cfg, want_sel=True)
    # without feature selection, or for a merge, nothing is left out
    assert not D.gram_full(dX, live_cfg=SLIMElastic({})._config(dX))._rt_live_only
    assert not D.gram_full(dX, live_cfg=op._config(dX, into_empty_w=False))._rt_live_only
    # and the operator takes the path by itself (the pruned fit, which would come first, is switched off here: its Gram rows
    # are accumulated with float atomics and agree to rounding only)
    D.set_option("fit_pruned", 0)
    try:
        op.fit(dX)
    finally:
        D.set_option("fit_pruned", 1)
    Wo = sp.csc_matrix(op.item_similarity)
    # (a second Gram pass: the shared-memory float atomics of the sparse kernel add in another order, continuous ratings
    # differ in the last bits)
    assert np.array_equal(Wo.indptr, W_full.indptr) and np.array_equal(Wo.indices, W_full.indices)
    assert np.abs(Wo.data - W_full.data).max() <= 1e-5 * np.abs(W_full.data).max()


@pytest.mark.parametrize("nn", [20, None])
def test_trivial_columns_shortcut_gives_the_same_w(nn):
    """Targets whose Gram row has no entry above the L1 threshold are zero before the first sweep.  Bulk fits return them
    without pairs (nn mode: rt_fit_config.skip_trivial, decided after the first pass over the row; all features: the
    prefilter kernel); W, and the stats of every column, equal the full path's."""
    import torch
    from rtrec_b200 import device as D
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    U, I, N = 30000, 1200, 150000
    u, i, ts, r = synth_events(U, I, N, seed=8, rating="half")          # a = 0.01 * 30000 = 300: most columns are trivial
    X = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))
    dX = D.DeviceMatrix.from_scipy(X)
    G = D.gram_full(dX)
    tg = torch.arange(I, dtype=torch.int32, device="cuda")
    op = SLIMElastic({"nn_feature_selection": nn} if nn else {})
    outs = []
    try:
        for fast in (True, False):
            D.set_option("solve_impl", 2 if fast else 1)               # 1 = CTA kernel for everything, no shortcut
            res = D.solve(G, I, tg, op._config(dX, into_empty_w=fast))
            W = D.w_merge(None, I, res)
            outs.append((W.to_scipy_csc(), res.stats.cpu().numpy(), res.cnt.cpu().numpy()))
    finally:
        D.set_option("solve_impl", 2)
    (Wa, sa, ca), (Wb, sb, cb) = outs
    assert 0 < Wb.nnz and (np.diff(Wb.indptr) == 0).mean() > 0.3, "the case must mix trivial and non-trivial columns"
    assert np.array_equal(Wa.indptr, Wb.indptr) and np.array_equal(Wa.indices, Wb.indices)
    triv = sb[:, 3] == 0
    assert np.array_equal(sa[triv], sb[triv]) and (sb[triv] == np.array([0, 0, 1, 0])).all()
    if nn:
        assert (ca[triv] == 0).all() and (cb == nn).all()
        # warp kernel vs CTA kernel on the non-trivial columns: the same bar as test_warp_solver_equals_block_solver
        assert np.abs(Wa.data - Wb.data).max() <= 1e-3 * np.abs(Wb.data).max()
    else:
        assert np.array_equal(Wa.data, Wb.data)


@pytest.mark.parametrize("j_range", [None, (400, 2300)])
def test_sparse_table_and_zero_work_queries_equal_dense_tile(j_range):
    """Large batches are split by work (score3.cu): queries with no W entry behind their items are answered without a
    launch of their own, queries with <= 1,024 entries go to the sparse-table kernel (one warp per query), the rest to the
    dense-tile kernel.  Sparse top-k semantics, ids / scores / counts bit-identical to the v2 kernel, which knows none of
    this; dense semantics (zeros eligible) must not take the shortcut."""
    import torch
    from rtrec_b200 import device as D
    from rtrec_b200._lib import RT_TOPK_DENSE, RT_TOPK_SPARSE
    rng = np.random.default_rng(17)
    n_items, n_users = 3000, 6000
    # W: rows of items < 1000 are empty, 1000..2499 hold a few entries, 2500.. are long
    rows, cols, vals = [], [], []
    for i in range(1000, 2500):
        c = rng.choice(n_items, int(rng.integers(1, 9)), replace=False)
        rows.append(np.full(len(c), i)); cols.append(c); vals.append(rng.random(len(c)).astype(np.float32))
    for i in range(2500, n_items):
        c = rng.choice(n_items, int(rng.integers(200, 900)), replace=False)
        rows.append(np.full(len(c), i)); cols.append(c); vals.append(rng.random(len(c)).astype(np.float32))
    W = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_items, n_items))
    # users: a third only rate items with empty rows (no work), a third a few light items, a third also heavy items
    xr, xc = [], []
    for u in range(n_users):
        kind = u % 3
        its = rng.choice(1000, int(rng.integers(1, 30)), replace=False)
        if kind >= 1:
            its = np.concatenate([its, 1000 + rng.choice(1500, int(rng.integers(1, 60)), replace=False)])
        if kind == 2:
            its = np.concatenate([its, 2500 + rng.choice(500, int(rng.integers(2, 40)), replace=False)])
        xr.append(np.full(len(its), u)); xc.append(its)
    xr, xc = np.concatenate(xr), np.concatenate(xc)
    X = sp.csr_matrix((rng.integers(1, 11, len(xr)).astype(np.float32) * 0.5, (xr, xc)), shape=(n_users, n_items))
    dX, dW = D.DeviceMatrix.from_scipy(X), D.DeviceW.from_scipy(W)
    users = torch.from_numpy(rng.permutation(n_users).astype(np.int32)).cuda()
    j0, j1 = (0, n_items) if j_range is None else j_range
    try:
        for mode in (RT_TOPK_SPARSE, RT_TOPK_DENSE):
            for filt in (True, False):
                for k in (10, 40):      # k = 40 > 32: the sparse-table kernel is not used
                    D.set_option("score_impl", 2)
                    a = [x.cpu().numpy() for x in D.recommend(dX, users, dW, k, filt, mode, j0, j1)]
                    D.set_option("score_impl", 3)
                    b = [x.cpu().numpy() for x in D.recommend(dX, users, dW, k, filt, mode, j0, j1)]
                    for x, y in zip(a, b):
                        assert np.array_equal(x, y), (mode, filt, k)
                    if mode == RT_TOPK_SPARSE:
                        cnt = b[2][np.argsort(users.cpu().numpy())]
                        assert (cnt[0::3] == 0).all() and (cnt[2::3] > 0).all()
    finally:
        D.set_option("score_impl", 3)


# ------------------------------------------------------------------------------------------ fit
@pytest.mark.parametrize("name,cfg", FIT_CASES)
def test_fit_matches_reference_golden(golden, name, cfg):
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    z = golden(name)
    X0 = csc_from(z, "X0")
    op = SLIMElastic(dict(cfg, keep_fit_details=True))
    sel0 = z["sel0"] if "sel0" in z else None
    items0 = z["fit_items0"] if "fit_items0" in z else np.arange(X0.shape[1])
    op._fit_device(op._as_device(X0), items0, keep_old=False, sel_in=sel0)
    assert_w_parity(op.item_similarity, w_from(z, "W0"), what=name + " W0")
    if "W1_data" in z:
        X1 = csc_from(z, "X1")
        W_before = op.item_similarity.copy()
        op.partial_fit_items(X1, z["fit_items1"].tolist(), sel_in=z["sel1"] if "sel1" in z else None)
        assert_w_parity(op.item_similarity, w_from(z, "W1"), what=name + " W1")
        # stale entries: columns that were not re-solved are bit-identical to the old matrix
        W0, W1 = W_before, op.item_similarity
        untouched = np.setdiff1d(np.arange(W0.shape[1]), z["fit_items1"])
        for j in untouched[:50]:
            assert np.array_equal(W1[:, j].toarray()[:W0.shape[0]], W0[:, j].toarray())


@pytest.mark.parametrize("name", ["slim_nn20_int", "slim_nn20_cont", "slim_nn20_decay"])
def test_candidate_selection_rule(golden, name):
    """Built-in selection = score desc, ties -> larger item id; equals the reference's picks whenever
    no tie crosses the cut (continuous ratings), and is a valid top-n otherwise."""
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    z = golden(name)
    X0 = csc_from(z, "X0")
    op = SLIMElastic({"nn_feature_selection": 20, "keep_fit_details": True})
    op.fit(X0)
    sel = op.last_fit_sel
    _, sel_or, _ = so.fit_columns(X0, np.arange(X0.shape[1]), 20)
    assert np.array_equal(sel, sel_or), "device picks differ from the oracle's deterministic rule"
    if name != "slim_nn20_int":
        assert np.array_equal(sel, z["sel0"]), "tie-free data: picks must equal the reference's"
    o = so.SlimOracle({"nn_feature_selection": 20})
    o.fit(X0, sel_in=sel)
    assert_w_parity(op.item_similarity, o.item_similarity, what=name)


@pytest.mark.parametrize("rating,nn,ncols", [("int", 50, 400), ("cont", 50, 400), ("int", None, 60)])
def test_fit_ml1m_shape_sampled_columns(rating, nn, ncols):
    """ML-1M shape (BASELINE configs[0]); oracle on a column sample via partial_fit_items semantics."""
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    U, I, N = 6040, 3706, 1_000_000
    u, i, ts, r = synth_events(U, I, N, seed=0, rating=rating)
    X = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))
    rng = np.random.default_rng(1)
    tg = np.sort(rng.choice(I, ncols, replace=False)).astype(np.int32)
    op = SLIMElastic({"nn_feature_selection": nn, "keep_fit_details": True})
    op.partial_fit_items(X, tg.tolist())
    sel = op.last_fit_sel
    o = so.SlimOracle({"nn_feature_selection": nn, "n_threads": 8})
    o.partial_fit_items(X, tg, sel_in=sel)
    assert_w_parity(op.item_similarity, o.item_similarity, cols=tg, what=f"ml1m {rating} nn={nn}")


# ------------------------------------------------------------------------------------------ scoring
@pytest.mark.parametrize("name", ["slim_all_int", "slim_nn20_cont", "slim_nn20_decay"])
@pytest.mark.parametrize("dense", [True, False])
def test_recommend_topk(golden, name, dense):
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    z = golden(name)
    S = z["scores_dense"]
    W = w_from(z, "W1") if "W1_data" in z else w_from(z, "W0")
    n_items = W.shape[0]
    # rebuild the final X (all events) with the store oracle
    ev = z["events"]
    decay = 180 if "decay" in name else None
    st = so.fold_events(ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64), ev[:, 2], ev[:, 3], decay_in_days=decay)
    X = so.state_to_matrix(st, decay_in_days=decay, fmt="csr")
    assert X.shape[1] == n_items
    op = SLIMElastic({})
    op.item_similarity = W
    users = z["rec_users"]
    for filt in (True, False):
        res = op.recommend_batch(users.tolist(), X, top_k=10, filter_interacted=filt, dense_output=dense, ret_scores=False)
        for r, uid in enumerate(users):
            row = X[uid]
            elig = np.ones(n_items, bool)
            if filt:
                elig[row.indices] = False
            if not dense:
                elig &= S[uid] != 0
            ok, why = topk_consistent(res[r], S[uid], 10, elig)
            assert ok, (name, dense, filt, int(uid), why)
    # the reference's own lists (filter on): same items wherever scores are separated
    if z["pass_through"] == (not dense):
        res = op.recommend_batch(users.tolist(), X, top_k=10, filter_interacted=True, dense_output=dense)
        same = sum(1 for r in range(len(users)) if res[r] == [x for x in z["rec_top10"][r] if x >= 0])
        assert same >= 0.9 * len(users), (same, len(users))


def test_recommend_candidates_and_errors(golden):
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    z = golden("slim_nn20_cont")
    W = w_from(z, "W0")
    ev = z["events"]
    X = so.state_to_matrix(so.fold_events(ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64), ev[:, 2], ev[:, 3]), fmt="csr")
    op = SLIMElastic({})
    with pytest.raises(RuntimeError):
        op.recommend_batch([0], X)
    with pytest.raises(RuntimeError):
        op.similar_items(0)
    op.item_similarity = W
    cand = [5, 3, 40, 41, 7, 100, 2]
    o = so.SlimOracle({})
    o.item_similarity = W
    users = z["rec_users"][:64].tolist()
    got = op.recommend_batch(users, X, candidate_item_ids=cand, top_k=3, ret_scores=True)
    exp = o.recommend_batch(users, X, candidate_item_ids=cand, top_k=3, ret_scores=True)
    for (gi, gs), (ei, es) in zip(got, exp):
        assert np.allclose(gs, es, rtol=1e-5, atol=1e-6)
        assert gi == ei or np.isclose(np.sort(gs), np.sort(es)).all()
    with pytest.raises(ValueError):
        op.fit(np.zeros((3, 3)))
    assert op.recommend_batch([], X) == []


def test_similar_items_matches_reference(golden):
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    for name in ("slim_all_int", "slim_nn20_cont"):
        z = golden(name)
        W = w_from(z, "W0")
        op = SLIMElastic({})
        op.item_similarity = W
        ids, scores, cnt = op.similar_items_batch(np.arange(W.shape[0]), 10)
        for j in range(W.shape[0]):
            c = int(cnt[j])
            ref_ids = [x for x in z["sim_ids"][j] if x >= 0]
            assert c == len(ref_ids)
            assert np.array_equal(scores[j, :c], z["sim_scores"][j, :c])   # same score sequence
            if len(set(z["sim_scores"][j, :c].tolist())) == c:
                assert ids[j, :c].tolist() == ref_ids


def test_shard_merge_equals_single_pass(golden):
    from rtrec_b200 import device as D
    from rtrec_b200._lib import RT_TOPK_DENSE, RT_TOPK_SPARSE
    z = golden("slim_nn20_cont")
    W = D.DeviceW.from_scipy(w_from(z, "W0"))
    ev = z["events"]
    X = D.DeviceMatrix.from_scipy(so.state_to_matrix(
        so.fold_events(ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64), ev[:, 2], ev[:, 3]), fmt="csr"))
    t = D.torch()
    users = D.to_dev(z["rec_users"].astype(np.int32))
    Q, k, I = users.numel(), 10, W.n_items
    for mode in (RT_TOPK_DENSE, RT_TOPK_SPARSE):
        full_ids, full_sc, full_cnt = D.recommend(X, users, W, k, True, mode)
        cuts = [0, 37, 90, I]
        parts = [D.recommend(X, users, W, k, True, mode, cuts[s], cuts[s + 1]) for s in range(3)]
        ids = t.stack([p[0] for p in parts]).contiguous()
        sc = t.stack([p[1] for p in parts]).contiguous()
        m_ids, m_sc, m_cnt = D.topk_merge(ids, sc, 3, Q, k)
        assert t.equal(m_cnt, full_cnt)
        assert t.equal(m_ids, full_ids)
        assert t.equal(m_sc, full_sc)


# ------------------------------------------------------------------------------------------ packed scoring (score3.cu)
@pytest.mark.parametrize("n_items,j_range", [(3000, None), (3000, (512, 2500)), (61000, None), (61000, (1000, 60001))])
def test_packed_scoring_equals_csr_scoring(n_items, j_range):
    """The bank-striped ELL pack changes the layout of W's heavy rows, not the arithmetic: ids, scores and
    counts equal the v2 kernel bit for bit (single tile, item sub-range, several tiles), and the scores equal
    a float32 numpy accumulation in ascending item order."""
    import torch
    from rtrec_b200 import device as D
    from rtrec_b200._lib import RT_TOPK_DENSE, RT_TOPK_SPARSE
    rng = np.random.default_rng(n_items)
    n_users = 700
    # W: 12 heavy source rows (a quarter of the columns each) + a sparse background
    rows, cols, vals = [], [], []
    heavy = rng.choice(n_items, 12, replace=False)
    for i in heavy:
        c = np.flatnonzero(rng.random(n_items) < 0.25)
        rows.append(np.full(len(c), i)); cols.append(c); vals.append(rng.random(len(c)).astype(np.float32))
    nb = 5 * n_items
    rows.append(rng.integers(0, n_items, nb)); cols.append(rng.integers(0, n_items, nb)); vals.append(rng.random(nb).astype(np.float32))
    W = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_items, n_items))
    W.sum_duplicates()
    # X: every user rates ~40 random items plus a few of the heavy ones
    xr = np.repeat(np.arange(n_users), 44)
    xc = np.concatenate([np.concatenate([rng.choice(n_items, 40, replace=False), rng.choice(heavy, 4, replace=False)])
                         for _ in range(n_users)])
    X = sp.csr_matrix((rng.integers(1, 6, len(xr)).astype(np.float32), (xr, xc)), shape=(n_users, n_items))
    X.sum_duplicates()
    dX, dW = D.DeviceMatrix.from_scipy(X), D.DeviceW.from_scipy(W)
    users = torch.arange(n_users, dtype=torch.int32, device="cuda")
    j0, j1 = (0, n_items) if j_range is None else j_range
    old_min = D.PACK_MIN_ROW
    try:
        D.PACK_MIN_ROW = 64
        for mode in (RT_TOPK_SPARSE, RT_TOPK_DENSE):
            for filt in (True, False):
                D.set_option("score_impl", 2)
                a = [x.cpu().numpy() for x in D.recommend(dX, users, dW, 10, filt, mode, j0, j1)]
                D.set_option("score_impl", 3)
                b = [x.cpu().numpy() for x in D.recommend(dX, users, dW, 10, filt, mode, j0, j1)]
                for x, y in zip(a, b):
                    assert np.array_equal(x, y)
        pk = D.score_pack(dW, j0, j1)
        assert pk.n_heavy >= 12 and pk.n_groups > 0
    finally:
        D.PACK_MIN_ROW = old_min
        D.set_option("score_impl", 3)
    # exact scores of a few users: ascending-item float32 accumulation (scipy csr_matmat order)
    Wr = W.tocsr(); Wr.sort_indices()
    ids, scores, cnt = b
    for u in range(0, n_users, 97):
        s = np.zeros(n_items, dtype=np.float32)
        for p in range(X.indptr[u], X.indptr[u + 1]):
            i, x = X.indices[p], X.data[p]
            sl = slice(Wr.indptr[i], Wr.indptr[i + 1])
            s[Wr.indices[sl]] = s[Wr.indices[sl]] + np.float32(x) * Wr.data[sl]
        for e in range(cnt[u]):
            assert scores[u, e] == s[ids[u, e]]


# ------------------------------------------------------------------------------------------ warp-per-column solver
@pytest.mark.parametrize("rating", ["int", "cont"])
def test_warp_solver_equals_block_solver(rating):
    """solve_impl 2 (one warp per target column, nn <= 64) against solve_impl 1 (one CTA per column): the same
    candidates in the same order for every column -- including empty columns (all scores tie at 0) and
    integer-rating ties across the cut -- the same sweep counts and coefficients (the reductions inside the
    duality gap are summed in a different order, so a last-bit difference is tolerated on a handful of
    columns), and the overflow hand-over to the CTA kernel (solve_impl 3 flags every 7th column)."""
    import torch
    from rtrec_b200 import device as D
    from rtrec_b200.models.internal.slim_elastic import SLIMElastic
    U, I, N = 5000, 1500, 200000
    u, i, ts, r = synth_events(U, I, N, seed=12, rating=rating)
    keep = (i % 97) != 5               # a few items without any interaction: all-zero Gram rows
    X = sp.csc_matrix((r[keep].astype(np.float32), (u[keep], i[keep])), shape=(U, I))
    dX = D.DeviceMatrix.from_scipy(X)
    G = D.gram_full(dX)
    tg = torch.arange(I, dtype=torch.int32, device="cuda")
    out = {}
    try:
        for nn in (50, 64, 7):
            cfg = SLIMElastic({"nn_feature_selection": nn})._config(dX)
            for impl in (1, 2, 3):
                D.set_option("solve_impl", impl)
                res = D.solve(G, I, tg, cfg, want_sel=True)
                out[impl] = [x.cpu().numpy() for x in (res.rows, res.vals, res.stats, res.sel, res.off, res.cnt)]
            for impl in (2, 3):
                a, b = out[1], out[impl]
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3]), f"candidates differ (nn={nn}, impl={impl})"
                assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])
                va, vb = a[1].reshape(I, nn), b[1].reshape(I, nn)
                differ = (va != vb).any(axis=1)
                assert differ.sum() <= max(2, I // 200), int(differ.sum())
                scale = np.maximum(np.abs(va).max(axis=1), 1e-30)
                assert (np.abs(va - vb).max(axis=1) / scale)[differ].max(initial=0.0) <= 1e-3
                assert (a[2][:, 0] != b[2][:, 0]).sum() <= max(2, I // 200)
    finally:
        D.set_option("solve_impl", 2)
