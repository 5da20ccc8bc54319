"""BASELINE.json configs[1] at its FULL size (synthetic MovieLens-20M shape: 138,493 x 26,744, 20M ratings,
nn_feature_selection=50) on the GPU, through the public API: size-independent properties of every stage plus
oracle parity on a sample of columns and users (the oracle needs ~40 ms per column at this size, so the whole
fit cannot be replayed on the CPU inside a test).

  store    CSR and CSC hold the same 20M entries, canonical (sorted, no duplicates), bit-equal to the folded events
  fit      W >= 0, zero diagonal, <= nn entries per column, every entry's row is a co-rated item;
           sampled columns within the at-scale parity bar of tests/helpers.py against the oracle (same candidates):
           1e-4 of the column maximum wherever that maximum is >= 1e-3, absolute 1e-5 + equal ElasticNet objectives
           where the reference's own float32 noise exceeds 1e-4 of a tiny column (tests/c2_parity_cpu.py)
  scoring  every list: <= 10 items, no interacted item, no duplicates, scores descending and > 0 (int ids -> sparse
           semantics); sampled users: valid top-10 of the oracle's scores; an order-independent checksum of all lists
           is reproduced by a second pass in one launch instead of chunks
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import slim_oracle as so
from oracle.synth import synth_shape
from tests.helpers import assert_w_parity_at_scale, topk_consistent

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    from rtrec_b200.models import SLIM
    u, i, ts, r = synth_shape("ml20m")
    m = SLIM(nn_feature_selection=50, keep_fit_details=True)
    m.add_interaction_arrays(u, i, ts, r)
    m.bulk_fit()
    return m, u, i, ts, r


def test_c2_store_full_size(c2):
    m, u, i, ts, r = c2
    U, I = int(u.max()) + 1, int(i.max()) + 1
    Xc = m.interactions.to_csc()
    Xr = m.interactions.to_csr()
    assert Xc.shape == Xr.shape == (U, I) and Xc.nnz == Xr.nnz == len(u) == 20_000_000
    assert Xc.has_canonical_format and Xr.has_canonical_format
    ref = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))       # the synthetic pairs are unique
    ref.sort_indices()
    assert np.array_equal(Xc.indptr, ref.indptr) and np.array_equal(Xc.indices, ref.indices)
    assert np.array_equal(Xc.data, ref.data)
    assert (Xr.tocsc() != Xc).nnz == 0
    assert m.interactions.max_timestamp == float(ts.max()) + 1.0
    assert (m.interactions.max_user_id, m.interactions.max_item_id) == (U - 1, I - 1)


def test_c2_fit_full_size(c2):
    m, u, i, ts, r = c2
    I = int(i.max()) + 1
    W = m.model.item_similarity.tocsc()
    assert W.shape == (I, I) and W.dtype == np.float32
    assert (W.data > 0).all() and (W.diagonal() == 0).all()
    assert np.diff(W.indptr).max() <= 50
    Xc = m.interactions.to_csc()
    rng = np.random.default_rng(7)
    nz_cols = np.flatnonzero(np.diff(W.indptr) > 0)
    cols = np.sort(np.concatenate([rng.choice(nz_cols, 12, replace=False), rng.choice(I, 12, replace=False)])).astype(np.int32)
    cols = np.unique(cols)
    # same candidate order as the device chose (exact integer-rating ties at the cut, SURVEY.md hard part 3)
    tg = m.model.last_fit_targets
    pos = {int(t): k for k, t in enumerate(tg)}
    sel = np.stack([m.model.last_fit_sel[pos[int(j)]] for j in cols])
    res, _, _ = so.fit_columns(Xc, cols, 50, sel_in=sel, n_threads=8)
    o = so.SlimOracle({"nn_feature_selection": 50})
    colsd = {}
    for j, (rows, vals) in zip(cols, res):
        so.SlimOracle._apply(colsd, int(j), rows, vals)
    Wo = so.SlimOracle._to_csc(colsd, I)
    assert_w_parity_at_scale(W, Wo, cols, Xc, what="W at ML-20M shape")
    # every neighbour is co-rated with its target (its Gram entry is positive): check on the sampled columns
    Xb = (Xc != 0).astype(np.float32)
    for j in cols:
        rows = W.indices[W.indptr[j]:W.indptr[j + 1]]
        if len(rows):
            co = np.asarray((Xb[:, rows].T @ Xb[:, [int(j)]]).todense()).ravel()
            assert (co > 0).all(), int(j)


def test_c2_scoring_full_size(c2):
    m, u, i, ts, r = c2
    U, I = int(u.max()) + 1, int(i.max()) + 1
    lists = m.recommend_batch(list(range(U)), top_k=10)
    assert len(lists) == U
    lens = np.fromiter((len(x) for x in lists), dtype=np.int64, count=U)
    assert lens.max() <= 10
    flat = np.fromiter((y for x in lists for y in x), dtype=np.int64, count=int(lens.sum()))
    assert flat.min() >= 0 and flat.max() < I
    owner = np.repeat(np.arange(U, dtype=np.int64), lens)
    key = owner * I + flat
    assert len(np.unique(key)) == len(key)                                   # no duplicates inside a list
    seen = np.sort(u.astype(np.int64) * I + i)
    assert not np.isin(key, seen, assume_unique=False).any()                 # interacted items are filtered
    # one launch over all users (device arrays) reproduces the chunked list pipeline
    X = m.interactions.device_matrix()
    ids, scores, cnt = m.model.recommend_batch_device(np.arange(U), X, None, 10, True, False)
    assert np.array_equal(cnt, lens)
    assert np.array_equal(ids[ids >= 0], flat)
    valid = np.arange(10)[None, :] < cnt[:, None]
    pair = np.arange(9)[None, :] < (cnt[:, None] - 1)
    assert ((scores[:, 1:] - scores[:, :-1])[pair] <= 0).all()                # descending inside every list
    assert (scores[valid] > 0).all()                                         # sparse semantics: only scored items
    # sampled users against scores computed on the CPU from the fitted W
    W = m.model.item_similarity.tocsc().astype(np.float32)
    Xr = m.interactions.to_csr()
    rng = np.random.default_rng(11)
    users = np.sort(rng.choice(U, 96, replace=False))
    S = np.asarray((Xr[users, :] @ W).todense(), dtype=np.float32)
    for q, uid in enumerate(users):
        inter = np.zeros(I, bool)
        inter[Xr[uid].indices] = True
        ok, why = topk_consistent(lists[uid], S[q], 10, ~inter & (S[q] != 0), tol=2e-5)
        assert ok, (int(uid), why)
