"""CPU tests that PIN the oracle (oracle/) before anything is compared against it.

1. slim_oracle.c vs the installed scikit-learn ElasticNet (the third-party code the reference calls,
   slim_elastic.py:197-208): coefficients bit-for-bit, same n_iter.
2. oracle vs golden vectors produced by importing the real reference (tests/golden/make_golden.py):
   W after bulk fit, W after a streaming partial fit (stale-entry merge), store matrices.
3. gram_model.c (the CPU model of the device algorithm) vs the exact port: the Gram-form replay
   stays within the north_star tolerance.
"""
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import slim_oracle as so
from rtrec_b200.utils.synth import synth_events
from tests.helpers import assert_w_parity, csc_from, w_from

CASES = [
    ("slim_all_int", {}),
    ("slim_nn20_int", {"nn_feature_selection": 20}),
    ("slim_nn20_cont", {"nn_feature_selection": 20}),
    ("slim_nn20_decay", {"nn_feature_selection": 20}),
    ("slim_all_decay_partial", {}),
    ("slim_nn20_partial", {"nn_feature_selection": 20}),
    ("slim_all_strids_fit", {}),
]


def _same(A, B):
    A = sp.csc_matrix(A); B = sp.csc_matrix(B)
    A.sort_indices(); B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data.astype(np.float32), B.data.astype(np.float32)))


def test_c_port_matches_sklearn_bit_for_bit():
    from sklearn.exceptions import ConvergenceWarning
    from sklearn.linear_model import ElasticNet
    rng = np.random.default_rng(0)
    U, I = 500, 120
    X = sp.random(U, I, density=0.08, random_state=1, format="csc", dtype=np.float32)
    X.data = np.ceil(X.data * 5).astype(np.float32)
    n = same = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", ConvergenceWarning)
        for mode in ("int", "dec"):
            Xm = X.copy()
            if mode == "dec":
                Xm.data = (Xm.data * 0.99 ** rng.uniform(0, 300, Xm.nnz)).astype(np.float32)
            for j in range(0, 40):
                y = np.asarray(Xm[:, j].todense()).ravel().astype(np.float32)
                Xz = Xm.copy()
                Xz.data[Xz.indptr[j]:Xz.indptr[j + 1]] = 0
                for nf in (15, None):
                    Xs = Xz
                    if nf:
                        s = Xz.T.dot(y)
                        Xs = Xz[:, np.argsort(s)[-1:-1 - nf:-1]]
                    m = ElasticNet(alpha=0.1, l1_ratio=0.1, fit_intercept=False, precompute=True, max_iter=100,
                                   copy_X=False, tol=1e-4, positive=True, random_state=43, selection="random")
                    m.fit(Xs, y)
                    w, it, _ = so.enet_solve(Xs, y)
                    n += 1
                    same += int(np.array_equal(w, m.coef_) and it == m.n_iter_)
    # the only non-replayable part is the BLAS reduction order inside the duality gap; allow 1 %
    assert same >= 0.99 * n, (same, n)


@pytest.mark.parametrize("name,cfg", CASES)
def test_oracle_reproduces_reference_w(golden, name, cfg):
    z = golden(name)
    X0 = csc_from(z, "X0")
    o = so.SlimOracle(cfg)
    sel0 = z["sel0"] if "sel0" in z else None
    if "fit_items0" in z:
        o.partial_fit_items(X0, z["fit_items0"], sel_in=sel0)
    else:
        o.fit(X0, sel_in=sel0)
    assert _same(o.item_similarity, w_from(z, "W0")), "bulk W differs from the reference"
    if "W1_data" in z:
        X1 = csc_from(z, "X1")
        o.partial_fit_items(X1, z["fit_items1"], sel_in=z["sel1"] if "sel1" in z else None)
        assert _same(o.item_similarity, w_from(z, "W1")), "W after partial fit differs from the reference"


@pytest.mark.parametrize("k", range(4))
def test_store_oracle_reproduces_reference(golden, k):
    z = golden(f"store_{k}")
    ev = z["events"]
    decay = None if z["decay"] < 0 else int(z["decay"])
    ups = bool(z["upsert"])
    u, i = ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64)
    st = None
    for a, b in ((0, 1500), (1500, 1501), (1501, len(u))):
        st = so.fold_events(u[a:b], i[a:b], ev[a:b, 2], ev[a:b, 3], upsert=ups, decay_in_days=decay, state=st)
    X = so.state_to_matrix(st, decay_in_days=decay)
    X.sort_indices()
    assert np.array_equal(X.indptr, z["X_indptr"]) and np.array_equal(X.indices, z["X_indices"])
    assert np.array_equal(X.data, z["X_data"])
    assert st[3] == float(z["max_timestamp"]) and st[4] == int(z["max_user_id"]) and st[5] == int(z["max_item_id"])
    s2 = so.StoreOracle(decay_in_days=decay)
    for a, b, c, d in ev:
        s2.add(int(a), int(b), float(c), float(d), upsert=ups)
    X2 = s2.to_csc()
    X2.sort_indices()
    assert np.array_equal(X2.data, z["X_data"]) and np.array_equal(X2.indices, z["X_indices"])


@pytest.mark.parametrize("rating,nn", [("int", 50), ("cont", 50), ("int", None), ("cont", None)])
def test_gram_form_model_matches_exact_port(rating, nn):
    U, I, N = 1500, 400, 70000
    u, i, ts, r = synth_events(U, I, N, seed=7, rating=rating)
    X = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))
    tg = np.arange(0, I, 3, dtype=np.int32)
    cols, sel, st = so.fit_columns(X, tg, nn, n_threads=4)
    gcols, gsel, gst = so.gram_model_fit_columns(X, tg, nn, sel_in=sel)
    worst, flips = 0.0, 0
    for t in range(len(tg)):
        a = np.zeros(I, np.float64); b = np.zeros(I, np.float64)
        a[cols[t][0]] = cols[t][1]; b[gcols[t][0]] = gcols[t][1]
        scale = max(np.abs(a).max(), 1e-30)
        e = np.abs(a - b).max() / scale
        worst = max(worst, e)
        flips += int(e > 1e-4)
    assert worst <= 1e-3, worst
    assert flips <= max(1, len(tg) // 50), (flips, len(tg))


def test_gram_form_model_all_columns_ml1m_shape():
    """BASELINE configs[0] at full size on the CPU: the model of the device algorithm (Gram form, float64 solver state,
    oracle/gram_model.c) against the exact port of the reference path (float32 residual form) on ALL 3,706 columns of the
    synthetic ML-1M-shaped matrix, same candidates -- the parity bar of tests/helpers.py (1e-4 of the column maximum, at
    most 2 % one-sweep flips agreeing to 1e-3 or in objective).  Observed: 7 columns above 1e-4, worst 5.4e-4."""
    from rtrec_b200.utils.synth import synth_shape
    from tests.helpers import assert_w_parity_at_scale
    u, i, ts, r = synth_shape("ml1m")
    U, I = int(u.max()) + 1, int(i.max()) + 1
    X = sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I))
    tg = np.arange(I, dtype=np.int32)
    cols, sel, st = so.fit_columns(X, tg, 50, n_threads=4)
    gcols, _, gst = so.gram_model_fit_columns(X, tg, 50, sel_in=sel)

    def assemble(res):
        rr, cc, vv = [], [], []
        for j, (rows, vals) in zip(tg, res):
            nz = vals != 0
            rr += rows[nz].tolist(); cc += [int(j)] * int(nz.sum()); vv += vals[nz].tolist()
        return sp.csc_matrix((np.array(vv, dtype=np.float32), (rr, cc)), shape=(I, I))

    rel, report = assert_w_parity_at_scale(assemble(gcols), assemble(cols), tg, X, what="Gram-form model at ML-1M shape")
    assert (rel > 1e-4).sum() <= 0.005 * I
    assert (st[:, 0] != gst[:, 0]).sum() <= 0.02 * I          # sweep counts differ on well under 2 % of the columns


def test_store_oracle_reference_known_answers():
    """The known answers of the reference's own store tests (/root/reference/tests/utils/test_interactions.py:21-35,
    65-116, 118-141), replayed on the oracle -- both on the event-at-a-time restatement and on the vectorised fold the
    GPU parity tests compare against.  Wall-clock sleeps are replaced by explicit timestamps: the reference decays
    relative to ``max_timestamp`` (= newest event + 1 s, interactions.py:99), not to the clock."""
    now = 1.7e9
    s = so.StoreOracle(min_value=-5, max_value=10)
    s.add(1, 10, now, 5.0)
    assert s.rating(1, 10) == 5.0                                     # :21-24
    s.add(1, 10, now, 3.0)
    assert s.rating(1, 10) == 8.0                                     # :26-30 (accumulate)
    s.add(1, 10, now, -8.0)
    assert s.rating(1, 10) == 0.0 and (1, 10) in s.pairs              # :32-37 (a zero stays stored)
    assert s.rating(99, 10) == 0.0                                    # :65-67 (unknown user)
    s.add(1, 10, now, 50.0)
    assert s.rating(1, 10) == 10.0                                    # clip at max_value (interactions.py:105-108)
    # decay over 7 days with decay_in_days=7: another user's event advances max_timestamp (:92-116)
    d = so.StoreOracle(min_value=-5, max_value=10, decay_in_days=7)
    d.add(1, 10, now - 7 * 86400, 5.0)
    d.add(2, 10, now, 5.0)
    factor = d.decay_rate ** (7 + 1 / 86400.0)
    assert abs(0.5 - factor) < 0.02
    assert abs(d.rating(1, 10) - 5.0 * factor) < 1e-12 and abs(2.5 - d.rating(1, 10)) < 0.1
    # to_csc, full and with select_items (:118-141)
    ev = [(0, 0, 12345, 5), (0, 1, 12346, 3), (1, 1, 12347, 4), (1, 2, 12348, 2), (2, 0, 12349, 1), (2, 2, 12350, 3)]
    t = so.StoreOracle()
    for u, i, ts, v in ev:
        t.add(u, i, float(ts), float(v))
    full = sp.csc_matrix(([5, 3, 4, 2, 1, 3], ([0, 0, 1, 1, 2, 2], [0, 1, 1, 2, 0, 2])), shape=(3, 3))
    part = sp.csc_matrix(([3, 4, 2, 3], ([0, 1, 1, 2], [1, 1, 2, 2])), shape=(3, 3))
    assert (t.to_csc() != full).nnz == 0 and (t.to_csc(select_items=[1, 2]) != part).nnz == 0
    a = np.array(ev, dtype=np.float64)
    st = so.fold_events(a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), a[:, 2], a[:, 3])
    assert (so.state_to_matrix(st, fmt="csc") != full).nnz == 0
    assert (so.state_to_matrix(st, fmt="csc", select_items=[1, 2]) != part).nnz == 0
    # the same decay case through the vectorised fold
    st = so.fold_events(np.array([1, 2]), np.array([10, 10]), np.array([now - 7 * 86400, now]), np.array([5.0, 5.0]),
                        decay_in_days=7)
    X = so.state_to_matrix(st, decay_in_days=7, fmt="csc")
    assert abs(float(X[1, 10]) - np.float32(5.0 * factor)) < 1e-6


def _toy_ids(events):
    users, items = {}, {}
    out = []
    for u, i, ts, r in events:
        out.append((users.setdefault(u, len(users)), items.setdefault(i, len(items)), ts, r))
    return out, users, items


def test_slim_oracle_reference_known_answers():
    """The two orderings the reference's own SLIM tests pin (/root/reference/tests/models/test_slim.py:58-79, 81-98; the
    README example :49-73 is the second one), replayed on the oracle through the same flow as slim.py:28-43 (ingest ->
    to_csc(item_ids) -> partial_fit_items -> similar_items / dense top-k for string ids)."""
    now = 1.7e9
    ev, users, items = _toy_ids([("user_1", "item_1", now, 5.0), ("user_1", "item_3", now, 4.0), ("user_1", "item_4", now, 3.0),
                                 ("user_2", "item_1", now, 3.0), ("user_2", "item_2", now, -2.0), ("user_2", "item_4", now, 3.0),
                                 ("user_3", "item_1", now, 4.0), ("user_3", "item_3", now, 2.0), ("user_3", "item_4", now, 4.0)])
    name = {v: k for k, v in items.items()}
    s = so.StoreOracle()
    for e in ev:
        s.add(*e)
    o = so.SlimOracle({})
    o.partial_fit_items(s.to_csc(select_items=sorted(set(items.values()))), sorted(set(items.values())))
    sims = o.similar_items(items["item_1"], top_k=5)
    assert [name[i] for i, _ in sims] == ["item_4", "item_3"] and sims[0][1] > sims[1][1]
    # fit twice (the second pass re-adds the same events: values accumulate and clip), then recommend for user_1
    ev, users, items = _toy_ids([("user_1", "item_1", now, 5.0), ("user_2", "item_2", now, -2.0), ("user_2", "item_1", now, 3.0),
                                 ("user_2", "item_4", now, 3.0), ("user_1", "item_3", now, 4.0)])
    name = {v: k for k, v in items.items()}
    s = so.StoreOracle()
    o = so.SlimOracle({})
    for _ in range(2):
        for e in ev:
            s.add(*e)
        o.partial_fit_items(s.to_csc(select_items=sorted(set(items.values()))), sorted(set(items.values())))
    assert s.rating(users["user_1"], items["item_1"]) == 10.0 and s.rating(users["user_2"], items["item_2"]) == -4.0
    rec = o.recommend_batch([users["user_1"]], s.to_csr(), top_k=5, dense_output=True)[0]
    assert [name[i] for i in rec] == ["item_4", "item_2"]
