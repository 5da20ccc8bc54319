"""Generate golden vectors by running the REAL reference (myui/rtrec at /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports ``rtrec`` from /root/reference (with two stub packages for the absent ``implicit``
and ``lightfm`` wheels that ``rtrec/models/__init__.py`` imports), drives the serial SLIM path on
small seeded inputs and stores inputs + outputs as ``tests/golden/*.npz``.  Nothing here is
imported by the product; tests only read the .npz files.

Reference entry points exercised (file:line under /root/reference):
  rtrec/models/slim.py:28-64 (fit / bulk_fit), rtrec/models/base.py:72-94 (add_interactions),
  rtrec/utils/interactions.py:81-119,259-303, rtrec/models/internal/slim_elastic.py:229-281,510-857.
"""
import os
import sys
import tempfile
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    stubs = tempfile.mkdtemp(prefix="rtrec_stubs_")
    os.makedirs(os.path.join(stubs, "implicit", "cpu"))
    os.makedirs(os.path.join(stubs, "lightfm"))
    open(os.path.join(stubs, "implicit", "__init__.py"), "w").close()
    open(os.path.join(stubs, "implicit", "cpu", "__init__.py"), "w").close()
    with open(os.path.join(stubs, "implicit", "cpu", "topk.py"), "w") as f:
        f.write("def topk(*a, **k):\n    raise NotImplementedError\n")
    with open(os.path.join(stubs, "lightfm", "__init__.py"), "w") as f:
        f.write("class LightFM:\n    pass\n")
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, stubs)
    from rtrec.models.slim import SLIM  # noqa
    from rtrec.utils.interactions import UserItemInteractions  # noqa
    return SLIM, UserItemInteractions


def synth_events(n_users, n_items, n_events, seed, rating="int", dup_frac=0.0, span_days=400.0):
    """Popularity-skewed synthetic events (recipe of SURVEY.md 8d, scaled down)."""
    rng = np.random.default_rng(seed)
    p_item = (np.arange(n_items) + 1.0) ** -0.9
    p_item /= p_item.sum()
    p_user = rng.lognormal(0, 1, n_users)
    p_user /= p_user.sum()
    u = rng.choice(n_users, size=int(n_events * 1.7), p=p_user)
    i = rng.choice(n_items, size=int(n_events * 1.7), p=p_item)
    key = u.astype(np.int64) * n_items + i
    _, first = np.unique(key, return_index=True)
    first.sort()
    first = first[:n_events]
    u, i = u[first], i[first]
    n_dup = int(len(u) * dup_frac)
    if n_dup:
        d = rng.integers(0, len(u), n_dup)
        u = np.concatenate([u, u[d]])
        i = np.concatenate([i, i[d]])
        perm = rng.permutation(len(u))
        u, i = u[perm], i[perm]
    n = len(u)
    ts = 1.0e9 + np.sort(rng.integers(0, int(span_days * 86400), n)).astype(np.float64)
    if rating == "int":
        r = rng.integers(1, 6, n).astype(np.float64)
    elif rating == "half":
        r = rng.integers(1, 11, n).astype(np.float64) * 0.5
    elif rating == "one":
        r = np.ones(n)
    else:  # "cont": tie-free
        r = rng.uniform(0.5, 5.0, n)
    return u.astype(np.int64), i.astype(np.int64), ts, r


def w_arrays(W):
    W = sp.csc_matrix(W)
    W.sort_indices()
    return W.data.astype(np.float32), W.indices.astype(np.int32), W.indptr.astype(np.int32)


def ref_sel(X_csc, items, nn):
    """Replays slim_elastic.py:141-143 with the same numpy calls to record the reference's picks."""
    X = X_csc.copy()
    out = np.full((len(items), nn), -1, dtype=np.int32)
    for t, j in enumerate(items):
        y = X.getcol(j).copy()
        a, b = X.indptr[j], X.indptr[j + 1]
        X.data[a:b] = 0
        s = X.T.dot(y.toarray().ravel()).flatten()
        sel = np.argsort(s)[-1:-1 - nn:-1]
        out[t, :len(sel)] = sel
        X.data[a:b] = y.data
    return out


def run_case(SLIM, name, *, n_users, n_items, n_events, seed, rating, kwargs, dup_frac=0.0, string_ids=False,
             partial_events=0, mode="bulk"):
    u, i, ts, r = synth_events(n_users, n_items, n_events + partial_events, seed, rating, dup_frac)
    n0 = len(u) - partial_events
    model = SLIM(**kwargs)
    uid = (lambda x: f"u{x}") if string_ids else int
    iid = (lambda x: f"i{x}") if string_ids else int
    ev = [(uid(a), iid(b), float(c), float(d)) for a, b, c, d in zip(u[:n0], i[:n0], ts[:n0], r[:n0])]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if mode == "bulk":
            model.add_interactions(ev)
            X0 = model.interactions.to_csc()
            model.bulk_fit(parallel=False, progress_bar=False)
        else:  # "fit": SLIM.fit(iterable) -> partial_fit_items path, float32 LIL
            items0 = list({model.item_ids.identify(e[1]) for e in ev}) if False else None
            model.fit(ev, progress_bar=False)
            X0 = model.interactions.to_csc()
    out = {}
    nn = kwargs.get("nn_feature_selection")
    out["events"] = np.stack([u.astype(np.float64), i.astype(np.float64), ts, r], axis=1)
    out["n0"] = np.array(n0)
    X0 = sp.csc_matrix(X0)
    X0.sort_indices()
    out["X0_data"], out["X0_indices"], out["X0_indptr"] = X0.data.astype(np.float32), X0.indices.astype(np.int32), X0.indptr.astype(np.int32)
    out["X0_shape"] = np.array(X0.shape)
    out["W0_data"], out["W0_indices"], out["W0_indptr"] = w_arrays(model.model.item_similarity)
    if nn and mode == "bulk":
        out["sel0"] = ref_sel(X0, list(range(X0.shape[1])), nn)
    if mode == "fit":
        # SLIM.fit solves the columns in ``list(set(item ids))`` order on a matrix holding only those columns
        ids = [model.item_ids.identify(e[1]) for e in ev]
        order = list(set(ids))
        out["fit_items0"] = np.array(order, dtype=np.int32)
        if nn:
            out["sel0"] = ref_sel(X0, order, nn)
    if partial_events:
        ev1 = [(uid(a), iid(b), float(c), float(d)) for a, b, c, d in zip(u[n0:], i[n0:], ts[n0:], r[n0:])]
        ids1 = []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.fit(ev1, update_interaction=True, progress_bar=False)
        ids1 = list({model.item_ids.identify(e[1]) for e in ev1})
        # the reference iterates list(set(...)) built inside fit(); rebuild the same set/order
        s = set()
        for e in ev1:
            s.add(model.item_ids.identify(e[1]))
        order1 = list(s)
        X1 = sp.csc_matrix(model.interactions.to_csc(order1))
        X1.sort_indices()
        out["fit_items1"] = np.array(order1, dtype=np.int32)
        out["X1_data"], out["X1_indices"], out["X1_indptr"] = X1.data.astype(np.float32), X1.indices.astype(np.int32), X1.indptr.astype(np.int32)
        out["X1_shape"] = np.array(X1.shape)
        out["W1_data"], out["W1_indices"], out["W1_indptr"] = w_arrays(model.model.item_similarity)
        if nn:
            out["sel1"] = ref_sel(X1, order1, nn)
    # scoring: top-10 for every known user, similar items for every item
    if mode == "bulk" and not string_ids:
        # serial bulk_fit leaves a float64 W; the sparse path accepts it
        pass
    users = sorted(set(u.tolist()))
    W = model.model.item_similarity
    if W.dtype != np.float32:
        model.model.item_similarity = W.astype(np.float32)
    recs = model.recommend_batch([uid(x) for x in users], top_k=10, filter_interacted=True)
    rec_arr = np.full((len(users), 10), -1, dtype=np.int64)
    for k, lst in enumerate(recs):
        ids = [int(x[1:]) if string_ids else int(x) for x in lst]
        rec_arr[k, :len(ids)] = ids
    out["rec_users"] = np.array(users, dtype=np.int64)
    out["rec_top10"] = rec_arr
    # dense scores for tie-aware comparison
    Xcsr = model.interactions.to_csr()
    S = (Xcsr @ model.model.item_similarity).toarray().astype(np.float32)
    out["scores_dense"] = S if S.size <= 400_000 else np.zeros((0, 0), np.float32)
    sim_ids = np.full((X0.shape[1], 10), -1, dtype=np.int64)
    sim_sc = np.zeros((X0.shape[1], 10), dtype=np.float32)
    n_items_now = model.model.item_similarity.shape[1]
    for j in range(min(X0.shape[1], n_items_now)):
        res = model.model.similar_items(j, top_k=10)
        for k, (a, b) in enumerate(res):
            sim_ids[j, k] = a
            sim_sc[j, k] = b
    out["sim_ids"], out["sim_scores"] = sim_ids, sim_sc
    out["pass_through"] = np.array(bool(model.item_ids.pass_through))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "X", X0.shape, "nnzW", model.model.item_similarity.nnz)


def store_case(UII, name, *, seed, decay, upsert, n=4000):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, 60, n)
    i = rng.integers(0, 40, n)
    ts = 1.0e9 + rng.uniform(0, 90 * 86400, n)  # out of order on purpose
    d = rng.integers(-3, 6, n).astype(np.float64)
    st = UII(min_value=-5, max_value=10, decay_in_days=decay)
    for a, b, c, e in zip(u, i, ts, d):
        st.add_interaction(int(a), int(b), float(c), float(e), upsert=upsert)
    X = sp.csc_matrix(st.to_csc())
    X.sort_indices()
    sel = [1, 5, 7, 30]
    Xs = sp.csc_matrix(st.to_csc(sel))
    Xs.sort_indices()
    R = sp.csr_matrix(st.to_csr([3, 9, 11]))
    R.sort_indices()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        events=np.stack([u.astype(np.float64), i.astype(np.float64), ts, d], axis=1),
        decay=np.array(-1 if decay is None else decay), upsert=np.array(upsert),
        X_data=X.data.astype(np.float32), X_indices=X.indices.astype(np.int32), X_indptr=X.indptr.astype(np.int32),
        X_shape=np.array(X.shape), sel_items=np.array(sel),
        Xs_data=Xs.data.astype(np.float32), Xs_indices=Xs.indices.astype(np.int32), Xs_indptr=Xs.indptr.astype(np.int32),
        sel_users=np.array([3, 9, 11]),
        R_data=R.data.astype(np.float32), R_indices=R.indices.astype(np.int32), R_indptr=R.indptr.astype(np.int32),
        max_timestamp=np.array(st.max_timestamp), max_user_id=np.array(st.max_user_id),
        max_item_id=np.array(st.max_item_id),
        hot_items=np.array(list(st.hot_items.get_freq_items(20)), dtype=np.int64))
    print(name, "nnz", X.nnz)


def main():
    SLIM, UII = _import_reference()
    run_case(SLIM, "slim_all_int", n_users=300, n_items=90, n_events=5000, seed=11, rating="int", kwargs={})
    run_case(SLIM, "slim_nn20_int", n_users=400, n_items=150, n_events=8000, seed=12, rating="int",
             kwargs={"nn_feature_selection": 20})
    run_case(SLIM, "slim_nn20_cont", n_users=400, n_items=150, n_events=8000, seed=13, rating="cont",
             kwargs={"nn_feature_selection": 20})
    run_case(SLIM, "slim_nn20_decay", n_users=400, n_items=150, n_events=8000, seed=14, rating="half",
             kwargs={"nn_feature_selection": 20, "decay_in_days": 180}, dup_frac=0.15)
    run_case(SLIM, "slim_all_decay_partial", n_users=300, n_items=90, n_events=5000, seed=15, rating="cont",
             kwargs={"decay_in_days": 30}, partial_events=600)
    run_case(SLIM, "slim_nn20_partial", n_users=400, n_items=150, n_events=8000, seed=16, rating="cont",
             kwargs={"nn_feature_selection": 20}, partial_events=900)
    run_case(SLIM, "slim_all_strids_fit", n_users=200, n_items=60, n_events=3000, seed=17, rating="cont",
             kwargs={}, string_ids=True, mode="fit")
    for k, (decay, upsert) in enumerate([(None, False), (None, True), (30, False), (30, True)]):
        store_case(UII, f"store_{k}", seed=20 + k, decay=decay, upsert=upsert)


if __name__ == "__main__":
    main()
