"""Model files written by the reference load into this package.

Fixtures: tests/golden/ref_model_*.pkl are the bytes of the reference's own ``SLIM.save`` and ref_model_*.npz
what the reference answers on them (tests/golden/make_ref_pickles.py, run where /root/reference exists).  The
reference package is not importable where these tests run; the loader must not need it.
"""
import io
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from tests.helpers import topk_consistent

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["int", "str_decay"]


def _load(name):
    from rtrec_b200.models import SLIM
    with open(os.path.join(GOLD, f"ref_model_{name}.pkl"), "rb") as f:
        m = SLIM.load(f)
    z = np.load(os.path.join(GOLD, f"ref_model_{name}.npz"))
    return m, z


@pytest.mark.parametrize("name", CASES)
def test_reference_pickle_loads_without_the_reference(name):
    """Store, W, id maps, tags and hot items of a reference model file, bit for bit, with no GPU and without
    importing ``rtrec``."""
    assert "rtrec" not in sys.modules
    m, z = _load(name)
    assert "rtrec" not in sys.modules
    st = m.interactions
    keys, vals, stamps = st._host_state
    assert np.array_equal(keys >> np.uint64(32), z["store_u"].astype(np.uint64))
    assert np.array_equal(keys & np.uint64(0xFFFFFFFF), z["store_i"].astype(np.uint64))
    assert np.array_equal(vals, z["store_v"]) and np.array_equal(stamps, z["store_ts"])
    assert (st.max_user_id, st.max_item_id, st.max_timestamp) == (int(z["max_user_id"]), int(z["max_item_id"]), float(z["max_timestamp"]))
    assert st.all_item_ids == set(int(x) for x in z["store_i"])
    W = m.model.item_similarity
    Wr = sp.csc_matrix((z["W_data"], z["W_indices"], z["W_indptr"]), shape=tuple(z["W_shape"]))
    assert W.shape == Wr.shape and W.dtype == np.float32
    W.sort_indices()
    assert np.array_equal(W.indptr, Wr.indptr) and np.array_equal(W.indices, Wr.indices)
    assert np.array_equal(W.data.astype(np.float64), Wr.data)          # float32 values held in a float64 container
    if name == "int":
        assert m.item_ids.pass_through is True and m.model.nn_feature_selection == 10
        assert [str(x) for x in st.hot_items.get_freq_items(5)] == list(z["hot"]) or len(z["hot"]) == 0
    else:
        assert m.item_ids.pass_through is False
        assert m.user_ids.identify("user_3") == m.user_ids.obj_to_id["user_3"]
        assert st.decay_rate == pytest.approx(1.0 - np.log(2) / 30) and (st.min_value, st.max_value) == (-2, 6)
        assert m.feature_store.num_user_features() == 2 and m.feature_store.num_item_features() == 1
    # and the result is an ordinary model of this package: it pickles again
    buf = io.BytesIO()
    m.save(buf)
    m2 = type(m).loads(buf.getvalue())
    assert np.array_equal(m2.interactions._host_state[0], keys)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_reference_pickle_serves_like_the_reference(name):
    """recommend_batch / similar_items on the loaded file against the reference's own answers: every list is a valid
    top-5 of X.W over the eligible items (score ties may order differently), most lists are
    identical, cold users get the reference's hot items."""
    m, z = _load(name)
    str_ids = name != "int"
    users = [x if str_ids else int(x) for x in z["users"]]
    got = m.recommend_batch(users, top_k=5, filter_interacted=True)
    want = [[y if str_ids else int(y) for y in r.split(",") if y != ""] for r in z["recs"]]
    # dense scores from the fixture
    U, I = int(z["max_user_id"]) + 1, int(z["max_item_id"]) + 1
    v = z["store_v"].copy()
    rate = m.interactions.decay_rate
    if rate is not None:
        v = v * rate ** ((float(z["max_timestamp"]) - z["store_ts"]) / 86400.0)
    X = sp.csr_matrix((v.astype(np.float32), (z["store_u"], z["store_i"])), shape=(U, I))
    W = sp.csc_matrix((z["W_data"].astype(np.float32), z["W_indices"], z["W_indptr"]), shape=tuple(z["W_shape"]))
    S = np.asarray((X @ W).todense(), dtype=np.float32)
    same = 0
    for user, g, w in zip(users, got, want):
        uid = m.user_ids.get_id(user) if str_ids else (user if user <= int(z["max_user_id"]) else None)
        if uid is None:
            assert g == w, (user, g, w)      # cold user: hot items, frequency order
            same += 1
            continue
        gi = [m.item_ids.get_id(x) if str_ids else x for x in g]
        inter = np.zeros(I, bool)
        inter[X[uid].indices] = True
        elig = ~inter if str_ids else (~inter & (S[uid] != 0))
        ok, why = topk_consistent(gi, S[uid], 5, elig, tol=2e-5)
        assert ok, (user, why, g, w)
        same += int(g == w)
    assert same >= 0.8 * len(users), f"only {same}/{len(users)} lists identical to the reference's"
    for q, line in zip(z["q_items"], z["sims"]):
        ref = [(a if str_ids else int(a), float(b)) for a, b in (p.rsplit(":", 1) for p in line.split(";") if p)]
        mine = m.similar_items(q if str_ids else int(q), top_k=3, ret_scores=True)
        assert [round(s, 6) for _, s in mine] == [round(s, 6) for _, s in ref]
        assert set(x for x, _ in mine) == set(x for x, _ in ref) or len(set(s for _, s in ref)) < len(ref)


def test_reference_pickle_of_an_unfitted_model():
    """``SLIM(n_recent_hot=7).save`` of the reference before any event: loads as an empty, unfitted model (W is None, the
    hot-item capacity survives), and `recommend` on it is the empty list like tests/models/test_slim.py:37-40 -- checked
    on the host only, no device needed for an empty store."""
    from rtrec_b200.models import SLIM
    with open(os.path.join(GOLD, "ref_model_empty.pkl"), "rb") as f:
        m = SLIM.load(f)
    assert m.model.item_similarity is None
    assert m.interactions._n_pairs == 0 and m.interactions._host_state[0].size == 0
    assert m.interactions.hot_items.capacity == 7 and len(m.interactions.hot_items) == 0
    assert m.interactions.max_user_id == 0 and m.interactions.max_item_id == 0
    assert m.user_ids.pass_through is None and m.item_ids.pass_through is None
