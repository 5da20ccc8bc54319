"""Model-level behaviour pinned by the reference's own tests, replayed against the GPU package
(/root/reference/tests/models/test_slim.py, tests/models/test_serialization.py:14-58,152-180,
README.md:49-73)."""
import io
import time

import numpy as np
import pandas as pd
import pytest

from rtrec_b200.utils.synth import synth_events

pytestmark = pytest.mark.gpu


@pytest.fixture
def model():
    from rtrec_b200.models import SLIM
    return SLIM()


def test_register_features(model):            # test_slim.py:11-27
    uid = model.register_user_feature("user_1", ["tag1", "tag2"])
    assert uid == model.user_ids.identify("user_1")
    assert model.feature_store.build_user_features_matrix(user_ids=[uid]).nnz == 2
    iid = model.register_item_feature("item_1", ["tagA", "tagB"])
    assert iid == model.item_ids.identify("item_1")
    assert model.feature_store.build_item_features_matrix(item_ids=[iid]).nnz == 2


def test_add_interactions(model):             # test_slim.py:29-35
    model.add_interactions([("user_1", "item_1", 1622470427.0, 5.0)])
    u, i = model.user_ids.identify("user_1"), model.item_ids.identify("item_1")
    assert model.interactions.get_user_item_rating(u, i) == 5.0


def test_recommend_no_interactions(model):    # test_slim.py:37-40
    assert model.recommend("user_1", top_k=5) == []


def test_similar_items_no_data(model):        # test_slim.py:42-56
    now = time.time()
    ev = [("user_1", "item_1", 1622470427.0, 5.0), ("user_2", "item_2", now, -2.0)]
    model.fit(ev)
    model.fit(iter(ev))
    assert model.similar_items("item_1", top_k=5) == []


def test_similar_items(model):                # test_slim.py:58-79
    now = time.time()
    ev = [("user_1", "item_1", now, 5.0), ("user_1", "item_3", now, 4.0), ("user_1", "item_4", now, 3.0),
          ("user_2", "item_1", now, 3.0), ("user_2", "item_2", now, -2.0), ("user_2", "item_4", now, 3.0),
          ("user_3", "item_1", now, 4.0), ("user_3", "item_3", now, 2.0), ("user_3", "item_4", now, 4.0)]
    model.fit(ev)
    assert model.similar_items("item_1", top_k=5) == ["item_4", "item_3"]
    res = model.similar_items("item_1", top_k=5, ret_scores=True)
    items, scores = map(list, zip(*res))
    assert items == ["item_4", "item_3"] and scores[0] > scores[1]


def test_fit_and_recommend(model):            # test_slim.py:81-98, README.md:49-73
    now = time.time()
    ev = [("user_1", "item_1", now, 5.0), ("user_2", "item_2", now, -2.0), ("user_2", "item_1", now, 3.0),
          ("user_2", "item_4", now, 3.0), ("user_1", "item_3", now, 4.0)]
    model.fit(ev)
    model.fit(iter(ev))
    assert model.recommend("user_1", top_k=5) == ["item_4", "item_2"]


def test_get_users_by_items(model):           # test_slim.py:100-119
    now = time.time()
    model.fit([("user_1", "item_1", now, 5.0), ("user_2", "item_1", now, 3.0), ("user_2", "item_2", now, 4.0),
               ("user_3", "item_2", now, 2.0)])
    assert set(model.get_users_by_items(["item_1"])) == {"user_1", "user_2"}
    assert set(model.get_users_by_items(["item_1", "item_2"])) == {"user_1", "user_2", "user_3"}
    assert model.get_users_by_items(["item_99"]) == []


def test_recommend_batch(model):              # test_slim.py:121-174
    now = time.time()
    model.fit([("user_1", "item_1", now, 5.0), ("user_1", "item_3", now, 4.0), ("user_2", "item_2", now, 3.0),
               ("user_2", "item_4", now, 4.0), ("user_3", "item_1", now, 2.0), ("user_3", "item_2", now, 3.0)])
    users = ["user_1", "user_2", "user_3"]
    recs = model.recommend_batch(users, top_k=2)
    assert len(recs) == 3 and all(isinstance(r, list) for r in recs)
    cands = ["item_1", "item_2", "item_3"]
    for r in model.recommend_batch(users, candidate_items=cands, top_k=2):
        assert all(x in cands for x in r)
    assert model.recommend_batch([], top_k=2) == []
    one = model.recommend_batch(["user_1"], top_k=2)
    assert len(one) == 1 and isinstance(one[0], list)
    with_seen = model.recommend_batch(["user_1"], top_k=3, filter_interacted=False)
    assert any(x in ("item_1", "item_3") for x in with_seen[0])


def test_recommend_batch_cold_start(model):   # test_slim.py:176-194
    now = time.time()
    model.fit([("user_1", "item_1", now, 5.0), ("user_1", "item_2", now, 4.0)])
    recs = model.recommend_batch(["user_1", "new_user"], top_k=2)
    assert len(recs) == 2 and len(recs[0]) <= 2 and len(recs[1]) <= 2
    assert recs[1] == ["item_1", "item_2"][:2] or set(recs[1]) <= {"item_1", "item_2"}


def test_int_ids_use_sparse_topk():           # tests/serving/test_app.py:80-109 (pass-through ids)
    from rtrec_b200.models import SLIM
    m = SLIM(min_value=-5, max_value=10, decay_in_days=365)
    now = time.time()
    m.fit([(1, 1, now, 5.0), (1, 3, now, 4.0), (2, 2, now, 3.0), (2, 4, now, 4.0), (3, 1, now, 2.0), (3, 2, now, 3.0),
           (2, 1, now, 3.0), (3, 4, now, 1.0)])
    rec = m.recommend(1, top_k=2)
    assert set(rec) <= {2, 4} and len(rec) <= 2
    assert m.recommend(99, top_k=2) == list(m.interactions.hot_items.get_freq_items(2))


def test_save_load_roundtrip(model):          # test_serialization.py:14-58,152-180
    from rtrec_b200.models import SLIM
    now = time.time()
    ev = [("user_1", "item_1", now, 5.0), ("user_1", "item_3", now, 4.0), ("user_2", "item_2", now, 3.0),
          ("user_2", "item_4", now, 4.0), ("user_3", "item_1", now, 2.0), ("user_3", "item_2", now, 3.0),
          ("user_2", "item_1", now, 3.0)]
    model.fit(ev)
    before = model.recommend_batch(["user_1", "user_2", "user_3"], top_k=3)
    buf = io.BytesIO()
    n = model.save(buf)
    assert n > 0
    buf.seek(0)
    m2 = SLIM.load(buf)
    assert m2.recommend_batch(["user_1", "user_2", "user_3"], top_k=3) == before
    m3 = SLIM.loads(buf.getvalue())
    assert m3.similar_items("item_1", top_k=3) == model.similar_items("item_1", top_k=3)
    m3.fit([("user_4", "item_2", now + 5, 4.0)], update_interaction=True)  # a loaded model keeps learning
    assert m3.interactions.get_user_item_rating(m3.user_ids.get_id("user_4"), m3.item_ids.get_id("item_2")) == 4.0


def test_recommender_facade_end_to_end(golden):
    """Recommender.bulk_fit / fit(update_interaction) / evaluate on the golden partial-fit case:
    W after the bulk fit and after the streaming fit match the reference."""
    from rtrec_b200.models import SLIM
    from rtrec_b200.recommender import Recommender
    from tests.helpers import assert_w_parity, w_from
    z = golden("slim_all_decay_partial")
    ev, n0 = z["events"], int(z["n0"])
    df = pd.DataFrame({"user": ev[:, 0].astype(np.int64), "item": ev[:, 1].astype(np.int64), "tstamp": ev[:, 2],
                       "rating": ev[:, 3]})
    rec = Recommender(SLIM(decay_in_days=30))
    rec.bulk_fit(df.iloc[:n0], parallel=False)
    assert_w_parity(rec.model.model.item_similarity, w_from(z, "W0"), what="facade bulk")
    rec.fit(df.iloc[n0:], update_interaction=True)
    assert_w_parity(rec.model.model.item_similarity, w_from(z, "W1"), what="facade partial")
    scores = rec.evaluate(df.iloc[n0:][["user", "item"]], recommend_size=10, filter_interacted=False)
    assert set(scores) == {"precision", "recall", "f1", "ndcg", "hit_rate", "mrr", "map", "tp", "auc"}
    assert 0.0 <= scores["ndcg"] <= 1.0
    sims = rec.similar_items([0, 1], top_k=3, ret_scores=True)
    assert len(sims) == 2


def test_upload_events_threaded_staging():
    """rt_upload_events: pageable int64/f64 columns -> int32/f64 device columns, exact, any chunk count / thread
    count, id ranges and max timestamp reduced on the way."""
    import ctypes as C
    import torch
    from rtrec_b200 import _lib, device as D
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for n, threads in [(1, 0), (1000, 1), ((1 << 18) + 17, 3), (3 * (1 << 18) + 5, 0)]:
        u = rng.integers(0, 2**31 - 1, n).astype(np.int64)
        i = rng.integers(0, 50_000, n).astype(np.int64)
        ts = rng.uniform(1e9, 2e9, n)
        d = rng.normal(size=n)
        du, di = D.empty(n, torch.int32), D.empty(n, torch.int32)
        dts, dd = D.empty(n, torch.float64), D.empty(n, torch.float64)
        lo_u, hi_u, lo_i, hi_i, mx = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_double(0)
        _lib.check(lib.rt_upload_events(u.ctypes.data, i.ctypes.data, ts.ctypes.data, d.ctypes.data, n, D.ptr(du), D.ptr(di),
                                        D.ptr(dts), D.ptr(dd), C.byref(lo_u), C.byref(hi_u), C.byref(lo_i), C.byref(hi_i),
                                        C.byref(mx), threads, D.stream_ptr()))
        assert np.array_equal(du.cpu().numpy(), u.astype(np.int32))
        assert np.array_equal(di.cpu().numpy(), i.astype(np.int32))
        assert np.array_equal(dts.cpu().numpy(), ts) and np.array_equal(dd.cpu().numpy(), d)
        assert (lo_u.value, hi_u.value, lo_i.value, hi_i.value) == (u.min(), u.max(), i.min(), i.max())
        assert mx.value == ts.max()


def test_batch_ingest_device_path_equals_event_loop():
    """A batch large enough for the device ingest path leaves exactly the state the per-event path leaves
    (store matrix, maxima, all_item_ids, hot items), and rejects ids outside int32."""
    from rtrec_b200.utils import interactions as inter
    n = inter._DEVICE_INGEST_AT + 1000
    u, i, ts, r = synth_events(3000, 500, n, seed=11, rating="int", dup_frac=0.2)
    r = r - 2.0  # some non-positive deltas (hot_items counts only delta > 0)
    a = inter.UserItemInteractions(decay_in_days=30)
    a.add_interactions_batch(u, i, ts, r)
    b = inter.UserItemInteractions(decay_in_days=30)
    step = 9973  # below the device threshold: host bookkeeping path
    for s in range(0, n, step):
        b.add_interactions_batch(u[s:s + step], i[s:s + step], ts[s:s + step], r[s:s + step])
    A, B = a.to_csc(), b.to_csc()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)
    assert (a.max_user_id, a.max_item_id, a.max_timestamp) == (b.max_user_id, b.max_item_id, b.max_timestamp)
    assert a.all_item_ids == b.all_item_ids
    assert list(a.hot_items.data.items()) == list(b.hot_items.data.items())
    bad = u.copy(); bad[5] = 2**31
    with pytest.raises(ValueError):
        inter.UserItemInteractions().add_interactions_batch(bad, i, ts, r)


def test_recommend_lists_chunked_pipeline_equals_single_launch(golden):
    """SLIMElastic.recommend_lists (chunked launches + pinned D2H + host list building) returns what one
    recommend_batch_device call returns, including lists shorter than k."""
    from rtrec_b200.models import SLIM
    z = golden("slim_nn20_cont")
    ev = z["events"]
    m = SLIM(nn_feature_selection=20)
    m.add_interaction_arrays(ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64), ev[:, 2], ev[:, 3])
    m.bulk_fit()
    X = m.interactions.device_matrix()
    users = np.arange(X.n_users, dtype=np.int64)
    for dense in (False, True):
        ids, _, cnt = m.model.recommend_batch_device(users, X, None, 10, True, dense)
        want = [row[:c] for row, c in zip(ids.tolist(), cnt.tolist())]
        old = m.model._LIST_CHUNK
        try:
            m.model._LIST_CHUNK = 37
            got = m.model.recommend_lists(users, X, 10, True, dense)
        finally:
            m.model._LIST_CHUNK = old
        assert got == want
    ids, _, cnt = m.model.recommend_batch_device(np.arange(50), X, None, 10, True, False)
    assert m.recommend_batch(list(range(50)), top_k=10) == [row[:c] for row, c in zip(ids.tolist(), cnt.tolist())]


# ------------------------------------------------------------------------------------------ round-2 API hardening
def _small_model(ids_kind="int", nn=20):
    from rtrec_b200.models import SLIM
    u, i, ts, r = synth_events(500, 400, 20000, seed=31, rating="cont")
    m = SLIM(nn_feature_selection=nn)
    if ids_kind == "int":
        m.add_interaction_arrays(u, i, ts, r)
    else:
        m.add_interaction_arrays(np.asarray([f"u{x}" for x in u], dtype=object), np.asarray([f"i{x}" for x in i], dtype=object), ts, r)
    m.bulk_fit()
    return m, u, i


@pytest.mark.parametrize("dense_output", [True, False])
def test_top_k_above_128_is_not_truncated(dense_output):
    """top_k > 128: the reference returns up to top_k items; the fused kernels keep 128 per user, so longer
    lists come from the candidate-scoring kernel + the reference's selection rule.  The first 128 entries must equal the
    fused kernel's list wherever scores are untied, and the whole list must be a valid top-k of X @ W."""
    import scipy.sparse as sp
    from tests.helpers import topk_consistent
    m, u, i = _small_model()
    X = m.interactions.to_csr()
    W = m.model.item_similarity.tocsc().astype(np.float32)
    users = np.arange(0, 60)
    k = 200
    out = m.model.recommend_batch(users.tolist(), X, top_k=k, filter_interacted=True, dense_output=dense_output, ret_scores=True)
    S = np.asarray((X[users, :] @ W).todense(), dtype=np.float32)
    n_items = X.shape[1]
    for q, (items, scores) in enumerate(out):
        inter = np.zeros(n_items, bool)
        inter[X[int(users[q])].indices] = True
        elig = ~inter if dense_output else (~inter & (S[q] != 0))
        assert len(items) == min(k, int(elig.sum())), (q, len(items))
        ok, why = topk_consistent(items, S[q], k, elig, tol=2e-5)
        assert ok, (q, why)
    # candidates: more than 128 candidates and k > 128 -> every candidate comes back, ordered by score
    cand = list(range(0, 300))
    res = m.model.recommend_batch([3, 4], X, candidate_item_ids=cand, top_k=250)
    assert all(len(x) == 250 for x in res)
    # through the model API as well
    lists = m.recommend_batch([1, 2, 3], top_k=150)
    assert all(len(set(x)) == len(x) for x in lists) and max(len(x) for x in lists) > 128


def test_out_of_range_ids_raise_instead_of_reading_out_of_bounds():
    """user ids / candidate ids reach kernels that index without bounds tests; the operator rejects them like
    scipy indexing does (IndexError), the model treats unseen or negative pass-through ids as cold users / non-candidates."""
    m, u, i = _small_model()
    X = m.interactions.to_csr()
    n_users, n_items = X.shape
    for bad in ([n_users], [-1], [0, n_users + 7]):
        with pytest.raises(IndexError):
            m.model.recommend_batch(bad, X, top_k=5)
    with pytest.raises(IndexError):
        m.model.recommend_batch([0], X, candidate_item_ids=[1, n_items], top_k=5)
    with pytest.raises(IndexError):
        m.model.recommend_batch([0], X, candidate_item_ids=[-5, 1], top_k=5)
    with pytest.raises(IndexError):
        m.model.similar_items(n_items + 3)
    hot = m.interactions.get_hot_items(5, filter_interacted=False)
    assert m.recommend(-1, top_k=5) == hot                       # negative pass-through id = unknown user
    assert m.recommend_batch([-3, n_users + 10], top_k=5) == [hot, hot]
    got = m.recommend(0, candidate_items=[-5, 1, 2, n_items + 100], top_k=5)
    assert set(got) <= {1, 2}


def test_bad_events_are_skipped_one_by_one(caplog):
    """base.py:85-94: only the malformed event is dropped (an id >= 2^31 used to drop the whole call)."""
    from rtrec_b200.models import SLIM
    m = SLIM()
    ev = [(1, 2, 1.0e9, 3.0), (2**31 + 5, 3, 1.0e9, 1.0), (4, "x", 1.0e9, 1.0), (5, 6, "bad", 1.0), (7, 8, 1.0e9 + 1, 2.0)]
    m.add_interactions(ev)
    X = m.interactions.to_csr()
    assert X.nnz == 2 and X[1, 2] == 3.0 and X[7, 8] == 2.0
    # the column path falls back to the per-event path when a column is malformed
    m2 = SLIM()
    m2.add_interaction_arrays([1, 2**31 + 5, 7], [2, 3, 8], [1.0e9, 1.0e9, 1.0e9 + 1], [3.0, 1.0, 2.0])
    X2 = m2.interactions.to_csr()
    assert X2.nnz == 2 and X2[1, 2] == 3.0 and X2[7, 8] == 2.0
    m3 = SLIM()
    m3.add_interaction_arrays([1, "a", 7], [2, 3, 8], [1.0e9, 1.0e9, 1.0e9 + 1], [3.0, 1.0, 2.0])   # mixed id kinds
    assert m3.interactions.to_csr().nnz == 2


@pytest.mark.parametrize("ids_kind", ["int", "str"])
def test_similar_items_batch_equals_one_by_one(ids_kind):
    from rtrec_b200.recommender import Recommender
    m, u, i = _small_model(ids_kind)
    rec = Recommender(m)
    q = list(range(0, 120)) if ids_kind == "int" else [f"i{x}" for x in range(0, 120)]
    for ret_scores in (False, True):
        got = rec.similar_items(q, top_k=7, ret_scores=ret_scores)
        want = [m.similar_items(x, None, 7, ret_scores) for x in q]
        assert got == want
    assert any(len(x) for x in got)
