"""Model-level behaviour pinned by the reference's own tests, replayed against the GPU package
(/root/reference/tests/models/test_slim.py, tests/models/test_serialization.py:14-58,152-180,
README.md:49-73)."""
import io
import time

import numpy as np
import pandas as pd
import pytest

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
