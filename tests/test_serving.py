"""Serving callers.  The route tests replay /root/reference/tests/serving/test_app.py:13-130 against
rtrec_b200.serving.app (same payloads, same expected answers); the token / root cases need no GPU."""
import json
import time

import pytest
from fastapi.testclient import TestClient

from rtrec_b200.serving.app import ModelGate, StreamIngestor, create_app

SECRET_TOKEN = "fake_secret_token"


@pytest.fixture
def client():
    return TestClient(create_app())


def test_read_root(client):                         # test_app.py:13-16
    r = client.get("/")
    assert r.status_code == 200 and r.json() == {"message": "Recommender System API is running"}


def test_fit_invalid_token(client):                 # test_app.py:33-45
    r = client.post("/fit", json=[{"user": "user1", "item": "item1", "timestamp": 1672531200.0, "rating": 5.0}],
                    headers={"X-Token": "wrong_token"})
    assert r.status_code == 400 and r.json() == {"detail": "Invalid X-Token header"}


def test_recommend_invalid_token(client):           # test_app.py:116-128
    r = client.post("/recommend", json={"user": "user1", "top_k": 5, "filter_interacted": True}, headers={"X-Token": "wrong_token"})
    assert r.status_code == 400 and r.json() == {"detail": "Invalid X-Token header"}


def test_fit_without_gpu_reports_training_failed(client):
    """No CPU fallback: without a CUDA device the model call raises and the route answers 500 like the reference does
    for any training error (app.py:66-68)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = client.post("/fit", json=[{"user": "user1", "item": "item1", "timestamp": 1672531200.0, "rating": 5.0}],
                    headers={"X-Token": SECRET_TOKEN})
    assert r.status_code == 500 and r.json() == {"detail": "Training failed"}


def test_stream_record_parsing():
    rec = {"user": "u1", "item": 3, "timestamp": 1.0e9, "rating": 2.5}
    want = ("u1", 3, 1.0e9, 2.5)
    assert StreamIngestor.parse(rec) == want
    assert StreamIngestor.parse(json.dumps(rec)) == want
    assert StreamIngestor.parse(json.dumps(rec).encode()) == want
    assert StreamIngestor.parse({"Data": json.dumps(rec).encode()}) == want          # raw Kinesis record
    with pytest.raises(Exception):
        StreamIngestor.parse(b"not json")


@pytest.mark.gpu
def test_fit(client):                               # test_app.py:18-31
    ev = [{"user": "user1", "item": "item1", "timestamp": 1672531200.0, "rating": 5.0},
          {"user": "user1", "item": "item2", "timestamp": 1672617600.0, "rating": 3.0},
          {"user": "user2", "item": "item1", "timestamp": 1672704000.0, "rating": 4.0}]
    r = client.post("/fit", json=ev, headers={"X-Token": SECRET_TOKEN})
    assert r.status_code == 200 and r.json() == {"message": "Training successful"}


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["str", "int"])    # test_app.py:48-80 and 82-114
def test_recommend(client, kind):
    U = (lambda k: f"user{k}") if kind == "str" else (lambda k: k)
    I = (lambda k: f"item{k}") if kind == "str" else (lambda k: k)
    ev = [{"user": U(1), "item": I(1), "timestamp": 1672531200.0, "rating": 5.0},
          {"user": U(1), "item": I(2), "timestamp": 1672617600.0, "rating": 3.0},
          {"user": U(2), "item": I(1), "timestamp": 1672704000.0, "rating": 4.0},
          {"user": U(2), "item": I(3), "timestamp": 1672704000.0, "rating": 4.0},
          {"user": U(2), "item": I(4), "timestamp": 1672704000.0, "rating": 3.0}]
    client.post("/fit", json=ev, headers={"X-Token": SECRET_TOKEN})
    r = client.post("/recommend", json={"user": U(1), "top_k": 5, "filter_interacted": True}, headers={"X-Token": SECRET_TOKEN})
    assert r.status_code == 200
    j = r.json()
    assert j["user"] == U(1) and isinstance(j["recommendations"], list) and len(j["recommendations"]) == 2
    assert all(it in [I(3), I(4)] for it in j["recommendations"])


@pytest.mark.gpu
def test_stream_ingestor_microbatches_equal_one_fit_per_batch():
    """Records pushed one by one end in the same store as the same events passed to fit() batch by batch, with far fewer
    partial fits than records; malformed records are dropped."""
    import numpy as np
    from rtrec_b200.models import SLIM
    from rtrec_b200.utils.synth import synth_events
    u, i, ts, r = synth_events(200, 80, 3000, seed=4, rating="cont")
    recs = [json.dumps({"user": int(a), "item": int(b), "timestamp": float(c), "rating": float(d)}).encode()
            for a, b, c, d in zip(u, i, ts, r)]
    m = SLIM(nn_feature_selection=10)
    ing = StreamIngestor(ModelGate(m), max_batch=1024, max_wait_s=0.05)
    for k in range(0, len(recs), 50):
        ing.put(recs[k:k + 50] + ([b"garbage"] if k == 100 else []))
    ing.flush()
    ing.close()
    assert ing.n_events == len(recs) and ing.n_batches < len(recs) // 50 + 1
    ref = SLIM(nn_feature_selection=10)
    ref.add_interactions(list(zip(u.tolist(), i.tolist(), ts.tolist(), r.tolist())))
    A, B = m.interactions.to_csr(), ref.interactions.to_csr()
    assert (A != B).nnz == 0 and np.array_equal(A.data, B.data)
    assert m.model.item_similarity is not None and len(m.recommend(0, top_k=5)) > 0
