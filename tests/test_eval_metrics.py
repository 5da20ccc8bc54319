"""``Recommender.evaluate``: the nine ranking metrics (/root/reference/rtrec/utils/metrics.py:6-313,
/root/reference/rtrec/recommender.py:163-200).

Golden vectors: tests/golden/eval_cases.npz, produced by the REAL reference's metrics module on random ragged lists
(tests/golden/make_eval_golden.py).  The host mirror (rtrec_b200/utils/metrics.py) and the device kernel
(rt_eval_metrics, csrc/eval.cu) must reproduce every per-query value and every mean BIT FOR BIT (float64)."""
import ctypes as C
import sys

import numpy as np
import pandas as pd
import pytest

from rtrec_b200.utils import metrics as M

KEYS = ("precision", "recall", "f1", "ndcg", "hit_rate", "mrr", "map", "tp", "auc")
FNS = (M.precision, M.recall, M.f1_score, M.ndcg, M.hit, M.reciprocal_rank, M.average_precision, M.true_positives, M.auc)
N_CASES = 5


def _case(z, c):
    ids, cnt, gptr, gidx = z[f"c{c}_ids"], z[f"c{c}_cnt"], z[f"c{c}_gptr"], z[f"c{c}_gidx"]
    pairs = [(ids[q, :cnt[q]].tolist(), gidx[gptr[q]:gptr[q + 1]].tolist()) for q in range(len(cnt))]
    return ids, cnt, gptr, gidx, int(z[f"c{c}_rs"]), pairs


def _same_python(z):
    return tuple(int(x) for x in z["python"][:2]) >= (3, 12) and sys.version_info >= (3, 12)


@pytest.mark.parametrize("c", range(N_CASES))
def test_host_metrics_match_reference_golden(golden, c):
    z = golden("eval_cases")
    if not _same_python(z):
        pytest.skip("golden made with CPython >= 3.12 (compensated float sum()); interpreter differs")
    ids, cnt, gptr, gidx, rs, pairs = _case(z, c)
    per = np.asarray([[float(f(r, g, rs)) for f in FNS] for r, g in pairs])
    assert np.array_equal(per, z[f"c{c}_per_query"])
    res = M.compute_scores(iter(pairs), rs)
    assert list(res.keys()) == list(KEYS)
    assert np.array_equal(np.asarray([float(res[k]) for k in KEYS]), z[f"c{c}_scores"])


@pytest.mark.gpu
@pytest.mark.parametrize("c", range(N_CASES))
def test_eval_kernel_matches_reference_golden(golden, c):
    """rt_eval_metrics through the C-ABI on the golden lists: per-query table and sequentially summed means bit-equal."""
    from math import log2
    from rtrec_b200 import _lib, device as D
    z = golden("eval_cases")
    ids, cnt, gptr, gidx, rs, pairs = _case(z, c)
    Q, k = ids.shape
    t = D.require_cuda()
    row = np.repeat(np.arange(Q), np.diff(gptr))
    order = np.lexsort((gidx, row))
    out = D.empty(Q * 9, t.float64)
    disc = D.to_dev(np.asarray([1 / log2(i + 2) for i in range(max(rs, k))]))
    d_ids, d_cnt, d_gptr = D.to_dev(ids), D.to_dev(cnt), D.to_dev(gptr)      # (named: the buffers must outlive the call)
    d_gidx = D.to_dev(np.clip(gidx[order], -1, 2**31 - 1).astype(np.int32))
    _lib.check(_lib.load().rt_eval_metrics(D.ptr(d_ids), D.ptr(d_cnt), Q, k, rs, D.ptr(d_gptr),
                                           D.ptr(d_gidx), D.ptr(disc),
                                           1 if tuple(z["python"][:2]) >= (3, 12) else 0, D.ptr(out), D.stream_ptr()),
               "rt_eval_metrics")
    per = out.view(Q, 9).cpu().numpy()
    assert np.array_equal(per, z[f"c{c}_per_query"])
    sums = np.cumsum(per, axis=0)[-1]
    got = [float(sums[m]) / Q for m in range(9)]
    got[7] = float(int(per[:, 7].sum()))
    assert np.array_equal(np.asarray(got), z[f"c{c}_scores"])


@pytest.mark.gpu
@pytest.mark.parametrize("ids_kind", ["int", "str"])
def test_evaluate_device_equals_list_loop(ids_kind):
    """Recommender.evaluate: device path (one scoring launch + rt_eval_metrics) == the reference's list-based loop over
    the same model, including unknown (cold-start) users, unknown ground-truth items and repeated ground-truth items."""
    from rtrec_b200.models import SLIM
    from rtrec_b200.recommender import Recommender
    from rtrec_b200.utils.metrics import compute_scores
    from rtrec_b200.utils.synth import synth_events
    u, i, ts, r = synth_events(900, 300, 30000, seed=21, rating="cont")
    conv = (lambda a, p: a) if ids_kind == "int" else (lambda a, p: np.asarray([f"{p}{x}" for x in a], dtype=object))
    df = pd.DataFrame({"user": conv(u, "u"), "item": conv(i, "i"), "tstamp": ts, "rating": r})
    train, test = df.iloc[:26000], df.iloc[26000:]
    extra = pd.DataFrame({"user": conv(np.asarray([5000, 5000, 5001]), "u"), "item": conv(np.asarray([3, 9999, 4]), "i"),
                          "tstamp": [ts[-1]] * 3, "rating": [1.0] * 3})       # cold users, one unknown item
    test = pd.concat([test, extra, test.iloc[:50]], ignore_index=True)        # repeated ground-truth rows
    rec = Recommender(SLIM(nn_feature_selection=20))
    rec.bulk_fit(train)
    got = rec.evaluate(test, recommend_size=10)
    grouped = test.groupby("user")["item"].apply(list).to_dict()
    users = list(grouped.keys())
    lists = rec.recommend_batch(users, top_k=10, filter_interacted=True)
    want = compute_scores(((l, grouped[uu]) for uu, l in zip(users, lists)), 10)
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k] == want[k], (k, got[k], want[k])
    assert got["tp"] > 0 and 0 < got["hit_rate"] <= 1
