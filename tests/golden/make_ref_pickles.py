"""Generate model files with the REAL reference (myui/rtrec at /root/reference) -- fixtures for the
"load a reference-produced pickle" row.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_ref_pickles.py

For two small seeded models (integer ids / string ids with decay and tags) it writes
``tests/golden/ref_model_<name>.pkl`` = exactly the bytes of ``rtrec.models.SLIM.save``
(/root/reference/rtrec/models/base.py:376-384) and ``ref_model_<name>.npz`` = what the reference itself answers
on the loaded model: the store as sorted (user, item, value, stamp) columns, W (CSC), top-5 lists for every user
(``recommend_batch``), ``similar_items`` for a few items, and the hot-item order.
"""
import io
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference, synth_events  # noqa: E402


def main():
    SLIM, _ = _import_reference()
    buf0 = io.BytesIO()
    SLIM(n_recent_hot=7).save(buf0)                      # an unfitted model, before any event
    with open(os.path.join(HERE, "ref_model_empty.pkl"), "wb") as f:
        f.write(buf0.getvalue())
    cases = {
        "int": dict(kwargs=dict(nn_feature_selection=10), U=60, I=40, N=900, str_ids=False, seed=31),
        "str_decay": dict(kwargs=dict(decay_in_days=30, min_value=-2, max_value=6), U=40, I=30, N=500, str_ids=True, seed=32),
    }
    for name, c in cases.items():
        u, i, ts, r = synth_events(c["U"], c["I"], c["N"], c["seed"], rating="int", dup_frac=0.2)
        uu = [f"user_{x}" for x in u] if c["str_ids"] else [int(x) for x in u]
        ii = [f"item_{x}" for x in i] if c["str_ids"] else [int(x) for x in i]
        m = SLIM(**c["kwargs"])
        if c["str_ids"]:
            m.register_user_feature(uu[0], ["tag_a", "tag_b"])
            m.register_item_feature(ii[0], ["tag_x"])
        m.fit(list(zip(uu, ii, ts.tolist(), r.tolist())), update_interaction=False, progress_bar=False)
        buf = io.BytesIO()
        m.save(buf)
        raw = buf.getvalue()
        with open(os.path.join(HERE, f"ref_model_{name}.pkl"), "wb") as f:
            f.write(raw)
        m2 = SLIM.loads(raw)
        st = m2.interactions.interactions
        rows = sorted((int(uq), int(iq), float(v), float(t)) for uq, d in st.items() for iq, (v, t) in d.items())
        su, si, sv, sts = (np.array(x) for x in zip(*rows))
        W = sp.csc_matrix(m2.model.item_similarity)
        W.sort_indices()
        users = sorted(set(uu), key=str)
        if not c["str_ids"]:
            users = users + [10_000]           # a cold user: hot items
        recs = m2.recommend_batch(users, top_k=5, filter_interacted=True)
        q_items = sorted(set(ii), key=str)[:6]
        sims = [m2.similar_items(q, top_k=3, ret_scores=True) for q in q_items]
        hot = list(m2.interactions.get_hot_items(5, filter_interacted=False)) if hasattr(m2.interactions, "get_hot_items") else []
        np.savez_compressed(
            os.path.join(HERE, f"ref_model_{name}.npz"),
            store_u=su.astype(np.int64), store_i=si.astype(np.int64), store_v=sv.astype(np.float64), store_ts=sts.astype(np.float64),
            max_user_id=m2.interactions.max_user_id, max_item_id=m2.interactions.max_item_id,
            max_timestamp=m2.interactions.max_timestamp,
            W_data=W.data.astype(np.float64), W_indices=W.indices, W_indptr=W.indptr, W_shape=np.array(W.shape),
            users=np.array([str(x) for x in users]), recs=np.array([",".join(str(x) for x in r_) for r_ in recs]),
            q_items=np.array([str(x) for x in q_items]),
            sims=np.array([";".join(f"{a}:{b!r}" for a, b in s_) for s_ in sims]),
            hot=np.array([str(x) for x in hot]),
            n_user_features=m2.feature_store.user_features.__len__() if hasattr(m2.feature_store.user_features, "__len__") else 0,
        )
        print(name, len(raw), "bytes;", len(rows), "pairs; nnz(W) =", W.nnz, "; first recs:", recs[:2])


if __name__ == "__main__":
    main()
