"""Golden vectors for the evaluation metrics, produced by the REAL reference (run in the build container only):

    python tests/golden/make_eval_golden.py

Random ranked lists / ground-truth lists (ragged, with duplicates in the ground truth, empty recommendation lists,
lists longer than recommend_size) are pushed through /root/reference/rtrec/utils/metrics.py::compute_scores and the
per-query functions; inputs and outputs go to tests/golden/eval_cases.npz.  Tests only read the .npz.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/rtrec/utils/metrics.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(7)
    out = {"python": np.asarray(sys.version_info[:3])}
    fns = [m.precision, m.recall, m.f1_score, m.ndcg, m.hit, m.reciprocal_rank, m.average_precision, m.true_positives, m.auc]
    for case, (Q, k, n_items, rs) in enumerate([(400, 10, 60, 10), (300, 20, 200, 10), (200, 5, 30, 10), (150, 50, 400, 50),
                                                 (64, 128, 1000, 128)]):
        ids = np.full((Q, k), -1, dtype=np.int32)
        cnt = np.zeros(Q, dtype=np.int32)
        gptr = [0]
        gidx = []
        pairs = []
        for q in range(Q):
            c = int(rng.integers(0, k + 1)) if q % 7 else (0 if q % 14 else k)
            row = rng.choice(n_items, size=c, replace=False)
            ids[q, :c] = row
            cnt[q] = c
            ng = int(rng.integers(1, 12))
            g = rng.integers(0, n_items + 20, ng)          # some ids outside the catalogue
            if q % 5 == 0:
                g = np.concatenate([g, g[:2]])              # duplicates count towards len(ground_truth)
            if c and q % 3 == 0:
                g = np.concatenate([g, row[: max(1, c // 3)]])   # guaranteed hits
            gidx.extend(g.tolist())
            gptr.append(len(gidx))
            pairs.append((row.tolist(), g.tolist()))
        res = m.compute_scores(iter(pairs), rs)
        per = np.asarray([[float(f(r, g, rs)) for f in fns] for r, g in pairs], dtype=np.float64)
        out[f"c{case}_ids"] = ids
        out[f"c{case}_cnt"] = cnt
        out[f"c{case}_gptr"] = np.asarray(gptr, dtype=np.int64)
        out[f"c{case}_gidx"] = np.asarray(gidx, dtype=np.int64)
        out[f"c{case}_rs"] = np.asarray(rs)
        out[f"c{case}_per_query"] = per
        keys = ("precision", "recall", "f1", "ndcg", "hit_rate", "mrr", "map", "tp", "auc")
        out[f"c{case}_scores"] = np.asarray([float(res[k_]) for k_ in keys], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "eval_cases.npz"), **out)
    print("wrote eval_cases.npz")


if __name__ == "__main__":
    main()
