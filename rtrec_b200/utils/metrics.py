"""Top-k ranking metrics for ``Recommender.evaluate`` (host side).

Same definitions and edge cases as /root/reference/rtrec/utils/metrics.py:6-313 (pinned by the
known answers of /root/reference/tests/utils/test_metrics.py); written around one shared
relevance vector per query instead of one scan per metric.
"""
from __future__ import annotations

from collections import defaultdict
from math import log2
from typing import Any, Dict, Iterable, List, Tuple


def _hits(ranked_list: List[Any], ground_truth: List[Any], recommend_size: int) -> List[bool]:
    k = min(len(ranked_list), recommend_size)
    return [ranked_list[i] in ground_truth for i in range(k)]


def true_positives(ranked_list, ground_truth, recommend_size: int) -> int:
    return sum(_hits(ranked_list, ground_truth, recommend_size))


def precision(ranked_list, ground_truth, recommend_size: int) -> float:
    if not ground_truth:
        return 0.0 if ranked_list else 1.0
    rel = _hits(ranked_list, ground_truth, recommend_size)
    return sum(rel) / len(rel) if rel else 0.0


def recall(ranked_list, ground_truth, recommend_size: int) -> float:
    if not ground_truth:
        return 0.0 if ranked_list else 1.0
    return sum(_hits(ranked_list, ground_truth, recommend_size)) / len(ground_truth)


def f1_score(ranked_list, ground_truth, recommend_size: int) -> float:
    if not ground_truth and not ranked_list:
        return 1.0
    p = precision(ranked_list, ground_truth, recommend_size)
    r = recall(ranked_list, ground_truth, recommend_size)
    return 2 * (p * r) / (p + r) if (p + r) > 0 else 0.0


def ndcg(ranked_list, ground_truth, recommend_size: int) -> float:
    rel = _hits(ranked_list, ground_truth, recommend_size)
    dcg = sum(1 / log2(pos + 2) for pos, hit_ in enumerate(rel) if hit_)
    ideal = sum(1 / log2(pos + 2) for pos in range(min(len(ground_truth), recommend_size)))
    return dcg / ideal if ideal > 0 else 0.0


def hit(ranked_list, ground_truth, recommend_size: int) -> float:
    return 1.0 if any(item in ground_truth for item in ranked_list[:recommend_size]) else 0.0


def reciprocal_rank(ranked_list, ground_truth, recommend_size: int) -> float:
    for pos, hit_ in enumerate(_hits(ranked_list, ground_truth, recommend_size)):
        if hit_:
            return 1.0 / (pos + 1)
    return 0.0


def auc(ranked_list, ground_truth, recommend_size: int) -> float:
    if not ground_truth:
        return 0.0 if ranked_list else 1.0
    if not ranked_list:
        return 0.0
    rel = _hits(ranked_list, ground_truth, recommend_size)
    tp = 0
    ordered_pairs = 0  # (relevant ranked above irrelevant)
    for hit_ in rel:
        if hit_:
            tp += 1
        else:
            ordered_pairs += tp
    fp = len(rel) - tp
    if tp == 0:
        return 0.0
    if fp == 0:
        return 1.0
    return ordered_pairs / (tp * fp)


def average_precision(ranked_list, ground_truth, recommend_size: int) -> float:
    if not ground_truth:
        return 0.0 if ranked_list else 1.0
    running, total = 0, 0.0
    for pos, hit_ in enumerate(_hits(ranked_list, ground_truth, recommend_size)):
        if hit_:
            running += 1
            total += running / (pos + 1)
    denom = min(len(ground_truth), recommend_size)
    return total / denom if denom > 0 else 0.0


def _mean_over(fn, ranked_lists, ground_truths, recommend_size) -> float:
    ranked_lists = ranked_lists if isinstance(ranked_lists, list) else list(ranked_lists)
    acc = sum(fn(r, g, recommend_size) for r, g in zip(ranked_lists, ground_truths))
    return acc / len(ranked_lists) if ranked_lists else 0.0


def mrr(ranked_lists: Iterable[List[Any]], ground_truths: Iterable[List[Any]], recommend_size: int) -> float:
    return _mean_over(reciprocal_rank, ranked_lists, ground_truths, recommend_size)


def map_score(ranked_lists: Iterable[List[Any]], ground_truths: Iterable[List[Any]], recommend_size: int) -> float:
    return _mean_over(average_precision, ranked_lists, ground_truths, recommend_size)


_PER_QUERY = (("precision", precision), ("recall", recall), ("f1", f1_score), ("ndcg", ndcg), ("hit_rate", hit),
              ("mrr", reciprocal_rank), ("map", average_precision), ("auc", auc))


def compute_scores(evaluation_pairs: Iterable[Tuple[List[Any], List[Any]]], recommend_size: int) -> Dict[str, float]:
    """Averages of the per-query metrics plus the total number of true positives (``tp``)."""
    sums = {name: 0.0 for name, _ in _PER_QUERY}
    tp_total = 0
    n = 0
    for ranked_list, ground_truth in evaluation_pairs:
        n += 1
        for name, fn in _PER_QUERY:
            sums[name] += fn(ranked_list, ground_truth, recommend_size)
        tp_total += true_positives(ranked_list, ground_truth, recommend_size)
    if n == 0:
        return defaultdict(float)
    out = {name: s / n for name, s in sums.items()}
    out["tp"] = tp_total
    # key order of the reference dict
    return {k: out[k] for k in ("precision", "recall", "f1", "ndcg", "hit_rate", "mrr", "map", "tp", "auc")}
