"""Synthetic interaction generators for the BASELINE.json shapes (SURVEY.md 8d recipe).

Test/bench infrastructure shared by tests/, bench.py and tests/golden/make_golden.py: item popularity
p_i ~ (i+1)^-0.9 capped so the top item reaches <= 50 % of users, user activity lognormal(0,1),
unique (user,item) pairs unless ``dup_frac`` > 0, ascending integer timestamps from 1.0e9.
"""
from __future__ import annotations

import numpy as np

SHAPES = {
    # name: (n_users, n_items, n_events, rating kind, span days, seed)
    "ml1m": (6040, 3706, 1_000_000, "int", 3 * 365, 0),
    "ml20m": (138_493, 26_744, 20_000_000, "half", 20 * 365, 1),
    "hm": (1_371_980, 105_542, 31_000_000, "one", 2 * 365, 2),
}


def synth_probs(rng, n_users, n_items, n_events, cap=0.5):
    """(p_user, p_item) of the recipe above; consumes exactly one lognormal draw of ``rng``."""
    p_item = (np.arange(n_items) + 1.0) ** -0.9
    p_item /= p_item.sum()
    p_user = rng.lognormal(0, 1, n_users)
    p_user /= p_user.sum()
    # cap the head so that the most popular item is rated by <= cap of the users
    lim = cap * n_users / max(n_events, 1)
    for _ in range(8):
        over = p_item > lim
        if not over.any():
            break
        p_item[over] = lim
        p_item /= p_item.sum()
    return p_user, p_item


def synth_events(n_users, n_items, n_events, seed=0, rating="int", dup_frac=0.0, span_days=365.0, cap=0.5):
    rng = np.random.default_rng(seed)
    p_user, p_item = synth_probs(rng, n_users, n_items, n_events, cap)
    n_unique = int(round(n_events * (1.0 - dup_frac)))
    u_parts, i_parts, have = [], [], 0
    seen = None
    while have < n_unique:
        n_draw = int((n_unique - have) * 1.5) + 1024
        u = rng.choice(n_users, size=n_draw, p=p_user).astype(np.int64)
        i = rng.choice(n_items, size=n_draw, p=p_item).astype(np.int64)
        key = u * n_items + i
        _, first = np.unique(key, return_index=True)
        first.sort()
        key = key[first]
        if seen is not None:
            keep = ~np.isin(key, seen)
            first, key = first[keep], key[keep]
        seen = key if seen is None else np.concatenate([seen, key])
        u_parts.append(u[first]); i_parts.append(i[first]); have += len(first)
    u = np.concatenate(u_parts)[:n_unique]
    i = np.concatenate(i_parts)[:n_unique]
    n_dup = n_events - n_unique
    if n_dup > 0:
        d = rng.integers(0, n_unique, n_dup)
        u = np.concatenate([u, u[d]]); i = np.concatenate([i, i[d]])
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
    return u, i, ts, r


def synth_shape(name: str, scale: float = 1.0, **over):
    U, I, N, kind, span, seed = SHAPES[name]
    U, I, N = max(8, int(U * scale)), max(8, int(I * scale)), max(64, int(N * scale * scale))
    kw = dict(seed=seed, rating=kind, span_days=span, dup_frac=0.15 if name == "hm" else 0.0)
    kw.update(over)
    return synth_events(U, I, N, **kw)


def synth_stream(name: str, base_u, base_i, n_batches: int, batch_events: int, seed: int = 3, frac_existing: float = 0.8):
    """Event batches that FOLLOW the base events ``(base_u, base_i)`` of shape ``name`` in time (SURVEY.md 8d, C4 =
    BASELINE configs[3], streaming partial fit): ``frac_existing`` of a batch re-rates (user, item) pairs drawn uniformly
    from the base events, the rest are pairs drawn from the same user-activity / item-popularity laws as the base
    (mostly new pairs); pairs may repeat inside a batch (``update_interaction=True``: the last event wins); batch b
    covers day b after the end of the base span, timestamps ascending."""
    U, I, N, kind, span, base_seed = SHAPES[name]
    base_u, base_i = np.asarray(base_u), np.asarray(base_i)
    U, I = int(base_u.max()) + 1, int(base_i.max()) + 1
    p_user, p_item = synth_probs(np.random.default_rng(base_seed), U, I, len(base_u))
    t_end = 1.0e9 + span * 86400
    n_old = int(round(batch_events * frac_existing))
    out = []
    for b in range(n_batches):
        rng = np.random.default_rng(seed + 1000 * b)
        pick = rng.integers(0, len(base_u), n_old)
        u = np.concatenate([base_u[pick], rng.choice(U, size=batch_events - n_old, p=p_user)]).astype(np.int64)
        i = np.concatenate([base_i[pick], rng.choice(I, size=batch_events - n_old, p=p_item)]).astype(np.int64)
        perm = rng.permutation(batch_events)
        u, i = u[perm], i[perm]
        ts = t_end + b * 86400 + np.sort(rng.integers(0, 86400, batch_events)).astype(np.float64)
        if kind == "half":
            r = rng.integers(1, 11, batch_events).astype(np.float64) * 0.5
        elif kind == "int":
            r = rng.integers(1, 6, batch_events).astype(np.float64)
        elif kind == "one":
            r = np.ones(batch_events)
        else:
            r = rng.uniform(0.5, 5.0, batch_events)
        out.append((u, i, ts, r))
    return out
