"""CPU tests of the host-side mirror: bookkeeping that must be bit-exact with the reference
(Identifier, LRUFreqSet, metrics) and the C-ABI surface (library loads, every declared symbol is
exported).  No compute call is made here (no GPU in this container)."""
import ctypes
import os
import re

import numpy as np
import pytest

from rtrec_b200 import _lib
from rtrec_b200.utils.identifiers import Identifier
from rtrec_b200.utils.lru import LRUFreqSet
from rtrec_b200.utils import metrics as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rtrec_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/rtrec_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.rt_version() >= 100


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rtrec_b200.models import SLIM
    m = SLIM()
    assert m.recommend("nobody", top_k=3) == []  # empty model: host-only answer
    with pytest.raises(_lib.RtrecB200Error):
        m.fit([("u", "i", 1.0, 1.0)])


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rtrec_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f"{f} mentions oracle"


def test_identifier_semantics():
    a = Identifier()
    assert a.identify(5) == 5 and a.pass_through is True
    with pytest.raises(ValueError):
        a.identify("x")
    assert a.get_id(7) == 7 and a.get(9) == 9
    b = Identifier()
    assert [b.identify(x) for x in ("p", "q", "p")] == [0, 1, 0]
    assert b.pass_through is False
    with pytest.raises(ValueError):
        b.identify(3)
    with pytest.raises(ValueError):
        b.get_id(3)
    assert b.get_id("zz") is None and b.get(1) == "q"
    c = Identifier(force_identify=True)
    assert c.identify(100) == 0 and c.identify(7) == 1 and c.get_id(100) == 0 and c.get(1) == 7
    # vectorised path == loop
    d1, d2 = Identifier(), Identifier()
    col = np.array(["a", "b", "a", "c", "b"], dtype=object)
    assert d1.identify_many(col).tolist() == [d2.identify(x) for x in col]
    e = Identifier()
    assert e.identify_many(np.array([4, 2, 4])).tolist() == [4, 2, 4] and e.pass_through is True


@pytest.mark.parametrize("cap", [1000, 7])
def test_lru_add_batch_equals_loop(cap):
    rng = np.random.default_rng(0)
    a, b = LRUFreqSet(cap), LRUFreqSet(cap)
    for _ in range(5):
        vals = rng.integers(0, 30, 200)
        a.add_batch(vals)
        for v in vals.tolist():
            b.add(v)
        assert list(a.data.items()) == list(b.data.items())
        assert list(a.get_freq_items(5)) == list(b.get_freq_items(5))
    assert list(a.get_freq_items(3, exclude_items=[list(a.data)[0]])) == list(b.get_freq_items(3, exclude_items=[list(b.data)[0]]))


def test_metrics_known_answers():
    # known answers of /root/reference/tests/utils/test_metrics.py (rel 1e-4)
    r, g = [1, 3, 2, 6], [1, 2, 4]
    assert M.precision(r, g, 4) == pytest.approx(0.5)
    assert M.recall(r, g, 4) == pytest.approx(2 / 3)
    assert M.f1_score(r, g, 4) == pytest.approx(2 * 0.5 * (2 / 3) / (0.5 + 2 / 3))
    assert M.hit(r, g, 2) == 1.0 and M.hit([9], g, 1) == 0.0
    assert M.reciprocal_rank([3, 1], g, 2) == pytest.approx(0.5)
    assert M.ndcg(r, g, 4) == pytest.approx((1 + 1 / np.log2(4)) / (1 + 1 / np.log2(3) + 1 / np.log2(4)))
    assert M.average_precision(r, g, 4) == pytest.approx((1 + 2 / 3) / 3)
    assert M.auc(r, g, 4) == pytest.approx(0.75)  # TP,FP,TP,FP: 3 of 4 (TP,FP) pairs ordered
    assert M.precision([], [], 5) == 1.0 and M.recall([1], [], 5) == 0.0 and M.auc([], [1], 5) == 0.0
    out = M.compute_scores([(r, g), ([], [])], 4)
    assert list(out) == ["precision", "recall", "f1", "ndcg", "hit_rate", "mrr", "map", "tp", "auc"]
    assert out["tp"] == 2
    assert M.compute_scores([], 3)["ndcg"] == 0.0


def test_every_referenced_helper_exists():
    """Every ``D.<name>`` / ``P.<name>`` the package, the bench, the tools and the tests refer to exists, and
    every ``rt_*`` entry point called through ctypes has a prototype (a missing wrapper would otherwise only
    show up on the GPU box)."""
    import glob
    import re
    from rtrec_b200 import _lib, device as D, pipeline as P
    files = (glob.glob(os.path.join(ROOT, "rtrec_b200", "**", "*.py"), recursive=True)
             + glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "*.py"))
             + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
    missing = []
    for f in files:
        src = open(f).read()
        for mod, obj, marker in (("D", D, "device as D"), ("P", P, "pipeline as P")):
            if marker not in src:
                continue
            for name in set(re.findall(r"(?<![\w.])" + mod + r"\.([A-Za-z_]\w*)", src)):
                if not hasattr(obj, name):
                    missing.append((os.path.relpath(f, ROOT), f"{mod}.{name}"))
        for a, b in set(re.findall(r"\blib\.(rt_\w+)|load\(\)\.(rt_\w+)", src)):
            if (a or b) not in _lib.PROTOTYPES:
                missing.append((os.path.relpath(f, ROOT), a or b))
    assert not missing, missing


@pytest.mark.parametrize("n_items,n_parts", [(1, 1), (63, 2), (64, 2), (65, 8), (1777, 3), (26744, 8), (105542, 5)])
def test_block_cyclic_row_ownership(n_items, n_parts):
    """Geometry of the owner-rows multi-GPU fit (rt_gram_block_rows; pure host code, no GPU): the 64-row blocks of
    the rank-space matrix are dealt round-robin, every part's buffer has the same height, the real rows of all parts
    add up to the catalogue, and the local-row -> global-row map used for the target lists is a bijection."""
    from rtrec_b200 import device as D
    nt = (n_items + 63) // 64
    seen = []
    for part in range(n_parts):
        rows_alloc, rows_own = D.gram_block_rows(n_items, n_parts, part)
        assert rows_alloc == -(-nt // n_parts) * 64 and rows_own <= rows_alloc
        l = np.arange(rows_own, dtype=np.int64)
        jp = ((l >> 6) * n_parts + part) * 64 + (l & 63)
        assert (jp < n_items).all()
        assert (((jp >> 6) % n_parts) == part).all()
        assert np.array_equal((((jp >> 6) // n_parts) << 6) | (jp & 63), l)   # blk_local_row
        seen.append(jp)
    allj = np.concatenate(seen)
    assert np.array_equal(np.sort(allj), np.arange(n_items))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys, on
    the smallest workload.  This is the one place outside tests/ that may execute oracle/ (bench.py docstring)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "ml1m",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "users/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(line["config"]) >= {"workload", "n_users", "n_items", "n_events"} and "model" not in line["config"]


def test_lru_native_replay_equals_python_loop():
    """rt_lru_replay (host code of the library) == LRUFreqSet.add per event, with evictions inside the batch
    (/root/reference/rtrec/utils/lru.py:33-47)."""
    import numpy as np
    from rtrec_b200.utils.lru import LRUFreqSet
    rng = np.random.default_rng(0)
    vals = rng.zipf(1.3, 60000) % 3000
    a, b = LRUFreqSet(700), LRUFreqSet(700)
    for v in vals[:5000].tolist():
        a.add(v); b.add(v)
    assert a._replay_native(vals[5000:])
    for v in vals[5000:].tolist():
        b.add(v)
    assert list(a.data.items()) == list(b.data.items())
    assert list(a.get_freq_items(20)) == list(b.get_freq_items(20))
    c = LRUFreqSet(10)
    c.add("x")
    assert not c._replay_native(np.arange(5000))      # non-integer key in the set: the Python loop handles it
