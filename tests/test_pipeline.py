"""Host logic of the chunked forward -> walk pipeline (sls_sat_solving_with_deep_learning_b200/pipeline.py): ordering, chunk
boundaries, seeds and failure handling of the staged executor -- no device needed; the CUDA run of the same function is
tests/test_gpu_gnn.py::test_pipeline_equals_the_monolithic_calls."""
import threading
import time
from types import SimpleNamespace

import numpy as np
import pytest

from sls_sat_solving_with_deep_learning_b200 import pipeline
from sls_sat_solving_with_deep_learning_b200.solver import PanicException


def test_staged_keeps_item_order_and_overlaps_stages():
    active = {"n": 0, "max": 0}
    lock = threading.Lock()

    def stage(tag):
        def f(x):
            with lock:
                active["n"] += 1
                active["max"] = max(active["max"], active["n"])
            time.sleep(0.01)
            with lock:
                active["n"] -= 1
            return x + [tag]
        return f

    out = pipeline.staged([stage("a"), stage("b"), stage("c")], [[i] for i in range(12)])
    assert out == [[i, "a", "b", "c"] for i in range(12)]
    assert active["max"] >= 2  # different items were in different stages at the same time


def test_staged_multi_worker_stage_hands_on_in_item_order():
    seen = []

    def slow_or_fast(tag):
        def f(x):
            time.sleep(0.02 if (x % 2 == 0) == (tag == "a") else 0.001)  # the two workers finish out of order
            return (x, tag)
        return f

    def last(y):
        seen.append(y[0])
        return y

    for n in (0, 1, 2, 7, 20):
        del seen[:]
        out = pipeline.staged([lambda x: x, [slow_or_fast("a"), slow_or_fast("b")], last], range(n), depth=2)
        assert [o[0] for o in out] == list(range(n)) and seen == list(range(n))
        if n >= 7:
            assert {o[1] for o in out} == {"a", "b"}  # both workers took part
    with pytest.raises(ValueError):
        pipeline.staged([lambda x: x, [slow_or_fast("a"), lambda x: (_ for _ in ()).throw(ValueError("boom"))], last], range(30))


def test_staged_without_items_and_with_one_stage():
    assert pipeline.staged([lambda x: x], []) == []
    assert pipeline.staged([lambda x: 2 * x], [1, 2, 3], depth=1) == [2, 4, 6]


@pytest.mark.parametrize("exc", [ValueError("bad chunk"), PanicException(4, "weights shorter than var_count")])
def test_staged_reraises_the_first_failure_and_terminates(exc):
    def middle(x):
        if x == 3:
            raise exc
        return x

    with pytest.raises(type(exc)):
        pipeline.staged([lambda x: x, middle, lambda x: x], range(50), depth=1)


def test_solve_with_oracle_chunks_seeds_and_order(monkeypatch):
    calls = []

    class FakeNet:
        last_kernel_ms = 1.0

        def probabilities(self, insts):
            return [np.full(i.n, 0.25 + 0.001 * i.tag) for i in insts]

    def fake_batch(insts, algo, wi, wr, nsteps, nruns, mapping, pfb, instance_seeds=None):
        assert all(a is b for a, b in zip(wi, wr))
        calls.append((tuple(i.tag for i in insts), tuple(instance_seeds)))
        return [SimpleNamespace(tag=i.tag, seed=s, p0=float(w[0]), kernel_ms=2.0) for i, s, w in zip(insts, instance_seeds, wi)]

    monkeypatch.setattr(pipeline._solver, "run_sls_batch", fake_batch)
    monkeypatch.setattr(pipeline._solver, "set_device", lambda d: None)
    reserved = []
    monkeypatch.setattr(pipeline._solver, "reserve_device_pool", reserved.append)
    stats = {}
    res, probs = pipeline.solve_with_oracle(FakeNet(), list(range(10)), "moser", 100, 4, chunk=4,
                                            make_instance=lambda t: SimpleNamespace(tag=t, n=3 + t, m=12, num_lits=36),
                                            instance_seeds=[100 + t for t in range(10)], stats=stats)
    assert [r.tag for r in res] == list(range(10))
    assert [r.seed for r in res] == [100 + t for t in range(10)]
    assert all(abs(r.p0 - (0.25 + 0.001 * r.tag)) < 1e-12 for r in res)
    assert [len(p) for p in probs] == [3 + t for t in range(10)]
    assert calls == [((0, 1, 2, 3), (100, 101, 102, 103)), ((4, 5, 6, 7), (104, 105, 106, 107)), ((8, 9), (108, 109))]
    assert stats["walk_kernel_ms"] == 6.0 and stats["forward_device_ms"] == 3.0
    assert len(reserved) == 1 and reserved[0] > 10 * (24 * 36 + 16 * 12)  # the pool's working set is mapped once, up front
    del calls[:]
    res2, _ = pipeline.solve_with_oracle([FakeNet(), FakeNet()], list(range(10)), "moser", 100, 4, chunk=3,
                                         make_instance=lambda t: SimpleNamespace(tag=t, n=3 + t, m=12, num_lits=36),
                                         instance_seeds=[100 + t for t in range(10)], reserve=False)
    assert [r.tag for r in res2] == list(range(10)) and [c[0] for c in calls] == [(0, 1, 2), (3, 4, 5), (6, 7, 8), (9,)]
    with pytest.raises(ValueError):
        pipeline.solve_with_oracle(FakeNet(), [1, 2], "moser", 1, 1)  # not handles and no make_instance
    with pytest.raises(ValueError):
        pipeline.solve_with_oracle(FakeNet(), [1, 2], "moser", 1, 1, make_instance=lambda t: t, instance_seeds=[1])
