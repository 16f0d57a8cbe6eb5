"""Property-based tests (hypothesis) of the host-side machinery every GPU path depends on: the piece planner, the
prefix-caching page allocator and the scheduler.  The reference has no equivalent (its tests are example-based,
SURVEY 4); these guard the invariants the device code silently assumes -- contiguous piece cover, pages never owned
twice unless shared through the prefix cache, every slot and page returned whatever the arrival / cancel pattern."""
import random

import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

from bee2bee_b200.engine.core import Engine, SamplingParams  # noqa: E402
from bee2bee_b200.engine.kv import PAGE, OutOfPages, PageAllocator  # noqa: E402
from bee2bee_b200.models.config import UNITS_PER_LAYER, ModelConfig, piece_units, resolve_config, unit_layers  # noqa: E402


# ------------------------------------------------------------------------------------------------ planner
@settings(max_examples=60, deadline=None, derandomize=True)
@given(layers=st.integers(2, 48), pieces=st.integers(1, 16), hidden=st.sampled_from([256, 1024, 4096]),
       ffn_mult=st.sampled_from([2, 3, 4]), vocab=st.sampled_from([1024, 32000, 128256]))
def test_piece_planner_covers_the_model_contiguously(layers, pieces, hidden, ffn_mult, vocab):
    base = resolve_config("tiny-llama")
    cfg = ModelConfig(**{**base.__dict__, "name": "prop", "n_layers": layers, "hidden_size": hidden,
                         "ffn_size": hidden * ffn_mult, "vocab_size": vocab, "n_heads": hidden // 64,
                         "n_kv_heads": max(1, hidden // 256), "head_dim": 64})
    ranges = piece_units(cfg, pieces)
    U = UNITS_PER_LAYER * layers
    assert 1 <= len(ranges) <= min(pieces, layers)
    assert ranges[0][0] == 0 and ranges[-1][1] == U
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])), ranges        # contiguous, no gap, no overlap
    assert all(u1 - u0 >= (2 if len(ranges) > 1 else 1) for u0, u1 in ranges), ranges   # no lone GEMM piece
    # the layers a piece touches follow from its units; together they cover every layer
    seen = set()
    for u0, u1 in ranges:
        ls = list(unit_layers((u0, u1)))
        assert ls == list(range(u0 // UNITS_PER_LAYER, (u1 - 1) // UNITS_PER_LAYER + 1))
        seen.update(ls)
    assert seen == set(range(layers))
    # the last piece carries the lm_head: with a big vocabulary it must not also be the longest in units
    if len(ranges) >= 4 and vocab >= 32000 and hidden >= 1024:
        assert (ranges[-1][1] - ranges[-1][0]) <= max(u1 - u0 for u0, u1 in ranges[:-1])


# ---------------------------------------------------------------------------------------------- allocator
def _check_allocator(a: PageAllocator):
    owned = [p for ps in a._owned.values() for p in ps]
    # reference counts equal the number of owners of every page
    for p in set(owned):
        assert a._ref[p] == owned.count(p), (p, a._ref[p], owned.count(p))
    assert set(a._ref) == set(owned)
    free, lru = set(a._free), set(a._lru)
    assert len(free) == len(a._free), "a page is on the free list twice"
    assert not (free & lru) and not (free & set(owned)) and not (lru & set(owned))
    assert 0 not in free | lru | set(owned), "page 0 is the reserved null page"
    assert free | lru | set(owned) == set(range(1, a.num_pages)), "a page leaked"
    # a page shared by several owners is always a committed prefix page
    for p in set(owned):
        if owned.count(p) > 1:
            assert p in a._key_of
    # the key maps are inverse of each other
    assert {k: p for p, k in a._key_of.items()} == a._page_of


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 2 ** 31 - 1), num_pages=st.integers(4, 40), n_ops=st.integers(5, 120))
def test_page_allocator_invariants_under_random_traffic(seed, num_pages, n_ops):
    rng = random.Random(seed)
    a = PageAllocator(num_pages, prefix_cache=True)
    families = [[rng.randrange(1000) for _ in range(PAGE * 3)] for _ in range(3)]     # shared prompt prefixes
    live = {}
    next_owner = 0
    for _ in range(n_ops):
        op = rng.random()
        if op < 0.55 or not live:
            fam = rng.choice(families)
            prompt = fam[:rng.randrange(1, len(fam))] + [rng.randrange(1000) for _ in range(rng.randrange(0, 40))]
            need = len(prompt) + rng.randrange(1, 80)
            ok = a.can_allocate(need, prompt)
            if ok:
                pages = a.allocate(next_owner, need, prompt)
                assert len(pages) == a.pages_for(need)
                assert a.cached_tokens(next_owner) % PAGE == 0 and a.cached_tokens(next_owner) < len(prompt)
                if rng.random() < 0.9:
                    a.commit(next_owner, prompt)
                live[next_owner] = prompt
                next_owner += 1
            else:
                with pytest.raises(OutOfPages):
                    a.allocate(next_owner, need, prompt)
        elif op < 0.9:
            o = rng.choice(list(live))
            a.release(o)
            del live[o]
        else:
            o = rng.choice(list(live))
            a.invalidate(o)               # failed prefill: keys of exclusively owned pages are forgotten
        _check_allocator(a)
    for o in list(live):
        a.release(o)
    _check_allocator(a)
    assert a.free_pages == num_pages - 1


def test_prefix_cache_never_serves_a_page_whose_content_was_evicted():
    a = PageAllocator(4, prefix_cache=True)                      # 3 usable pages
    p1 = list(range(PAGE + 5))
    a.allocate(0, len(p1) + 1, p1)
    a.commit(0, p1)
    a.release(0)                                                  # page stays resident as an evictable cache entry
    assert a.cached_tokens(0) == 0
    a.allocate(1, len(p1) + 1, p1)
    assert a.cached_tokens(1) == PAGE                             # hit
    a.release(1)
    other = [7] * (PAGE * 3 - 1)
    a.allocate(2, len(other), other)                              # needs all 3 pages: evicts the cached one
    a.release(2)
    a.allocate(3, len(p1) + 1, p1)
    assert a.cached_tokens(3) == 0                                # the key went away with the page


# ---------------------------------------------------------------------------------------------- scheduler
@settings(max_examples=12, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 2 ** 31 - 1), max_batch=st.sampled_from([1, 2, 4]), burst=st.sampled_from([1, 3, 8]))
def test_scheduler_returns_every_slot_and_page_under_random_arrivals_and_cancels(seed, max_batch, burst):
    rng = random.Random(seed)
    eng = Engine("tiny-llama", device="cpu", max_batch=max_batch, max_seq_len=128, decode_burst=burst)
    reqs, cancelled = [], set()
    for _ in range(rng.randrange(3, 9)):
        for _ in range(rng.randrange(0, 3)):
            prompt = [rng.randrange(4, 200) for _ in range(rng.randrange(1, 70))]
            sp = SamplingParams(max_new_tokens=rng.randrange(1, 12), temperature=0.0, ignore_eos=True)
            reqs.append((eng.submit(prompt, sp), sp))
        if reqs and rng.random() < 0.35:
            i = rng.randrange(len(reqs))
            eng.cancel(reqs[i][0])
            cancelled.add(i)
        eng.step()
    guard = 0
    while not all(r.done.is_set() for r, _ in reqs):
        eng.step()
        guard += 1
        assert guard < 500, "scheduler does not drain"
    for i, (r, sp) in enumerate(reqs):
        if r.finish_reason == "cancelled":
            assert i in cancelled and len(r.out_ids) <= sp.max_new_tokens
        else:
            assert r.error is None and len(r.out_ids) == sp.max_new_tokens, (r.finish_reason, r.error)
    eng.step()
    assert not eng._running and not eng._pending
    assert len(eng._free_slots) == max_batch and eng.alloc.free_pages == eng.alloc.num_pages - 1
