"""Host logic of the engine's plan cache (no GPU): least-recently-used order, eviction clears the captured graphs before the
workspace goes, and an evicted plan drops its buffers at once instead of waiting for the cyclic garbage collector."""
import gc
import weakref

import torch

from rvc_maker_b200 import engine


def _stub_engine(max_plans=4):
    e = engine.Engine.__new__(engine.Engine)   # no device, no library: only the cache logic is exercised
    e.plans, e.graphs, e._warm, e.max_plans = {}, {}, set(), max_plans
    e.dev, e.poison = torch.device("cpu"), ""
    built = []

    def make(kind):
        def build(B, T):
            p = engine.Plan()
            p.out_arena = e._buf(p, 8)
            z = p.out_arena
            p.add_py(lambda: z.add_(1))        # closures capture tensors, never the plan itself
            built.append((kind, B, T))
            return p
        return build

    e._enc_plan, e._dec_plan = make("enc"), make("dec")
    return e, built


def test_plan_cache_is_lru_and_eviction_drops_only_the_graphs_built_on_the_evicted_plan():
    e, built = _stub_engine(max_plans=4)
    a_enc, a_dec = e.plan("enc", 1, 10), e.plan("dec", 1, 10)
    e.plan("enc", 1, 20), e.plan("dec", 1, 20)
    assert e.plan("enc", 1, 10) is a_enc and len(built) == 4          # a hit builds nothing ...
    assert list(e.plans)[-1] == ("enc", 1, 10)                         # ... and becomes the most recent entry
    # a graph records the plans it replays into (engine.infer): one on the shape about to lose its decoder plan, one on another shape
    e.graphs[(1, 10, 0)] = (object(), (("enc", 1, 10), ("dec", 1, 10)))
    e.graphs[(1, 20, 0)] = (object(), (("enc", 1, 20), ("dec", 1, 20)))
    e._warm.update({(1, 10, 0), (1, 20, 0)})
    e.plan("enc", 1, 30)                                               # evicts the least recently used: ("dec", 1, 10)
    assert ("dec", 1, 10) not in e.plans and ("enc", 1, 10) in e.plans and len(e.plans) == 4
    assert list(e.graphs) == [(1, 20, 0)] and e._warm == {(1, 20, 0)}  # only the graph that replayed into the evicted buffers goes
    assert e.plan("dec", 1, 10) is not a_dec and built[-1] == ("dec", 1, 10)


def test_poison_mode_fills_every_dtype_and_guards_every_workspace():
    e, _ = _stub_engine()
    e.poison = "nan"
    p = engine.Plan()
    f, i, z = e._buf(p, 3, 5), e._buf(p, 7, dtype=torch.int32), e._zbuf(p, 48)
    assert torch.isnan(f).all() and (i == 0x7E7E7E7E).all() and (z == 0).all()
    assert len(p.keep) == 3 and all(k.numel() >= e.GUARD and (k[-e.GUARD:] == 0x7E).all() for k in p.keep)
    assert f.shape == (3, 5) and i.shape == (7,) and z.numel() == 48


def test_evicted_plan_is_freed_without_a_gc_pass():
    e, _ = _stub_engine(max_plans=2)
    gc.collect()
    gc.disable()
    try:
        p = e.plan("enc", 1, 10)
        refs = [weakref.ref(p), weakref.ref(p.out_arena)]
        del p
        e.plan("dec", 1, 10)
        e.plan("enc", 1, 20)                                           # evicts ("enc", 1, 10)
        assert all(r() is None for r in refs)
    finally:
        gc.enable()


def test_release_drops_buffers_even_if_someone_still_holds_the_plan():
    e, _ = _stub_engine(max_plans=2)
    p = e.plan("enc", 1, 10)
    buf = weakref.ref(p.out_arena)
    gc.disable()
    try:
        p.release()
        assert buf() is None and p.ops == [] and p.keep == [] and not hasattr(p, "out_arena")
    finally:
        gc.enable()
