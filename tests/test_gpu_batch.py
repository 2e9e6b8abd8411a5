"""GPU tests of the C batch API (wavecu_batch_*; BASELINE config 5): scans matched
against a map that is uploaded and indexed once per GPU must give exactly the results of matching each
scan on its own handle with its own copy of the target - and of the oracle; per-scan targets
(MultiMatcher::insert(id, src, target)) and the information matrix (estimateInfo's fall-through to
estimateLUMold) go through the same entry point."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W():
    import libwave_b200 as W
    return W


def test_batch_shared_map_equals_individual_matches_and_oracle(W, oracle):
    from libwave_b200 import batch, synth
    ids = [0, 1, 2, 3, 4, 5, 6]
    sources, target = synth.scan_batch(10_000, 0, ids=ids)
    sb = batch.ScanBatch(W.ICPMatcherParams(res=-1), devices=[0], workers_per_device=3)
    sb.set_map(target)
    recs = sb.match([synth.to_xyzw(s) for s in sources], scan_ids=ids, with_info=True)
    for k, src in zip(ids, sources):
        r = recs[ids.index(k)]
        assert r.scan_id == k and r.device == 0
        m = W.ICPMatcher(W.ICPMatcherParams(res=-1))
        m.setup(src, target)
        ok = m.match()
        m.estimateInfo()
        T = np.frombuffer(r.T, dtype=np.float64).reshape(4, 4)
        assert bool(r.converged) == ok and r.iterations == m.iterations
        assert np.array_equal(T, m.getResult())
        assert np.array_equal(np.frombuffer(r.info, dtype=np.float64).reshape(6, 6), m.getInfo())
        ref = oracle.icp_align(src, target, sum_mode=oracle.SUM_EXACT)
        assert r.iterations == ref.iterations and np.array_equal(T.astype(np.float32), ref.T)
    # the same batch object again (handles, streams and the shared index are reused), in another order
    recs2 = sb.match([synth.to_xyzw(s) for s in sources[::-1]], scan_ids=ids[::-1])
    for a in recs2:
        b = recs[ids.index(a.scan_id)]
        assert np.array_equal(np.frombuffer(a.T, dtype=np.float64), np.frombuffer(b.T, dtype=np.float64))
    table = batch.records_to_table(sb.allgather(recs, len(ids)), len(ids))   # one rank: a copy
    assert np.isfinite(table[:, 16]).all() and (table[:, 17] > 0).all()


def test_batch_per_scan_targets(W):
    from libwave_b200 import batch, synth
    pairs = [synth.scan_pair(10_000, scan_id=k) for k in (1, 2, 3)]
    sb = batch.ScanBatch(W.ICPMatcherParams(res=-1), devices=[0], workers_per_device=2)
    recs = sb.match([synth.to_xyzw(p[0]) for p in pairs], targets=[synth.to_xyzw(p[1]) for p in pairs])
    for r, (src, tgt) in zip(recs, pairs):
        m = W.ICPMatcher(W.ICPMatcherParams(res=-1))
        m.setup(src, tgt)
        assert m.match() == bool(r.converged)
        assert np.array_equal(np.frombuffer(r.T, dtype=np.float64).reshape(4, 4), m.getResult())


def test_shared_target_needs_a_built_owner(W):
    from libwave_b200 import capi
    import ctypes as C
    L = capi.lib()
    a, b = W.ICPMatcher(W.ICPMatcherParams(res=-1)), W.ICPMatcher(W.ICPMatcherParams(res=-1))
    assert L.wavecu_icp_share_target(b._h, a._h) == -1          # owner's target was never built
    assert L.wavecu_icp_share_target(b._h, b._h) == -1
    assert L.wavecu_icp_share_target(b._h, None) == 0
