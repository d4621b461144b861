"""Randomised agreement of the pruned search with the exhaustive one (and the float64
oracle) over cloud shapes the hand-picked cases do not cover: lines, planes, duplicates,
clusters, wildly different sizes, far-apart clouds, several ICP starts with hints."""
import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _cloud(rng, kind, n):
    if kind == "gauss":
        return rng.normal(scale=rng.uniform(0.5, 80), size=(n, 3))
    if kind == "line":
        t = rng.uniform(-100, 100, size=(n, 1))
        return t * rng.normal(size=(1, 3)) + rng.normal(scale=0.01, size=(n, 3))
    if kind == "plane":
        uv = rng.uniform(-60, 60, size=(n, 2))
        return np.concatenate([uv, rng.normal(scale=0.05, size=(n, 1))], axis=1)
    if kind == "dups":
        base = rng.normal(scale=30, size=(max(1, n // 7), 3))
        return base[rng.integers(0, len(base), n)]
    if kind == "clusters":
        cen = rng.normal(scale=150, size=(25, 3))
        return cen[rng.integers(0, 25, n)] + rng.normal(scale=rng.uniform(0.1, 8), size=(n, 3))
    if kind == "shell":
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        return v * np.array([60.0, 42.0, 30.0]) + rng.normal(scale=0.3, size=(n, 3))
    raise ValueError(kind)


KINDS = ["gauss", "line", "plane", "dups", "clusters", "shell"]


@pytest.mark.parametrize("seed", range(24))
def test_pruned_search_equals_exhaustive_on_random_shapes(gpu, seed):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api
    rng = np.random.default_rng(1000 + seed)
    nq = int(rng.choice([1, 31, 255, 256, 257, 1500, 9000, 40000]))
    nt = int(rng.choice([1, 63, 64, 65, 1023, 1025, 5000, 70000, 150000]))
    kq, kt = KINDS[seed % 6], KINDS[(seed // 6 + seed) % 6]
    q = _cloud(rng, kq, nq)
    t = _cloud(rng, kt, nt)
    if seed % 5 == 0:
        q = q + rng.normal(scale=400, size=(1, 3))      # far apart: nothing can be pruned early
    if seed % 7 == 0 and nq <= nt:
        q = t[rng.integers(0, nt, nq)]                    # queries ARE target points: distance 0
    off = rng.normal(scale=[0.0, 700.0][seed % 2], size=(1, 3))
    q, t = (q + off).astype(np.float32), (t + off).astype(np.float32)
    try:
        pr = gpu.nearest_neighbors(q, t)
        api.set_nn_pruning(False)
        ex = gpu.nearest_neighbors(q, t)
    finally:
        api.set_nn_pruning(True)
    np.testing.assert_array_equal(pr.idx.cpu().numpy(), ex.idx.cpu().numpy())
    np.testing.assert_array_equal(pr.d2.cpu().numpy().view(np.uint32), ex.d2.cpu().numpy().view(np.uint32))
    sub = rng.choice(nq, size=min(nq, 1500), replace=False)
    _, ridx = c_oracle.nn_f64(q[sub], t)
    np.testing.assert_array_equal(pr.idx.cpu().numpy()[sub], ridx)


@pytest.mark.parametrize("seed", range(4))
def test_multistart_icp_with_hints_equals_exhaustive(gpu, seed):
    """Hints carry over between iterations and between very different starts: every start of a
    batch must still reproduce the exhaustive search bit for bit."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth
    rng = np.random.default_rng(50 + seed)
    ns, nt = int(rng.choice([900, 5000, 12000])), int(rng.choice([1100, 7000, 20000]))
    src, tgt, _ = synth.icp_pair(ns, nt, 10 + seed, 20 + seed)
    inits = np.tile(np.eye(4), (5, 1, 1))
    for k in range(1, 5):
        inits[k, :3, :3] = synth.random_rotation(rng) if k == 4 else inits[k, :3, :3]
        inits[k, :3, 3] = rng.normal(scale=[1.0, 5.0, 30.0, 80.0][k - 1], size=3)
    out = {}
    try:
        for on in (True, False):
            api.set_nn_pruning(on)
            m = gpu.multistart_icp(src, tgt, inits, 20.0, max_iteration=7)
            out[on] = ([r.transformation for r in m.results], [r.fitness for r in m.results],
                       [r.inlier_rmse for r in m.results], [r.iterations for r in m.results], m.chamfer)
    finally:
        api.set_nn_pruning(True)
    for a, b in zip(out[True], out[False]):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
