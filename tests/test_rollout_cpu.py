"""CPU tests of the rollout shim's host-side maths (normaliser with norm groups, policy head) against numpy restatements of
R/learning/normalizer.py and the actor construction of R/learning/pg_agent.py:140-160."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from deepmimic_b200.rollout import DeviceNormalizer, build_policy


def _np_normalizer_update(mean, mean_sq, count, x, groups, eps=0.02):
    new_mean, new_mean_sq = x.mean(0), (x * x).mean(0)

    def proc(new, old):
        out = new.copy()
        for g in np.unique(groups):
            idx = np.nonzero(groups == g)[0]
            if g == -1:
                out[idx] = old[idx]
            elif g != 0:
                out[idx] = new[idx].mean()
        return out
    new_mean, new_mean_sq = proc(new_mean, mean), proc(new_mean_sq, mean_sq)
    tot = count + x.shape[0]
    mean = count / tot * mean + x.shape[0] / tot * new_mean
    mean_sq = count / tot * mean_sq + x.shape[0] / tot * new_mean_sq
    std = np.maximum(np.sqrt(np.maximum(mean_sq - mean * mean, 0)), eps)
    return mean, mean_sq, tot, std


def test_normalizer_matches_reference_semantics():
    rng = np.random.default_rng(0)
    size = 12
    groups = np.array([-1, 0, 0, 0, 1, 1, 1, 2, 2, 0, 0, 0])
    n = DeviceNormalizer(size, groups)
    mean0, std0 = rng.normal(size=size), rng.uniform(0.5, 2, size=size)
    n.set_mean_std(mean0, std0)
    mean, mean_sq, count = mean0.copy(), std0 ** 2 + mean0 ** 2, 0
    for _ in range(3):
        x = rng.normal(1.0, 2.0, size=(50, size))
        n.record(torch.tensor(x, dtype=torch.float32))
        n.update()
        mean, mean_sq, count, std = _np_normalizer_update(mean, mean_sq, count, x, groups)
        np.testing.assert_allclose(n.mean.numpy(), mean, rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(n.std.numpy(), std, rtol=2e-4, atol=2e-5)
    assert n.mean[0].item() == pytest.approx(mean0[0])          # NORM_GROUP_NONE keeps its statistics (the phase slot)
    assert np.allclose(n.mean.numpy()[4:7], n.mean.numpy()[4])   # a shared group gets one mean
    x = torch.tensor(rng.normal(size=(5, size)), dtype=torch.float32)
    torch.testing.assert_close(n.unnormalize(n.normalize(x)), x, rtol=1e-5, atol=1e-5)


def test_policy_head_shapes_and_logp():
    torch.manual_seed(0)
    p = build_policy(227, 28, init_output_scale=0.01, noise=0.05)
    assert [l.weight.shape for l in p.hidden] == [(1024, 227), (512, 1024)] and p.mean.weight.shape == (28, 512)
    assert float(p.mean.weight.abs().max()) <= 0.01 and float(p.logstd[0]) == pytest.approx(math.log(0.05))
    s = torch.randn(7, 227)
    mask = torch.tensor([1, 0, 1, 1, 0, 1, 1], dtype=torch.bool)
    a, logp = p.sample(s, mask)
    mu = p(s)
    torch.testing.assert_close(a[~mask], mu[~mask])               # non-exploring rows take the mode
    z = (a - mu) / p.logstd.exp()
    ref = (-0.5 * z * z - p.logstd - 0.5 * math.log(2 * math.pi)).sum(-1)
    torch.testing.assert_close(logp, ref, rtol=1e-4, atol=1e-4)
