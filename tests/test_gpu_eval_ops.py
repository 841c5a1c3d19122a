"""GPU parity of the SURVEY.md 8(f) widening: device-side objective distortions (train.py:399-432) and
inference MLPG with real variances (evaluation_tts.py:70-72,92-94).  Oracles: the numpy restatements of
the nnmnkwii metrics / paramgen (package not vendored: parity unpinned, except that unit-variance MLPG must
agree with the golden-pinned R matrix path)."""
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import WINDOWS, ROOT
from oracle import nnmnkwii_port as nnp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__
    __graft_entry__.build()
    return torch.device("cuda:0")


def _reference_compute_distortions(y, yh, Ym, Ys, lengths, hp):
    """reference train.py:358-432 as restated in oracle/ (gantts_port.compute_distortions over nnmnkwii_port's metrics --
    the checker shares no code with the product's numpy shim compat/nnmnkwii/metrics.py)."""
    from oracle import gantts_port as gp
    return gp.compute_distortions(y, yh, Ym, Ys, lengths, hp)


@pytest.mark.parametrize("B,T", [(3, 50), (32, 1000)])
def test_compute_distortions_acoustic(dev, B, T):
    from gantts_b200 import metrics
    rng = np.random.RandomState(3)
    hp = types.SimpleNamespace(name="acoustic", windows=WINDOWS, stream_sizes=[180, 3, 1, 3],
                               has_dynamic_features=[True, True, False, True])
    D = 63
    y = rng.randn(B, T, D).astype(np.float32)
    yh = (y + 0.3 * rng.randn(B, T, D)).astype(np.float32)
    Ym = rng.randn(187) * 0.5
    Ys = 0.5 + rng.rand(187)
    Ym[183], Ys[183] = 0.5, 0.5                       # V/UV statistics around the 0.5 threshold
    Ym[180], Ys[180] = 5.0, 0.3                       # log-F0 scale
    lengths = sorted([T] + list(rng.randint(T // 2, T, B - 1)), reverse=True)
    got = metrics.compute_distortions(torch.from_numpy(y).to(dev), torch.from_numpy(yh).to(dev), Ym, Ys,
                                      lengths=lengths, hp=hp)
    want = _reference_compute_distortions(y.astype(np.float64), yh.astype(np.float64), Ym, Ys, lengths, hp)
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 2e-5 * abs(want[k]) + 1e-7, (k, got[k], want[k])


def test_compute_distortions_vc_and_duration(dev):
    from gantts_b200 import metrics
    from oracle import gantts_port as gp
    rng = np.random.RandomState(4)
    B, T = 4, 70
    lengths = [70, 66, 41, 40]
    y = rng.randn(B, T, 59).astype(np.float32)
    yh = (y + 0.2 * rng.randn(B, T, 59)).astype(np.float32)
    Ym, Ys = rng.randn(177), 0.5 + rng.rand(177)
    hp = types.SimpleNamespace(name="vc", order=59)
    got = metrics.compute_distortions(torch.from_numpy(y).to(dev), torch.from_numpy(yh).to(dev), Ym, Ys, lengths, hp)
    want = gp.compute_distortions(y, yh, Ym, Ys, lengths, hp)["mcd"]
    assert abs(got["mcd"] - want) <= 2e-5 * want
    hp = types.SimpleNamespace(name="duration")
    y5, yh5 = y[:, :, :5], yh[:, :, :5]
    got = metrics.compute_distortions(torch.from_numpy(y5).to(dev), torch.from_numpy(yh5).to(dev), Ym[:5], Ys[:5],
                                      lengths, hp)
    want = gp.compute_distortions(y5, yh5, Ym[:5], Ys[:5], lengths, hp)["dur_rmse"]
    assert abs(got["dur_rmse"] - want) <= 2e-5 * want


def test_distortions_no_voiced_frames_is_nan(dev):
    from gantts_b200 import metrics
    hp = types.SimpleNamespace(name="acoustic", windows=WINDOWS, stream_sizes=[180, 3, 1, 3],
                               has_dynamic_features=[True, True, False, True])
    y = torch.zeros(2, 10, 63, device=dev)
    Ym, Ys = np.zeros(187), np.ones(187)              # V/UV = 0 everywhere -> no voiced frame
    got = metrics.compute_distortions(y, y + 1.0, Ym, Ys, [10, 7], hp)
    assert math.isnan(got["f0_rmse"]) and got["vuv_err"] == 1.0   # 0 vs 1 after the shift


@pytest.mark.parametrize("T,sd", [(1, 2), (2, 3), (7, 1), (120, 60), (1000, 3)])
def test_mlpg_var_vs_dense_f64(dev, T, sd):
    from gantts_b200 import ops
    rng = np.random.RandomState(T + sd)
    mu = rng.randn(T, 3 * sd).astype(np.float32)
    for var in (0.2 + rng.rand(3 * sd).astype(np.float32), 0.2 + rng.rand(T, 3 * sd).astype(np.float32)):
        want = nnp.mlpg(mu, var, WINDOWS)
        got = ops.mlpg_var(torch.from_numpy(mu).to(dev), torch.from_numpy(var).to(dev),
                           ops.windows_key(WINDOWS)).cpu().numpy()
        assert got.shape == (T, sd)
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6 * np.abs(want).max())


def test_mlpg_var_unit_variance_equals_R_path(dev):
    """Ties the unpinned variance solver to the golden-pinned R matrix arithmetic."""
    from gantts_b200 import ops
    rng = np.random.RandomState(0)
    T, sd = 200, 4
    mu = rng.randn(2, T, 3 * sd).astype(np.float32)
    R = nnp.unit_variance_mlpg_matrix(WINDOWS, T).astype(np.float64)
    got = ops.mlpg_var(torch.from_numpy(mu).to(dev), torch.ones(3 * sd, device=dev), ops.windows_key(WINDOWS))
    for b in range(2):
        wm = np.concatenate([mu[b, :, w * sd:(w + 1) * sd] for w in range(3)], axis=0).astype(np.float64)
        np.testing.assert_allclose(got[b].cpu().numpy(), R @ wm, rtol=0, atol=3e-6 * np.abs(R @ wm).max())


def test_mlpg_var_wide_windows_and_batch(dev):
    """5-tap windows (half bandwidth 4, the widest the ABI admits), per-frame variances, batch of 3."""
    from gantts_b200 import ops
    rng = np.random.RandomState(9)
    wins = [(0, 0, np.array([1.0])), (2, 2, np.array([-0.2, -0.1, 0.0, 0.1, 0.2])),
            (1, 1, np.array([1.0, -2.0, 1.0]))]
    B, T, sd = 3, 40, 2
    mu = rng.randn(B, T, 3 * sd).astype(np.float32)
    var = (0.3 + rng.rand(B, T, 3 * sd)).astype(np.float32)
    got = ops.mlpg_var(torch.from_numpy(mu).to(dev), torch.from_numpy(var).to(dev), ops.windows_key(wins)).cpu().numpy()
    for b in range(B):
        np.testing.assert_allclose(got[b], nnp.mlpg(mu[b], var[b], wins), rtol=2e-5, atol=2e-6)


def test_compat_paramgen_mlpg_numpy_api(dev):
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        from nnmnkwii import paramgen
        rng = np.random.RandomState(1)
        mu = rng.randn(30, 9)
        out = paramgen.mlpg(mu, np.ones(9), WINDOWS)
        assert out.shape == (30, 3) and out.dtype == np.float64
        np.testing.assert_allclose(out, nnp.mlpg(mu, np.ones(9), WINDOWS), rtol=2e-5, atol=2e-6)
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))


def test_evaluation_gen_parameters_on_cuda_paramgen(dev):
    """The evaluation call site of MLPG (reference evaluation_tts.py:50-97, both branches) through the product's
    nnmnkwii.paramgen.mlpg (compat shim -> gantts_mlpg_var on the GPU) and nnmnkwii.preprocessing.inv_scale, against the
    outputs of the unmodified reference function (tests/golden/eval.npz)."""
    if os.path.join(ROOT, "compat") not in sys.path:
        sys.path.insert(1, os.path.join(ROOT, "compat"))
    from nnmnkwii import paramgen, preprocessing
    import evaltts_mirror
    g = np.load(os.path.join(ROOT, "tests", "golden", "eval.npz"))
    for tag, mge in (("mge", True), ("var", False)):
        got = evaltts_mirror.gen_parameters(g["eval_y"].copy(), g["eval_mean"], g["eval_std"], mge, [180, 3, 1, 3], WINDOWS,
                                            paramgen, preprocessing)
        for k, v in zip(("mgc", "lf0", "vuv", "bap"), got):
            want = g["eval_%s_%s" % (tag, k)]
            assert np.abs(np.asarray(v, dtype=np.float64) - want).max() <= 2e-5 * np.abs(want).max(), (tag, k)

