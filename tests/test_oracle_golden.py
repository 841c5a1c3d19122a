"""Pin the CPU oracle (oracle/gantts_port.py + oracle/nnmnkwii_port.py) against golden vectors
produced by the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import WINDOWS, TTS_HP, rel_err
from oracle import gantts_port as gp
from oracle import nnmnkwii_port as nnp

F32_TOL = 2e-6   # same torch CPU ops in the same order; allow for thread-count dependent reductions


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_sequence_mask_bit_exact(golden_ops):
    lengths = T(golden_ops["mask_lengths"])
    assert np.array_equal(gp.sequence_mask(lengths).numpy(), golden_ops["mask"])
    assert np.array_equal(gp.sequence_mask(lengths, 30).numpy(), golden_ops["mask_maxlen30"])


def test_masked_mse(golden_ops):
    a = T(golden_ops["mse_in"]).requires_grad_(True)
    b = T(golden_ops["mse_tgt"])
    lengths = T(golden_ops["mask_lengths"])
    loss = gp.masked_mse(a, b, lengths=lengths)
    loss.backward()
    assert rel_err(loss.detach().numpy(), golden_ops["mse_loss"]) < F32_TOL
    assert rel_err(a.grad.numpy(), golden_ops["mse_grad"]) < F32_TOL
    m = gp.sequence_mask(lengths).unsqueeze(-1)
    assert rel_err(gp.masked_mse(a, b, mask=m).detach().numpy(), golden_ops["mse_loss_mask"]) < F32_TOL
    with pytest.raises(RuntimeError):
        gp.masked_mse(a, b)


def test_stream_indexing_bit_exact(golden_ops):
    x = torch.arange(0, 63).float().expand(2, 4, 63)
    for name in ("1111", "1000", "1001", "0010", "0101"):
        streams = [c == "1" for c in name]
        assert np.array_equal(gp.select_streams(x, [60, 1, 1, 1], streams).numpy(),
                              golden_ops["select_" + name])
    assert np.array_equal(gp.get_static_stream_sizes([180, 3, 1, 3], [True, True, False, True], 3),
                          golden_ops["static_sizes"])
    y = T(golden_ops["ms_in"])
    assert np.array_equal(gp.get_static_features(y, 3).numpy(), golden_ops["static_all"])
    assert np.array_equal(
        gp.get_static_features(y, 3, streams=[True, False, False, True]).numpy(),
        golden_ops["static_1001"])


def test_multi_stream_mlpg(golden_ops):
    y = T(golden_ops["ms_in"]).requires_grad_(True)
    R = T(nnp.unit_variance_mlpg_matrix(WINDOWS, y.shape[1]))
    z = gp.multi_stream_mlpg(y, R)
    z.backward(T(golden_ops["mlpg_gout"]))
    assert rel_err(z.detach().numpy(), golden_ops["mlpg_out"]) < F32_TOL
    assert rel_err(y.grad.numpy(), golden_ops["mlpg_gin"]) < F32_TOL
    # vuv stream is copied through bit-exactly (reference tests/test_gantts.py:158)
    assert np.array_equal(z.detach().numpy()[:, :, 61], golden_ops["ms_in"][:, :, 183])
    z2 = gp.multi_stream_mlpg(y.detach(), R, streams=[True, False, True, False])
    assert rel_err(z2.numpy(), golden_ops["mlpg_out_1010"]) < F32_TOL
    with pytest.raises(RuntimeError):
        gp.multi_stream_mlpg(y.detach()[:, :, :100], R)


def test_mlpg_dense_vs_banded_f64(golden_ops):
    """The reference arithmetic (dense fp32 R matmul) against the independent fp64 banded solve."""
    y = golden_ops["ms_in"][:, :, :180]
    ref64 = nnp.mlpg_solve_f64(WINDOWS, y)
    assert rel_err(golden_ops["mlpg_out"][:, :, :60], ref64) < 5e-6
    v = golden_ops["vc_in"]
    assert rel_err(golden_ops["vc_out"], nnp.mlpg_solve_f64(WINDOWS, v)) < 5e-6
    assert rel_err(golden_ops["w2_out"], nnp.mlpg_solve_f64(WINDOWS[:2], golden_ops["w2_in"])) < 5e-6


def test_normal_matrix_is_pentadiagonal():
    """P interior [0.75,-4,7.5,-4,0.75], corners 6.25."""
    P = nnp.normal_matrix(WINDOWS, 12)
    assert np.allclose(np.diagonal(P)[1:-1], 7.5) and P[0, 0] == 6.25 and P[-1, -1] == 6.25
    assert np.allclose(np.diagonal(P, 1), -4.0) and np.allclose(np.diagonal(P, 2), 0.75)
    assert np.count_nonzero(np.triu(P, 3)) == 0


def _layers(g, prefix, n_hidden):
    ls = [(T(g[prefix + "layers.%d.weight" % i]), T(g[prefix + "layers.%d.bias" % i]))
          for i in range(n_hidden)]
    ls.append((T(g[prefix + "last_linear.weight"]), T(g[prefix + "last_linear.bias"])))
    return ls


def test_mlp_forward_backward(golden_models):
    g = golden_models
    layers = [(W.requires_grad_(True), b.requires_grad_(True)) for W, b in _layers(g, "mlpg_", 3)]
    x = T(g["mlp_g_x"]).requires_grad_(True)
    y = gp.mlp_forward(x, layers)
    y.backward(T(g["mlp_g_gy"]))
    assert rel_err(y.detach().numpy(), g["mlp_g_y"]) < F32_TOL
    assert rel_err(x.grad.numpy(), g["mlp_g_gx"]) < F32_TOL
    for i, (W, b) in enumerate(layers[:-1]):
        assert rel_err(W.grad.numpy(), g["mlp_g_grad_layers.%d.weight" % i]) < F32_TOL
        assert rel_err(b.grad.numpy(), g["mlp_g_grad_layers.%d.bias" % i]) < F32_TOL
    yd = gp.mlp_forward(T(g["mlp_d_x"]), _layers(g, "mlpd_", 3), last_sigmoid=True)
    assert rel_err(yd.numpy(), g["mlp_d_y"]) < F32_TOL


def test_in2out_highway(golden_models):
    g = golden_models
    layers = [(T(g["hw_H.%d.weight" % i]), T(g["hw_H.%d.bias" % i])) for i in range(2)]
    layers.append((T(g["hw_last_linear.weight"]), T(g["hw_last_linear.bias"])))
    gate = (T(g["hw_T.weight"]), T(g["hw_T.bias"]))
    x = T(g["hw_x"])
    R = T(nnp.unit_variance_mlpg_matrix(WINDOWS, x.shape[1]))
    y, ys = gp.in2out_highway_forward(x, R, gate, layers, static_dim=10)
    assert rel_err(y.numpy(), g["hw_y"]) < F32_TOL
    assert rel_err(ys.numpy(), g["hw_ystatic"]) < F32_TOL


def test_lstm(golden_models):
    g = golden_models
    lstm = torch.nn.LSTM(12, 16, 2, batch_first=True, bidirectional=True)
    lstm.load_state_dict({k[len("lstm_lstm."):]: T(g[k]) for k in g.files if k.startswith("lstm_lstm.")})
    lstm.eval()
    h2o = (T(g["lstm_hidden2out.weight"]), T(g["lstm_hidden2out.bias"]))
    with torch.no_grad():
        y = gp.lstm_forward(T(g["lstm_x"]), g["lstm_lengths"], lstm, h2o)
    assert rel_err(y.numpy(), g["lstm_y"]) < F32_TOL


@pytest.mark.parametrize("tag,cond", [("u_", False), ("c_", True)])
def test_gan_step_two_iterations(golden_step, tag, cond):
    """Losses, generator outputs and post-step weights of two consecutive reference mini-batches."""
    g = golden_step
    state = gp.GanStepState(_layers(g, tag + "g0_", 3), _layers(g, tag + "d0_", 3))
    hp = dict(TTS_HP, discriminator_linguistic_condition=cond)
    R = T(nnp.unit_variance_mlpg_matrix(WINDOWS, 30))
    for it in range(2):
        p = "%sit%d_" % (tag, it)
        out, y_hat, y_hat_static = gp.gan_step_mlp(
            state, T(g[p + "x"]), T(g[p + "y"]), g[p + "lengths"], R, hp,
            w_d=1.0, mse_w=0.0, mge_w=1.0, adv_w=1.0)
        ref = g[p + "losses"]
        got = [out["loss_d"], out["loss_fake_d"], out["loss_real_d"], out["loss_mse"],
               out["loss_mge"], out["loss_adv"], out["loss_g"]]
        assert np.allclose(got, ref, rtol=2e-6, atol=0), (got, ref)
        assert [out["real_correct"], out["fake_correct"]] == list(g[p + "counts"])
        assert rel_err(y_hat.numpy(), g[p + "y_hat"]) < F32_TOL
        assert rel_err(y_hat_static.numpy(), g[p + "y_hat_static"]) < F32_TOL
        names = ["layers.0", "layers.1", "layers.2", "last_linear"]
        for (W, b), n in zip(state.g, names):
            assert rel_err(W.detach().numpy(), g[p + "g_" + n + ".weight"]) < 1e-5
            assert rel_err(b.detach().numpy(), g[p + "g_" + n + ".bias"]) < 1e-5
        for (W, b), n in zip(state.d, names):
            assert rel_err(W.detach().numpy(), g[p + "d_" + n + ".weight"]) < 1e-5
            assert rel_err(b.detach().numpy(), g[p + "d_" + n + ".bias"]) < 1e-5


VC_TOY_HP = dict(stream_sizes=[27], has_dynamic_features=[True], adversarial_streams=[True],
                 mask_nth_mgc_for_adv_loss=0, num_windows=3, discriminator_linguistic_condition=False)
STEP_MODEL_CASES = {
    "hw_": ("highway", dict(static_dim=9), VC_TOY_HP),
    "rhw_": ("rnn_highway", dict(static_dim=9, num_hidden=2, hidden_dim=12, bidirectional=True), VC_TOY_HP),
    "lstm_": ("lstm", dict(num_hidden=2, hidden_dim=16, bidirectional=True), TTS_HP),
}
LOSS_KEYS = ("loss_d", "loss_fake_d", "loss_real_d", "loss_mse", "loss_mge", "loss_adv", "loss_g",
             "real_correct", "fake_correct")


@pytest.mark.parametrize("tag", sorted(STEP_MODEL_CASES))
def test_gan_step_non_mlp_generators(golden_step_models, tag):
    """The generic step of the port (gp.gan_step + GeneratorOracle) against the reference's own step functions
    run with In2OutHighwayNet (no discriminator, BASELINE cfg1), In2OutRNNHighwayNet + MLP D (cfg3) and
    LSTMRNN + MLP D (cfg5) at toy sizes: losses, counts, outputs and post-step weights of two mini-batches."""
    g = golden_step_models
    kind, kw, hp = STEP_MODEL_CASES[tag]
    sub = lambda pre: {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
    gen = gp.GeneratorOracle(kind, sub(tag + "g0_"), **kw)
    dsd = sub(tag + "d0_")
    d_layers = gp.discriminator_layers(dsd) if dsd else None
    d_sum = [torch.zeros_like(t) for pair in d_layers for t in pair] if d_layers else None
    w_d, mse_w, mge_w = [float(v) for v in g[tag + "cfg"]]
    R = T(nnp.unit_variance_mlpg_matrix(WINDOWS, 24))
    for it in range(2):
        p = "%sit%d_" % (tag, it)
        x, y, lens = T(g[p + "x"]), T(g[p + "y"]), [int(v) for v in g[p + "lengths"]]
        out, y_hat, y_hat_static = gp.gan_step(lambda: gen.forward(x, R, lens, hp, training=True), gen.params(),
                                               gen.sums, d_layers, d_sum, x, y, lens, R, hp, w_d=w_d, mse_w=mse_w,
                                               mge_w=mge_w, adv_w=1.0)
        ref = dict(zip(LOSS_KEYS, g[p + "losses"]))
        for k, v in ref.items():
            if np.isnan(v):
                assert k not in out or w_d == 0
            elif k.endswith("correct"):
                assert out[k] == v, (k, out[k], v)
            else:
                assert abs(out[k] - v) <= 2e-6 * max(abs(v), 1e-3), (k, out[k], v)
        assert rel_err(y_hat.numpy(), g[p + "y_hat"]) < F32_TOL
        assert rel_err(y_hat_static.numpy(), g[p + "y_hat_static"]) < F32_TOL
        for k, v in gen.named.items():
            assert rel_err(v.detach().numpy(), g[p + "g_" + k]) < 1e-5, k
        if d_layers:
            names = ["layers.%d" % i for i in range(len(d_layers) - 1)] + ["last_linear"]
            for (W, b), n in zip(d_layers, names):
                assert rel_err(W.detach().numpy(), g[p + "d_" + n + ".weight"]) < 1e-5
                assert rel_err(b.detach().numpy(), g[p + "d_" + n + ".bias"]) < 1e-5


def test_variance_mlpg_restatement_reduces_to_R():
    """oracle mlpg (evaluation-time, real variances) with unit variance == the pinned R-matrix product."""
    rng = np.random.RandomState(2)
    T, sd = 37, 3
    mu = rng.randn(T, 3 * sd)
    R = nnp.unit_variance_mlpg_matrix(WINDOWS, T).astype(np.float64)
    wm = np.concatenate([mu[:, w * sd:(w + 1) * sd] for w in range(3)], axis=0)
    np.testing.assert_allclose(nnp.mlpg(mu, np.ones(3 * sd), WINDOWS), R @ wm, rtol=0, atol=1e-6)
    # a per-dimension rescaling of all windows of one dimension leaves its trajectory unchanged
    v = np.ones(3 * sd)
    v[[0, sd, 2 * sd]] = 7.0
    np.testing.assert_allclose(nnp.mlpg(mu, v, WINDOWS), nnp.mlpg(mu, np.ones(3 * sd), WINDOWS), atol=1e-9)


def test_metrics_restatement_hand_values_and_agrees_with_the_shim():
    """oracle/nnmnkwii_port metrics (the checker of csrc/metrics.cu) against hand-computed values, and against the product's
    separately written numpy shim (compat/nnmnkwii/metrics.py): two restatements of the published nnmnkwii definitions
    (package not vendored: parity unpinned) that share no code."""
    import importlib.util
    import os
    import types
    from oracle import nnmnkwii_port as M
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_shim_metrics", os.path.join(root, "compat", "nnmnkwii", "metrics.py"))
    shim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shim)
    rs = np.random.RandomState(1)
    X, Y = rs.randn(2, 5, 4), rs.randn(2, 5, 4)
    lens = [5, 3]
    exp = 10 / np.log(10) * np.sqrt(2) * (np.sqrt(((X[0] - Y[0]) ** 2).sum(-1)).sum() +
                                          np.sqrt(((X[1, :3] - Y[1, :3]) ** 2).sum(-1)).sum()) / 8
    assert abs(M.melcd(X, Y, lens) - exp) < 1e-9
    assert abs(M.mean_squared_error(X, Y, lens) - (((X[0] - Y[0]) ** 2).sum() + ((X[1, :3] - Y[1, :3]) ** 2).sum()) / 8) < 1e-12
    v1 = np.array([[1, 1, 0, 1, 1], [1, 0, 1, 1, 1]], float)
    v2 = np.array([[1, 0, 0, 1, 1], [1, 1, 1, 0, 0]], float)
    assert M.vuv_error(v1, v2, lens) == 2 / 8
    f1, f2 = rs.rand(2, 5, 1), rs.rand(2, 5, 1)
    both = [(0, 0), (0, 3), (0, 4), (1, 0), (1, 2)]
    exp = sum((f1[b, t, 0] - f2[b, t, 0]) ** 2 for b, t in both) / len(both)
    assert abs(M.lf0_mean_squared_error(f1, v1, f2, v2, lens) - exp) < 1e-12
    with pytest.raises(ZeroDivisionError):
        M.lf0_mean_squared_error(f1, v1 * 0, f2, v2, lens)
    for fn in ("melcd", "mean_squared_error"):
        assert abs(getattr(M, fn)(X, Y, lens) - getattr(shim, fn)(X, Y, lens)) < 1e-12
    assert abs(M.lf0_mean_squared_error(f1, v1, f2, v2, lens, linear_domain=True) -
               shim.lf0_mean_squared_error(f1, v1, f2, v2, lens, linear_domain=True)) < 1e-12
    # train.py:399-432 on top of them: acoustic split + inverse scaling
    from oracle import gantts_port as gp
    from conftest import WINDOWS
    hp = types.SimpleNamespace(name="acoustic", windows=WINDOWS, stream_sizes=[180, 3, 1, 3],
                               has_dynamic_features=[True, True, False, True])
    y = rs.randn(2, 6, 63)
    yh = y + 0.1 * rs.randn(2, 6, 63)
    Ym, Ys = rs.randn(187) * 0.1, 0.5 + rs.rand(187)
    Ym[183], Ys[183] = 0.5, 0.5
    d = gp.compute_distortions(y, yh, Ym, Ys, [6, 4], hp)
    mgc = y[:, :, 1:60] * Ys[1:60] + Ym[1:60]
    mgc_h = yh[:, :, 1:60] * Ys[1:60] + Ym[1:60]
    assert abs(d["mcd"] - M.melcd(mgc, mgc_h, [6, 4])) < 1e-12
    assert set(d) == {"mcd", "bap_mcd", "f0_rmse", "vuv_err"}


def test_adam_stepper_matches_torch_optim_adam():
    """oracle AdamStepper (the duration model's optimiser, reference hparams.py:125-130: Adam, lr 1e-3, betas (0.5, 0.9),
    weight_decay 0) against torch.optim.Adam over several steps, with and without weight decay."""
    import torch
    from oracle import gantts_port as gp
    for wd in (0.0, 1e-3):
        torch.manual_seed(0)
        ps = [torch.randn(7, 5), torch.randn(5)]
        a = [p.clone().requires_grad_(True) for p in ps]
        b = [p.clone() for p in ps]
        opt = torch.optim.Adam(a, lr=1e-3, betas=(0.5, 0.9), eps=1e-8, weight_decay=wd)
        st = gp.AdamStepper(b, lr=1e-3, betas=(0.5, 0.9), eps=1e-8, weight_decay=wd)
        for k in range(5):
            gs = [torch.randn_like(p) * (k + 1) for p in ps]
            for p, g in zip(a, gs):
                p.grad = g.clone()
            opt.step()
            st(b, [g.clone() for g in gs])
            for p, q in zip(a, b):
                assert float((p.detach() - q).abs().max()) < 2e-7


def test_evaluation_gen_parameters_mirror_matches_reference():
    """tests/evaltts_mirror.gen_parameters (what the GPU suite drives the CUDA paramgen.mlpg through) against the outputs of the
    UNMODIFIED reference evaluation_tts.py:50-97 (tests/golden/eval.npz, generated by make_golden.py::gen_eval), both
    branches: MLPG with unit variance on normalised features, and with the real variances after inverse scaling."""
    import os
    import types
    import evaltts_mirror
    from conftest import WINDOWS
    from oracle import nnmnkwii_port as nnp
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval.npz"))
    P = types.SimpleNamespace(inv_scale=lambda x, m, s: s * x + m)
    pg = types.SimpleNamespace(mlpg=nnp.mlpg)
    for tag, mge in (("mge", True), ("var", False)):
        got = evaltts_mirror.gen_parameters(g["eval_y"].copy(), g["eval_mean"], g["eval_std"], mge, [180, 3, 1, 3], WINDOWS, pg, P)
        for k, v in zip(("mgc", "lf0", "vuv", "bap"), got):
            assert np.abs(np.asarray(v) - g["eval_%s_%s" % (tag, k)]).max() < 1e-10, (tag, k)

