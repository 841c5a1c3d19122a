"""Device-side `compute_distortions` (reference train.py:399-432 with split_streams :383-396 and inv_scale
:358-380): same dictionary, same formulas, ONE kernel pass + one 32-byte read instead of pulling both
(B, T, D) tensors to the host and looping over utterances in numpy."""
import math

import numpy as np
import torch

from . import ops
from .multistream import get_static_stream_sizes

_LOGDB = 10.0 / math.log(10.0) * math.sqrt(2.0)      # nnmnkwii.metrics.melcd constant


def _dev(a, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(a, dtype=np.float32), device=device)


def _lengths(lengths, B, T, device):
    if lengths is None:
        return torch.full((B,), T, dtype=torch.int64, device=device)
    if torch.is_tensor(lengths):
        return lengths.to(device=device, dtype=torch.int64)
    return torch.as_tensor([int(v) for v in lengths], dtype=torch.int64, device=device)


def compute_distortions(y_static, y_hat_static, Y_data_mean, Y_data_std, lengths=None, hp=None):
    """Drop-in for train.compute_distortions (train.py:399): `hp` carries name, stream_sizes,
    has_dynamic_features, windows (acoustic) or order (vc)."""
    device = y_static.device
    B, T, D = y_static.shape
    lens = _lengths(lengths, B, T, device)
    Ym, Ys = np.asarray(Y_data_mean, dtype=np.float64), np.asarray(Y_data_std, dtype=np.float64)
    if hp.name == "acoustic":
        nw = len(hp.windows)
        mgc_dim, lf0_dim, vuv_dim, bap_dim = hp.stream_sizes
        s_mgc, s_lf0, s_vuv, s_bap = [int(v) for v in get_static_stream_sizes(
            hp.stream_sizes, hp.has_dynamic_features, nw)]
        lf0_i, vuv_i, bap_i = mgc_dim, mgc_dim + lf0_dim, mgc_dim + lf0_dim + vuv_dim
        # per static column (mean, std): the reference indexes the static+dynamic-domain statistics
        src = (list(range(0, mgc_dim // nw)) + list(range(lf0_i, lf0_i + lf0_dim // nw)) + [vuv_i] +
               list(range(bap_i, bap_i + bap_dim // nw)))
        assert len(src) == D, "static width %d does not match the stream layout (%d)" % (D, len(src))
        mean, std = _dev(Ym[src], device), _dev(Ys[src], device)
        lf0_c, vuv_c, bap_c = s_mgc, s_mgc + s_lf0, s_mgc + s_lf0 + s_vuv
        s = ops.distortion_sums(y_static, y_hat_static, lens, mean, std, mcd=(1, s_mgc - 1), bap=(bap_c, s_bap),
                                lf0_col=lf0_c, vuv_col=vuv_c, lf0_linear=True).cpu().double().numpy()
        frames = s[5]
        f0_mse = s[2] / s[3] if s[3] > 0 else float("nan")     # the reference maps ZeroDivisionError to nan
        return {"mcd": _LOGDB * s[0] / frames, "bap_mcd": _LOGDB * s[1] / frames / 10.0,
                "f0_rmse": math.sqrt(f0_mse) if f0_mse == f0_mse else float("nan"), "vuv_err": s[4] / frames}
    if hp.name == "duration":
        mean, std = _dev(Ym.reshape(-1)[:D], device), _dev(Ys.reshape(-1)[:D], device)
        s = ops.distortion_sums(y_static, y_hat_static, lens, mean, std, mse=(0, D)).cpu().double().numpy()
        return {"dur_rmse": math.sqrt(s[6] / s[5])}
    if hp.name == "vc":
        sdim = hp.order
        mean, std = _dev(Ym[:sdim], device), _dev(Ys[:sdim], device)
        s = ops.distortion_sums(y_static, y_hat_static, lens, mean, std, mcd=(0, D)).cpu().double().numpy()
        return {"mcd": _LOGDB * s[0] / s[5]}
    raise AssertionError("unknown hparams name %r" % (hp.name,))
