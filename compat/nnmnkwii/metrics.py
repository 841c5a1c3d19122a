"""nnmnkwii.metrics shim: the objective distortions logged by reference train.py:399-432 (MCD, lf0 MSE on
voiced frames, V/UV error, MSE).  Restated from the published definitions (the package is not vendored:
parity unpinned); logging only, not on the training hot path.  Accept numpy arrays or torch tensors
(B, T, D) with per-utterance `lengths`."""
import numpy as np

_logdb_const = 10.0 / np.log(10.0) * np.sqrt(2.0)


def _np(a):
    try:
        import torch
        if torch.is_tensor(a):
            return a.detach().float().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(a, dtype=np.float64)


def _lengths(lengths, X):
    if lengths is None:
        return [X.shape[1]] * X.shape[0]
    return [int(v) for v in _np(lengths).reshape(-1)]


def _as3(a):
    a = _np(a)
    if a.ndim == 2:
        a = a[:, :, None]
    return a


def melcd(X, Y, lengths=None):
    """Mel-cepstral distortion in dB: 10/ln10 * sqrt(2) * mean over valid frames of ||x - y||_2."""
    X, Y = _as3(X), _as3(Y)
    s, T = 0.0, 0
    for x, y, n in zip(X, Y, _lengths(lengths, X)):
        z = x[:n] - y[:n]
        s += np.sqrt((z * z).sum(-1)).sum()
        T += n
    return _logdb_const * float(s) / float(T)


def mean_squared_error(X, Y, lengths=None):
    X, Y = _as3(X), _as3(Y)
    s, T = 0.0, 0
    for x, y, n in zip(X, Y, _lengths(lengths, X)):
        z = x[:n] - y[:n]
        s += (z * z).sum()
        T += n
    return float(s) / float(T)


def lf0_mean_squared_error(src_f0, src_vuv, tgt_f0, tgt_vuv, lengths=None, linear_domain=False):
    """MSE of (log-)F0 over frames voiced in BOTH source and target; ZeroDivisionError when there is none."""
    sf, tf, sv, tv = _as3(src_f0), _as3(tgt_f0), _as3(src_vuv), _as3(tgt_vuv)
    if linear_domain:
        sf, tf = np.exp(sf), np.exp(tf)
    s, T = 0.0, 0
    for x, y, a, b, n in zip(sf, tf, sv, tv, _lengths(lengths, sf)):
        voiced = ((a[:n] + b[:n]) >= 2).reshape(-1)
        z = x[:n].reshape(n, -1)[voiced] - y[:n].reshape(n, -1)[voiced]
        s += (z * z).sum()
        T += int(voiced.sum())
    return float(s) / float(T)


def vuv_error(src_vuv, tgt_vuv, lengths=None):
    sv, tv = _as3(src_vuv), _as3(tgt_vuv)
    s, T = 0.0, 0
    for a, b, n in zip(sv, tv, _lengths(lengths, sv)):
        s += (a[:n] != b[:n]).sum()
        T += n
    return float(s) / float(T)
