"""nnmnkwii.paramgen shim: ``unit_variance_mlpg_matrix(windows, T)`` (reference train.py:511) and
``mlpg(mean_frames, variance_frames, windows)`` (reference evaluation_tts.py:70-72,92-94).

The reference rebuilds this dense (T, num_windows*T) float32 matrix on the CPU for every batch
(O(T^2) banded inverse + O(T^3) dense product in the real package).  Here it is assembled once per
(windows, T) from the rows of P^-1 that libgantts_b200.so computes by banded Cholesky (entries
beyond +-25 frames are < 1e-9 of the diagonal and are left at zero) and memoised; the CUDA MLPG
only uses its SHAPE (it re-derives the same P^-1 rows on the device side)."""
import numpy as np

from gantts_b200 import ops, _lib

_cache = {}


def unit_variance_mlpg_matrix(windows, T):
    T = int(T)
    key = (ops.windows_key(windows), T)
    R = _cache.get(key)
    if R is not None:
        return R
    ops.register_windows(windows)
    tab = ops.mlpg_table_host(windows, T).astype(np.float64)       # (T, 49): P^-1[t, t-24 .. t+24]
    K = _lib.MLPG_HALF_TAPS
    nw = len(windows)
    R = np.zeros((T, nw * T), dtype=np.float64)
    for t in range(T):
        lo, hi = max(0, t - K), min(T, t + K + 1)
        prow = tab[t, lo - t + K:hi - t + K]                        # P^-1[t, lo:hi]
        for w, (l, u, coef) in enumerate(windows):
            # R[t, w*T + r] = sum_k P^-1[t, r+k] * coef[k+l]
            for k in range(-l, u + 1):
                c = float(coef[k + l])
                if c == 0.0:
                    continue
                r_lo, r_hi = max(0, lo - k), min(T, hi - k)
                if r_hi > r_lo:
                    R[t, w * T + r_lo:w * T + r_hi] += c * prow[r_lo + k - lo:r_hi + k - lo]
    R = np.ascontiguousarray(R.astype(np.float32))
    _cache[key] = R
    return R


def mlpg(mean_frames, variance_frames, windows):
    """Static trajectory (T, sd) from (T, nw*sd) means and (nw*sd,) or (T, nw*sd) variances: numpy in, numpy
    out like the real package; the banded solve runs on the GPU (gantts_mlpg_var)."""
    import torch
    mean_frames = np.asarray(mean_frames)
    dtype = mean_frames.dtype if mean_frames.dtype in (np.float32, np.float64) else np.float64
    nw = len(windows)
    if nw == 1 and tuple(windows[0][:2]) == (0, 0):
        return mean_frames                                        # static-only features: nothing to solve
    dev = torch.device("cuda")
    mu = torch.as_tensor(np.ascontiguousarray(mean_frames, dtype=np.float32), device=dev)
    var = torch.as_tensor(np.ascontiguousarray(variance_frames, dtype=np.float32), device=dev)
    return ops.mlpg_var(mu, var, ops.windows_key(windows)).cpu().numpy().astype(dtype)
