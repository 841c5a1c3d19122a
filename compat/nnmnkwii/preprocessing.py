"""nnmnkwii.preprocessing shim (host-side numpy; data pipeline, not on the GPU hot path)."""
import numpy as np


def delta_features(x, windows):
    T, D = x.shape
    out = np.zeros((T, D * len(windows)), dtype=x.dtype)
    for wi, (l, u, coef) in enumerate(windows):
        for k in range(-l, u + 1):
            lo, hi = max(0, -k), min(T, T - k)
            out[lo:hi, wi * D:(wi + 1) * D] += coef[k + l] * x[lo + k:hi + k]
    return out


def scale(x, data_mean, data_std):
    return (x - data_mean) / data_std


def inv_scale(x, data_mean, data_std):
    return data_std * x + data_mean


def minmax_scale_params(data_min, data_max, feature_range=(0, 1)):
    data_range = data_max - data_min
    data_range = np.where(data_range == 0, 1.0, data_range)
    scale_ = (feature_range[1] - feature_range[0]) / data_range
    min_ = feature_range[0] - data_min * scale_
    return min_, scale_


def minmax_scale(x, data_min=None, data_max=None, feature_range=(0, 1), scale_=None, min_=None):
    if scale_ is None or min_ is None:
        min_, scale_ = minmax_scale_params(data_min, data_max, feature_range)
    return x * scale_ + min_


def meanvar(dataset, lengths=None, mean_=0.0, var_=0.0, last_sample_count=0, return_last_sample_count=False):
    n, s, ss = last_sample_count, 0.0, 0.0
    for idx in range(len(dataset)):
        x = dataset[idx]
        if lengths is not None:
            x = x[:lengths[idx]]
        s = s + x.sum(axis=0)
        ss = ss + (x.astype(np.float64) ** 2).sum(axis=0)
        n += len(x)
    mean = s / n
    var = ss / n - mean ** 2
    if return_last_sample_count:
        return mean, var, n
    return mean, var


def minmax(dataset, lengths=None):
    mn, mx = None, None
    for idx in range(len(dataset)):
        x = dataset[idx]
        if lengths is not None:
            x = x[:lengths[idx]]
        a, b = x.min(axis=0), x.max(axis=0)
        mn = a if mn is None else np.minimum(mn, a)
        mx = b if mx is None else np.maximum(mx, b)
    return mn, mx
