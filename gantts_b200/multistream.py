"""Drop-in for reference ``gantts/multistream.py`` (multi-stream select / static features / MLPG).

Column bookkeeping is host logic (numpy, bit-exact by construction); the data movement and the
MLPG arithmetic run in CUDA (gantts_gather_cols, gantts_mlpg_fwd/bwd)."""
import numpy as np

from . import ops


def _delta_features(x, windows):
    """numpy 'same' correlation per window (nnmnkwii.preprocessing.delta_features semantics)."""
    T, D = x.shape
    out = np.zeros((T, D * len(windows)), dtype=x.dtype)
    for wi, (l, u, coef) in enumerate(windows):
        for k in range(-l, u + 1):
            lo, hi = max(0, -k), min(T, T - k)
            out[lo:hi, wi * D:(wi + 1) * D] += coef[k + l] * x[lo + k:hi + k]
    return out


def recompute_delta_features(Y, Y_data_mean, Y_data_std, windows, stream_sizes=[180, 3, 1, 3],
                             has_dynamic_features=[True, True, False, True]):
    """Loader-side numpy helper (reference gantts/multistream.py:15-30); stays on the host."""
    start_indices = np.hstack(([0], np.cumsum(stream_sizes)[:-1]))
    end_indices = np.cumsum(stream_sizes)
    static_stream_sizes = get_static_stream_sizes(stream_sizes, has_dynamic_features, len(windows))
    for start_idx, end_idx, static_size, has_dynamic in zip(
            start_indices, end_indices, static_stream_sizes, has_dynamic_features):
        if has_dynamic:
            y_static = Y[:, start_idx:start_idx + static_size]
            Y[:, start_idx:end_idx] = _delta_features(y_static, windows)
    return Y


def get_static_stream_sizes(stream_sizes, has_dynamic_features, num_windows):
    """[180,3,1,3], [T,T,F,T], 3 -> array([60,1,1,1]) (reference gantts/multistream.py:46-53)."""
    static_stream_sizes = np.array(stream_sizes)
    sel = np.asarray(has_dynamic_features, dtype=bool)
    static_stream_sizes[sel] = static_stream_sizes[sel] / num_windows
    return static_stream_sizes


def select_stream_columns(stream_sizes, streams):
    """Host logic: input columns kept by select_streams."""
    cols, start = [], 0
    for size, enabled in zip(stream_sizes, streams):
        size = int(size)
        if enabled:
            cols.extend(range(start, start + size))
        start += size
    return cols


def static_feature_columns(num_windows, stream_sizes, has_dynamic_features, streams):
    """Host logic: input columns kept by get_static_features (multi-stream case)."""
    cols, start = [], 0
    for size, dyn, enabled in zip(stream_sizes, has_dynamic_features, streams):
        size = int(size)
        if enabled:
            width = size // num_windows if dyn else size
            cols.extend(range(start, start + width))
        start += size
    return cols


def select_streams(inputs, stream_sizes=[60, 1, 1, 1], streams=[True, True, True, True]):
    """Column gather of the enabled streams (reference gantts/multistream.py:33-43)."""
    return ops.gather_cols(inputs, select_stream_columns(stream_sizes, streams))


def get_static_features(inputs, num_windows, stream_sizes=[180, 3, 1, 3],
                        has_dynamic_features=[True, True, False, True],
                        streams=[True, True, True, True]):
    """Static columns of a static+delta tensor (reference gantts/multistream.py:56-79)."""
    _, _, D = inputs.size()
    if stream_sizes is None or (len(stream_sizes) == 1 and has_dynamic_features[0]):
        return inputs[:, :, :D // num_windows]
    if len(stream_sizes) == 1 and not has_dynamic_features[0]:
        return inputs
    return ops.gather_cols(inputs, static_feature_columns(num_windows, stream_sizes,
                                                          has_dynamic_features, streams))


def mlpg_stream_entries(stream_sizes, has_dynamic_features, streams, num_windows):
    """Host logic: (in_start, sd, dyn, out_start) per enabled stream + number of output columns."""
    entries, in_start, out_start = [], 0, 0
    for size, dyn, enabled in zip(stream_sizes, has_dynamic_features, streams):
        size = int(size)
        if enabled:
            sd = size // num_windows if dyn else size
            entries.append((in_start, sd, bool(dyn), out_start))
            out_start += sd
        in_start += size
    return entries, out_start


def multi_stream_mlpg(inputs, R, stream_sizes=[180, 3, 1, 3],
                      has_dynamic_features=[True, True, False, True],
                      streams=[True, True, True, True]):
    """Per-stream MLPG for streams with dynamics, copy for the others, concatenated
    (reference gantts/multistream.py:82-123).  One fused launch over all streams; per-column
    arithmetic is identical to a stand-alone unit_variance_mlpg(R, slice) call (bitwise, as
    reference tests/test_gantts.py:156-159 demands)."""
    B, T, D = inputs.size()
    if D != sum(stream_sizes):
        raise RuntimeError("You probably have specified wrong dimention params.")
    if R is None:
        windows, num_windows = ops.windows_for(1), 1
    else:
        windows, TR = ops.windows_from_R(R)
        num_windows = len(windows)
        if TR != T:
            raise RuntimeError("R was built for T=%d but inputs have %d frames" % (TR, T))
    entries, ncols = mlpg_stream_entries(stream_sizes, has_dynamic_features, streams, num_windows)
    if not entries:
        raise RuntimeError("no stream enabled")
    return ops.mlpg(inputs, windows, entries, ncols)
