"""torch.autograd bindings of the C-ABI ops.  Every op requires CUDA float32 tensors; there is no
CPU path (the product must fail loudly rather than fall back)."""
import ctypes

import numpy as np
import torch

from . import _lib
from . import config

LEAKY_SLOPE = 0.01          # nn.LeakyReLU() default used by the reference (gantts/models.py:37,132)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw getter is ~20x cheaper than building a
    torch.cuda.Stream object; it is called once per native launch)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("gantts_b200: CUDA tensor required (got a %s tensor); this package has "
                               "no CPU fallback" % t.device.type)
        if t.dtype != torch.float32:
            raise RuntimeError("gantts_b200: float32 tensor required (got %s)" % t.dtype)


def _rows2d(x):
    """View (…, D) as (rows, D) with unit column stride; returns (tensor2d, row_stride)."""
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < D):
        x2 = x2.contiguous()
    return x2, (x2.stride(0) if x2.shape[0] > 1 else D)


_ws_cache = {}


def workspace(nbytes, device, tag="ws"):
    """Cached per-(device, tag) scratch buffer, grown geometrically (stream-ordered reuse on the
    current stream only)."""
    key = (device.index, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ------------------------------------------------------------------------------------- MLPG
STANDARD_WINDOWS = {
    1: [(0, 0, (1.0,))],
    2: [(0, 0, (1.0,)), (1, 1, (-0.5, 0.0, 0.5))],
    3: [(0, 0, (1.0,)), (1, 1, (-0.5, 0.0, 0.5)), (1, 1, (1.0, -2.0, 1.0))],
}
_registered_windows = {}
_table_cache = {}
_validated_R = set()


def windows_key(windows):
    return tuple((int(l), int(u), tuple(float(c) for c in coef)) for l, u, coef in windows)


def register_windows(windows):
    """Declare the delta windows behind the R matrices of a given window count (the defaults are
    the reference's hparams windows, hparams.py:22-26,183-187)."""
    _registered_windows[len(windows)] = windows_key(windows)


def windows_for(num_windows):
    w = _registered_windows.get(num_windows)
    if w is None:
        if num_windows not in STANDARD_WINDOWS:
            raise RuntimeError("gantts_b200: no windows registered for num_windows=%d" % num_windows)
        w = windows_key(STANDARD_WINDOWS[num_windows])
    return w


def mlpg_table_full_host(windows, T):
    """The coefficient table the kernels read, float32 (T, GANTTS_MLPG_TABLE_COLS): rows of P^-1 within +-24 taps
    followed by the rows of the banded Cholesky factor of P; host fp64 computation inside the C library."""
    lib = _lib.load()
    w = _lib.make_windows(windows)
    tab = np.zeros((int(T), _lib.MLPG_TABLE_COLS), dtype=np.float32)
    _lib.check(lib.gantts_mlpg_table(ctypes.byref(w), int(T), tab.ctypes.data))
    return tab


def mlpg_table_host(windows, T):
    """Rows of P^-1 within +-24 taps, float32 (T, 49)."""
    return np.ascontiguousarray(mlpg_table_full_host(windows, T)[:, :_lib.MLPG_NTAPS])


def mlpg_table(windows, T, device):
    key = (windows_key(windows), int(T), device.index)
    t = _table_cache.get(key)
    if t is None:
        t = torch.from_numpy(mlpg_table_full_host(windows, T)).to(device)
        _table_cache[key] = t
    return t


def _validate_R(R, windows, T):
    """One-time check per (num_windows, T) that a dense R handed in by the caller really is
    (W^T W)^-1 W^T for the registered windows (the kernels never read R on the hot path)."""
    key = (windows, int(T))
    if key in _validated_R:
        return
    tab = mlpg_table_host(windows, T)
    t = int(T) // 2
    row = R[t].detach().float().cpu().numpy()
    K = _lib.MLPG_HALF_TAPS
    worst = 0.0
    for w, (l, u, coef) in enumerate(windows):
        for r in range(max(0, t - K + 2), min(int(T), t + K - 1)):
            exp = 0.0
            for k in range(-l, u + 1):
                c = r + k
                j = c - t + K
                if 0 <= c < T and 0 <= j < _lib.MLPG_NTAPS:
                    exp += tab[t, j] * coef[k + l]
            worst = max(worst, abs(exp - row[w * int(T) + r]))
    if worst > 1e-4:
        raise RuntimeError("gantts_b200: the R matrix does not match the registered delta windows "
                           "(max deviation %.3g); call gantts_b200.ops.register_windows(...)" % worst)
    _validated_R.add(key)


def windows_from_R(R):
    """(windows, T) for a dense MLPG matrix R of shape (T, num_windows*T) (reference train.py:511)."""
    T = int(R.shape[0])
    nw = int(R.shape[1]) // T
    if nw * T != int(R.shape[1]):
        raise RuntimeError("gantts_b200: R must have shape (T, num_windows*T)")
    windows = windows_for(nw)
    _validate_R(R, windows, T)
    return windows, T


class _MLPG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, streams, windows, ncols_out):
        require_cuda(x, table)
        lib = _lib.load()
        B, T, D = x.shape
        if x.stride(2) != 1:
            x = x.contiguous()
        out = torch.empty(B, T, ncols_out, dtype=torch.float32, device=x.device)
        _lib.check(lib.gantts_mlpg_fwd(x.data_ptr(), x.stride(0), x.stride(1), out.data_ptr(),
                                       out.stride(0), out.stride(1), table.data_ptr(),
                                       ctypes.byref(streams), ctypes.byref(windows), B, T, _stream()))
        ctx.table, ctx.streams, ctx.windows = table, streams, windows
        ctx.in_shape = (B, T, D)
        return out

    @staticmethod
    def backward(ctx, go):
        require_cuda(go)
        lib = _lib.load()
        B, T, D = ctx.in_shape
        if go.stride(2) != 1:
            go = go.contiguous()
        gi = torch.zeros(B, T, D, dtype=torch.float32, device=go.device)
        _lib.check(lib.gantts_mlpg_bwd(go.data_ptr(), go.stride(0), go.stride(1), gi.data_ptr(),
                                       gi.stride(0), gi.stride(1), ctx.table.data_ptr(),
                                       ctypes.byref(ctx.streams), ctypes.byref(ctx.windows), B, T, 0,
                                       _stream()))
        return gi, None, None, None, None


def mlpg(x, windows, stream_entries, ncols_out):
    """x: (B, T, D) CUDA float32.  stream_entries: [(in_start, sd, dyn, out_start)]."""
    squeeze = x.dim() == 2
    if squeeze:
        x = x.unsqueeze(0)
    table = mlpg_table(windows, x.shape[1], x.device)
    out = _MLPG.apply(x, table, _lib.make_streams(stream_entries), _lib.make_windows(windows), ncols_out)
    return out.squeeze(0) if squeeze else out


def mlpg_var(mean, variance, windows):
    """nnmnkwii.paramgen.mlpg on the device (reference evaluation_tts.py:70-72,92-94): mean (T, nw*sd) or
    (B, T, nw*sd) CUDA float32, variance (nw*sd,) [time-invariant], (T, nw*sd) or (B, T, nw*sd);
    returns the static trajectory (…, T, sd)."""
    require_cuda(mean, variance)
    lib = _lib.load()
    squeeze = mean.dim() == 2
    if squeeze:
        mean = mean.unsqueeze(0)
    B, T, D = mean.shape
    nw = len(windows)
    if D % nw:
        raise RuntimeError("gantts_b200: feature width %d is not a multiple of the window count %d" % (D, nw))
    sd = D // nw
    if mean.stride(2) != 1:
        mean = mean.contiguous()
    variance = variance.contiguous()
    if variance.shape[-1] != D:
        raise RuntimeError("gantts_b200: variance width %d != mean width %d" % (variance.shape[-1], D))
    if variance.dim() == 1:
        v_bs, v_ts = 0, 0
    elif variance.dim() == 2:
        v_bs, v_ts = 0, variance.stride(0)
    else:
        v_bs, v_ts = variance.stride(0), variance.stride(1)
    if variance.dim() >= 2 and variance.shape[-2] != T:
        raise RuntimeError("gantts_b200: variance has %d frames, mean has %d" % (variance.shape[-2], T))
    w = _lib.make_windows(windows)
    out = torch.empty(B, T, sd, dtype=torch.float32, device=mean.device)
    nbytes = lib.gantts_mlpg_var_workspace_bytes(ctypes.byref(w), B, T, sd)
    ws = workspace(nbytes, mean.device, "mlpg_var")
    _lib.check(lib.gantts_mlpg_var(mean.data_ptr(), mean.stride(0), mean.stride(1), variance.data_ptr(), v_bs, v_ts,
                                   out.data_ptr(), out.stride(0), out.stride(1), ctypes.byref(w), B, T, sd,
                                   ws.data_ptr(), ws.numel(), _stream()))
    return out.squeeze(0) if squeeze else out


def distortion_sums(y, y_hat, lengths, mean, std, mcd=(0, 0), bap=(0, 0), lf0_col=-1, vuv_col=-1,
                    lf0_linear=True, mse=(0, 0)):
    """Eight sums behind the objective metrics of reference train.py:399-432 (see include/gantts_b200.h,
    gantts_distortions).  y, y_hat: (B, T, D) CUDA float32; lengths: int64 CUDA (B,); mean/std: (D,) CUDA.
    Returns a CUDA float32 tensor of 8 values (no host synchronisation here)."""
    require_cuda(y, y_hat, mean, std)
    lib = _lib.load()
    B, T, D = y.shape
    if y.stride(2) != 1:
        y = y.contiguous()
    if y_hat.stride(2) != 1:
        y_hat = y_hat.contiguous()
    cols = _lib.DistortionColsT(int(mcd[0]), int(mcd[1]), int(bap[0]), int(bap[1]), int(lf0_col), int(vuv_col),
                                1 if lf0_linear else 0, int(mse[0]), int(mse[1]))
    out = torch.empty(8, dtype=torch.float32, device=y.device)
    ws = workspace(lib.gantts_distortions_workspace_bytes(), y.device, "distortions")
    _lib.check(lib.gantts_distortions(y.data_ptr(), y.stride(0), y.stride(1), y_hat.data_ptr(), y_hat.stride(0),
                                      y_hat.stride(1), lengths.data_ptr(), B, T, D, mean.contiguous().data_ptr(),
                                      std.contiguous().data_ptr(), ctypes.byref(cols), out.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _stream()))
    return out


def unit_variance_mlpg(R, means):
    """Drop-in for nnmnkwii.autograd.unit_variance_mlpg(R, means) (reference
    gantts/multistream.py:120, gantts/models.py:66,115): means (B, T, nw*sd) or (T, nw*sd)."""
    windows, T = windows_from_R(R)
    if means.shape[-2] != T:
        raise RuntimeError("gantts_b200: means has %d frames but R was built for T=%d" % (means.shape[-2], T))
    nw = len(windows)
    D = means.shape[-1]
    if D % nw:
        raise RuntimeError("gantts_b200: feature dim %d not divisible by num_windows %d" % (D, nw))
    sd = D // nw
    return mlpg(means, windows, [(0, sd, True, 0)], sd)


# ---------------------------------------------------------------------------- column gather
_cols_cache = {}


def _cols_tensor(cols, device):
    key = (tuple(cols), device.index)
    t = _cols_cache.get(key)
    if t is None:
        t = torch.tensor(list(cols), dtype=torch.int32, device=device)
        _cols_cache[key] = t
    return t


class _GatherCols(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cols):
        require_cuda(x)
        lib = _lib.load()
        x2, rs = _rows2d(x)
        rows = x2.shape[0]
        out = torch.empty(x.shape[:-1] + (len(cols),), dtype=torch.float32, device=x.device)
        ct = _cols_tensor(cols, x.device)
        _lib.check(lib.gantts_gather_cols(x2.data_ptr(), rs, out.data_ptr(), len(cols), ct.data_ptr(),
                                          len(cols), rows, _stream()))
        ctx.cols, ctx.in_shape = cols, tuple(x.shape)
        return out

    @staticmethod
    def backward(ctx, go):
        lib = _lib.load()
        go = go.contiguous()
        gi = torch.zeros(ctx.in_shape, dtype=torch.float32, device=go.device)
        rows = gi.numel() // ctx.in_shape[-1]
        ct = _cols_tensor(ctx.cols, go.device)
        _lib.check(lib.gantts_scatter_cols_add(go.data_ptr(), len(ctx.cols), gi.data_ptr(),
                                               ctx.in_shape[-1], ct.data_ptr(), len(ctx.cols), rows,
                                               _stream()))
        return gi, None


def gather_cols(x, cols):
    cols = tuple(int(c) for c in cols)
    if len(cols) == 0:
        raise RuntimeError("gantts_b200: empty column selection")
    return _GatherCols.apply(x, cols)


# ------------------------------------------------------------------------------ masks/losses
def sequence_mask(lengths, max_len):
    lib = _lib.load()
    if not lengths.is_cuda:
        raise RuntimeError("gantts_b200: CUDA lengths tensor required; this package has no CPU fallback")
    lengths = lengths.long().contiguous().view(-1)
    B = lengths.numel()
    mask = torch.empty(B, int(max_len), dtype=torch.float32, device=lengths.device)
    _lib.check(lib.gantts_sequence_mask(lengths.data_ptr(), mask.data_ptr(), B, int(max_len), _stream()))
    return mask


class _MaskedSSE(torch.autograd.Function):
    """sums = [sum(((a-b)*m)^2), sum(m)]"""

    @staticmethod
    def forward(ctx, a, b, mask):
        require_cuda(a, b, mask)
        lib = _lib.load()
        a2, ars = _rows2d(a)
        b2, brs = _rows2d(b)
        m = mask.reshape(-1).contiguous()
        rows, D = a2.shape
        if m.numel() != rows or b2.shape != a2.shape:
            raise RuntimeError("gantts_b200: masked MSE shape mismatch")
        sums = torch.empty(2, dtype=torch.float32, device=a.device)
        nb = lib.gantts_masked_sse_workspace_bytes()
        ws = workspace(nb, a.device, "red")
        _lib.check(lib.gantts_masked_sse_fwd(a2.data_ptr(), ars, b2.data_ptr(), brs, m.data_ptr(), rows, D,
                                             sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        ctx.save_for_backward(a2, b2, m)
        ctx.strides = (ars, brs)
        ctx.shape = tuple(a.shape)
        return sums

    @staticmethod
    def backward(ctx, gsums):
        lib = _lib.load()
        a2, b2, m = ctx.saved_tensors
        rows, D = a2.shape
        scale = gsums[0:1].contiguous()
        ga = torch.empty(rows, D, dtype=torch.float32, device=a2.device)
        _lib.check(lib.gantts_masked_sse_bwd(a2.data_ptr(), ctx.strides[0], b2.data_ptr(), ctx.strides[1],
                                             m.data_ptr(), rows, D, scale.data_ptr(), ga.data_ptr(), D, 0,
                                             _stream()))
        return ga.view(ctx.shape), None, None


def masked_mse(inp, target, mask):
    """reference gantts/seqloss.py:41-43: sum((in*m - tgt*m)^2) / sum(m), m of shape (B, T, 1)."""
    sums = _MaskedSSE.apply(inp, target, mask)
    return sums[0] / sums[1]


class _MaskedBCE(torch.autograd.Function):
    """out = [-(log(arg) * m).sum(), count, sum(m)], arg = D+1e-20 (kind 0) or 1-D+1e-20 (kind 1)."""

    @staticmethod
    def forward(ctx, D, mask, kind):
        require_cuda(D, mask)
        lib = _lib.load()
        d = D.reshape(-1).contiguous()
        m = mask.reshape(-1).contiguous()
        if d.numel() != m.numel():
            raise RuntimeError("gantts_b200: masked BCE shape mismatch")
        out = torch.empty(3, dtype=torch.float32, device=D.device)
        ws = workspace(lib.gantts_masked_sse_workspace_bytes(), D.device, "red")
        _lib.check(lib.gantts_masked_bce_fwd(d.data_ptr(), m.data_ptr(), d.numel(), int(kind), out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _stream()))
        ctx.save_for_backward(d, m)
        ctx.kind, ctx.shape = int(kind), tuple(D.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        d, m = ctx.saved_tensors
        scale = gout[0:1].contiguous()
        gD = torch.empty_like(d)
        _lib.check(lib.gantts_masked_bce_bwd(d.data_ptr(), m.data_ptr(), d.numel(), ctx.kind, scale.data_ptr(),
                                             gD.data_ptr(), _stream()))
        return gD.view(ctx.shape), None, None


def masked_bce(D, mask, kind):
    """Un-normalised adversarial BCE sum, correct-count and sum(mask) (reference train.py:258-271)."""
    return _MaskedBCE.apply(D, mask, kind)


# --------------------------------------------------------------------------- fused linear
class _LinearAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, act, slope, p, seed, engine):
        require_cuda(x, W, b)
        lib = _lib.load()
        x2, xrs = _rows2d(x)
        M, K = x2.shape
        N = W.shape[0]
        if W.shape[1] != K:
            raise RuntimeError("gantts_b200: linear shape mismatch (x has %d features, W expects %d)" % (K, W.shape[1]))
        Wc = W.contiguous()
        bc = b.contiguous() if b is not None else None
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        nb = lib.gantts_linear_workspace_bytes(M, N, K, engine)
        ws = workspace(nb, x.device)
        _lib.check(lib.gantts_linear_fwd(x2.data_ptr(), xrs, Wc.data_ptr(), bc.data_ptr() if bc is not None else None,
                                         y.data_ptr(), N, M, N, K, act, slope, p, seed, engine,
                                         ws.data_ptr(), ws.numel(), _stream()))
        ctx.save_for_backward(x2, Wc, y)
        ctx.cfg = (xrs, act, slope, p, engine, b is not None, tuple(x.shape))
        return y.view(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x2, W, y = ctx.saved_tensors
        xrs, act, slope, p, engine, has_bias, xshape = ctx.cfg
        M, K = x2.shape
        N = W.shape[0]
        gy2, gyrs = _rows2d(gy)
        need_gx, need_gW, need_gb = ctx.needs_input_grad[0], ctx.needs_input_grad[1], (has_bias and ctx.needs_input_grad[2])
        gz = torch.empty(M, N, dtype=torch.float32, device=gy.device)
        gx = torch.empty(M, K, dtype=torch.float32, device=gy.device) if need_gx else None
        gW = torch.empty(N, K, dtype=torch.float32, device=gy.device) if need_gW else None
        gb = torch.empty(N, dtype=torch.float32, device=gy.device) if need_gb else None
        nb = lib.gantts_linear_workspace_bytes(M, N, K, engine)
        ws = workspace(nb, gy.device)
        _lib.check(lib.gantts_linear_bwd(gy2.data_ptr(), gyrs, y.data_ptr(), N, x2.data_ptr(), xrs, W.data_ptr(),
                                         gz.data_ptr(), gx.data_ptr() if gx is not None else None, K,
                                         gW.data_ptr() if gW is not None else None,
                                         gb.data_ptr() if gb is not None else None,
                                         M, N, K, act, slope, p, 0, engine, ws.data_ptr(), ws.numel(), _stream()))
        return (gx.view(xshape) if gx is not None else None), gW, gb, None, None, None, None, None


_seed_state = [None, 0]      # [torch.initial_seed() the counter belongs to, draws since then]


def draw_seed():
    """62-bit dropout seed, reproducible under torch.manual_seed: splitmix64 of (torch.initial_seed(), number
    of draws since the last manual_seed).  Pure host arithmetic -- no tensor op, no synchronisation (the old
    form went through torch.randint(...).item() once per dropout layer)."""
    base = torch.initial_seed()
    if _seed_state[0] != base:
        _seed_state[0], _seed_state[1] = base, 0
    _seed_state[1] += 1
    z = (base + 0x9E3779B97F4A7C15 * _seed_state[1]) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return (z ^ (z >> 31)) & ((1 << 62) - 1)


def peek_seeds(n):
    """The next `n` values draw_seed() will return (test hook: regenerate the masks of a step about to run)."""
    saved = list(_seed_state)
    out = [draw_seed() for _ in range(n)]
    _seed_state[0], _seed_state[1] = saved
    if _seed_state[0] is None:
        _seed_state[0], _seed_state[1] = torch.initial_seed(), 0
    return out


def dropout_mask(rows, cols, p, seed, device):
    """The dropout multiplier {0, 1/(1-p)} every engine applies for (seed, rows, cols): float32 (rows, cols).
    Test hook for injected-mask parity against the CPU checker (gantts_dropout on a tensor of ones)."""
    lib = _lib.load()
    ones = torch.ones(int(rows), int(cols), dtype=torch.float32, device=device)
    out = torch.empty_like(ones)
    _lib.check(lib.gantts_dropout(ones.data_ptr(), out.data_ptr(), int(rows), int(cols), float(p), int(seed),
                                  _stream()))
    return out


def mlp_dropout_masks(rows, hidden_dims, p, seed, device):
    """Masks of the hidden layers of an MLP run (mlp_stack / gantts_mlp_fwd) with dropout seed `seed`."""
    lib = _lib.load()
    return [dropout_mask(rows, n, p, lib.gantts_mlp_layer_seed(int(seed), l), device)
            for l, n in enumerate(hidden_dims)]


def linear_act(x, W, b, act=_lib.ACT_NONE, p=0.0, training=False, slope=LEAKY_SLOPE, engine=None, seed=None):
    """act(x W^T + b): the reference's Linear -> LeakyReLU(0.01) -> Dropout(p) hidden layer
    (gantts/models.py:137-139), ``last_linear`` (act NONE) or Linear -> Sigmoid."""
    eng = config.engine_id(engine)
    p_eff = float(p) if (training and act == _lib.ACT_LEAKY_DROPOUT) else 0.0
    if p_eff > 0.0 and seed is None:
        seed = draw_seed()
    return _LinearAct.apply(x, W, b, int(act), float(slope), p_eff, int(seed or 0), eng)


# ------------------------------------------------------------------ whole MLP (tensor cores)
class _MLPStack(torch.autograd.Function):
    """y = MLP(x) through gantts_mlp_fwd / gantts_mlp_bwd: one tcgen05 GEMM per layer and direction,
    activations resident as bf16 hi/lo planes (the tape)."""

    @staticmethod
    def forward(ctx, x, slope, p, last_act, seed, *params):
        require_cuda(x, *params)
        lib = _lib.load()
        L = len(params) // 2
        if L > _lib.MAX_LAYERS:
            raise RuntimeError("gantts_b200: at most %d layers" % _lib.MAX_LAYERS)
        x2, xrs = _rows2d(x)
        M = x2.shape[0]
        Ws = [params[2 * i].contiguous() for i in range(L)]
        bs = [params[2 * i + 1].contiguous() for i in range(L)]
        d = _lib.MlpT()
        d.num_layers = L
        d.dims[0] = x2.shape[1]
        for i in range(L):
            if Ws[i].shape[1] != d.dims[i]:
                raise RuntimeError("gantts_b200: MLP layer %d expects %d inputs, got %d" % (i, Ws[i].shape[1], d.dims[i]))
            d.dims[i + 1] = Ws[i].shape[0]
            d.W[i] = Ws[i].data_ptr()
            d.b[i] = bs[i].data_ptr()
        d.slope, d.dropout_p, d.last_act, d.seed = float(slope), float(p), int(last_act), int(seed)
        N = d.dims[L]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        tape = torch.empty(lib.gantts_mlp_tape_bytes(ctypes.byref(d), M), dtype=torch.uint8, device=x.device)
        _lib.check(lib.gantts_mlp_fwd(ctypes.byref(d), x2.data_ptr(), xrs, M, y.data_ptr(), N, tape.data_ptr(),
                                      tape.numel(), _stream()))
        ctx.save_for_backward(tape, y, *Ws, *bs)
        ctx.cfg = (d.dims[:L + 1], float(slope), float(p), int(last_act), int(seed), tuple(x.shape))
        return y.view(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        saved = ctx.saved_tensors
        tape, y = saved[0], saved[1]
        dims, slope, p, last_act, seed, xshape = ctx.cfg
        L = len(dims) - 1
        Ws, bs = saved[2:2 + L], saved[2 + L:2 + 2 * L]
        d = _lib.MlpT()
        d.num_layers = L
        for i, v in enumerate(dims):
            d.dims[i] = v
        for i in range(L):
            d.W[i], d.b[i] = Ws[i].data_ptr(), bs[i].data_ptr()
        d.slope, d.dropout_p, d.last_act, d.seed = slope, p, last_act, seed
        gy2, gyrs = _rows2d(gy)
        M = gy2.shape[0]
        dev = gy.device
        need_gx = ctx.needs_input_grad[0]
        gx = torch.empty(M, dims[0], dtype=torch.float32, device=dev) if need_gx else None
        gWs = [torch.empty_like(W) if ctx.needs_input_grad[5 + 2 * i] else None for i, W in enumerate(Ws)]
        gbs = [torch.empty_like(b) if ctx.needs_input_grad[6 + 2 * i] else None for i, b in enumerate(bs)]
        arr = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() if t is not None else None for t in ts])
        ws = workspace(lib.gantts_mlp_workspace_bytes(ctypes.byref(d), M), dev, "mlp")
        _lib.check(lib.gantts_mlp_bwd(ctypes.byref(d), gy2.data_ptr(), gyrs, y.data_ptr(), dims[L], M,
                                      tape.data_ptr(), tape.numel(), gx.data_ptr() if gx is not None else None,
                                      dims[0], arr(gWs), arr(gbs), 0, ws.data_ptr(), ws.numel(), _stream()))
        grads = []
        for i in range(L):
            grads += [gWs[i], gbs[i]]
        return (gx.view(xshape) if gx is not None else None, None, None, None, None) + tuple(grads)


def mlp_stack(x, weights, biases, p=0.0, training=False, last_act=_lib.ACT_NONE, slope=LEAKY_SLOPE, seed=None):
    """Whole MLP (hidden: Linear -> LeakyReLU -> Dropout; last: Linear [-> sigmoid]) on the tcgen05
    engine.  reference gantts/models.py:137-141."""
    p_eff = float(p) if training else 0.0
    if p_eff > 0.0 and seed is None:
        seed = draw_seed()
    params = []
    for W, b in zip(weights, biases):
        params += [W, b]
    return _MLPStack.apply(x, float(slope), p_eff, int(last_act), int(seed or 0), *params)


def highway_combine(x_static, Tx, Gx):
    """y = x_static + Tx * Gx (reference gantts/models.py:69)."""
    require_cuda(x_static, Tx, Gx)
    return torch.addcmul(x_static, Tx, Gx)
