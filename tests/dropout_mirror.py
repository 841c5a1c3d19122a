"""numpy mirror of the product's counter-hash dropout mask (gantts_b200/csrc/common.cuh: dropout_quad_bits / dropout_keep) --
test infrastructure: documents the mask function bit for bit, checks its statistics on the CPU and pins the device kernels
to it on the GPU.  keep(row, col) = field(hash(seed, row * ceil(N/4) + col/4)) >= round(p * 65536)."""
import numpy as np

_M = np.uint64(0xFFFFFFFF)


def quad_bits(seed, row, quarter_n, quad):
    """(a, b) 32-bit words of the quad: a -> columns 4q, 4q+1 (low, high 16-bit field), b -> columns 4q+2, 4q+3."""
    u = np.uint64
    x = ((row.astype(u) * u(quarter_n) + quad.astype(u)) * u(0x9E3779B1) + u(seed & 0xFFFFFFFF)) & _M
    x ^= x >> u(16)
    x = (x * u(0x7FEB352D)) & _M
    x ^= x >> u(15)
    x = (x * u(0x846CA68B)) & _M
    x ^= x >> u(16)
    y = (x * u(0x9E3779B1)) & _M
    y ^= y >> u(15)
    s = u((seed >> 32) & 0xFFFFFFFF)
    return (x ^ s) & _M, (y ^ s) & _M


def keep_mask(seed, rows, n_cols, p):
    """Boolean (rows, n_cols): True where the element is kept."""
    thresh = int(round(p * 65536.0))
    r, c = np.meshgrid(np.arange(rows), np.arange(n_cols), indexing="ij")
    a, b = quad_bits(int(seed), r, (n_cols + 3) // 4, c // 4)
    w = np.where((c & 2) > 0, b, a)
    f = np.where((c & 1) > 0, w >> np.uint64(16), w & np.uint64(0xFFFF))
    return f >= np.uint64(thresh)


def dropout_multiplier(seed, rows, n_cols, p):
    """float32 {0, 1/(1-p)} like gantts_dropout on a tensor of ones."""
    return keep_mask(seed, rows, n_cols, p).astype(np.float32) * np.float32(1.0 / (1.0 - p))
