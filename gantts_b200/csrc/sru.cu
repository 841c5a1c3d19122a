// SRU (Simple Recurrent Unit, v1) recurrence -- the scan behind the third-party `cuda_functional.SRU`
// that reference gantts/models.py:144-167 (SRURNN) imports (github.com/taolei87/sru, NOT vendored, no
// reference test touches it: parity unpinned; restated from the published recurrence):
//   U = x W  with k = 3 (n_in == dirs*d) or 4 gates per hidden unit, laid out [.., column, k] (k fastest)
//   f = sigmoid(U_1 + b_f);  r = sigmoid(U_2 + b_r);  c_t = f c_{t-1} + (1 - f) U_0
//   h_t = r (g(c_t) mask) + (1 - r) x'_t,   x' = x (k == 3) or U_3 (k == 4),  g = tanh | relu | identity
// The GEMM is the tensor-core engine (caller); this kernel is the element-wise scan: one thread per
// (batch row, column) walking the T steps, columns [0,d) forward in time, [d,2d) backward.  HBM-bound.
#include "common.cuh"

namespace gantts {

struct SruParams {
  const float* u;       // [B][T][ncols*k]
  const float* x;       // [B][T][ncols] highway input when k == 3 (else null)
  const float* bias;    // [2*ncols]
  const float* mask_h;  // [B][ncols] output dropout mask (already scaled) or null
  float* h;             // [B][T][ncols]
  float* c;             // [B][T][ncols] cell states (saved for the backward)
  const float* dh;      // bwd
  float* du;            // bwd [B][T][ncols*k]
  float* dx;            // bwd [B][T][ncols] (+=) when k == 3
  float* dbias_part;    // bwd [B][2*ncols]
  int B, T, d, k, bidir, act;   // act: 0 identity, 1 tanh, 2 relu
};

__device__ __forceinline__ float sru_act(float c, int act) { return act == 1 ? tanhf(c) : (act == 2 ? fmaxf(c, 0.f) : c); }
__device__ __forceinline__ float sru_dact(float c, float val, int act) {
  return act == 1 ? (1.f - val * val) : (act == 2 ? (c > 0.f ? 1.f : 0.f) : 1.f);
}

__global__ void sru_fwd_kernel(const SruParams p) {
  const int ncols = p.d * (p.bidir ? 2 : 1);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * ncols) return;
  const int b = idx / ncols, col = idx - b * ncols;
  const bool rev = p.bidir && col >= p.d;
  const float bf = p.bias[col], br = p.bias[col + ncols];
  const float m = p.mask_h ? p.mask_h[idx] : 1.f;
  float c = 0.f;
  for (int s = 0; s < p.T; ++s) {
    const int t = rev ? p.T - 1 - s : s;
    const int64_t row = (int64_t)b * p.T + t;
    const float* up = p.u + (row * ncols + col) * p.k;
    const float u0 = up[0];
    const float g1 = 1.f / (1.f + expf(-(up[1] + bf)));
    const float g2 = 1.f / (1.f + expf(-(up[2] + br)));
    const float xp = p.k == 3 ? p.x[row * ncols + col] : up[3];
    c = (c - u0) * g1 + u0;
    p.c[row * ncols + col] = c;
    const float val = sru_act(c, p.act);
    p.h[row * ncols + col] = (val * m - xp) * g2 + xp;
  }
}

__global__ void sru_bwd_kernel(const SruParams p) {
  const int ncols = p.d * (p.bidir ? 2 : 1);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * ncols) return;
  const int b = idx / ncols, col = idx - b * ncols;
  const bool rev = p.bidir && col >= p.d;
  const float bf = p.bias[col], br = p.bias[col + ncols];
  const float m = p.mask_h ? p.mask_h[idx] : 1.f;
  float dc = 0.f, gbf = 0.f, gbr = 0.f;
  for (int s = p.T - 1; s >= 0; --s) {              // reverse of the forward scan order
    const int t = rev ? p.T - 1 - s : s;
    const int tprev = rev ? t + 1 : t - 1;
    const int64_t row = (int64_t)b * p.T + t;
    const float* up = p.u + (row * ncols + col) * p.k;
    float* dup = p.du + (row * ncols + col) * p.k;
    const float u0 = up[0];
    const float g1 = 1.f / (1.f + expf(-(up[1] + bf)));
    const float g2 = 1.f / (1.f + expf(-(up[2] + br)));
    const float xp = p.k == 3 ? p.x[row * ncols + col] : up[3];
    const float c = p.c[row * ncols + col];
    const float cprev = s > 0 ? p.c[((int64_t)b * p.T + tprev) * ncols + col] : 0.f;
    const float val = sru_act(c, p.act);
    const float dhv = p.dh[row * ncols + col];
    const float dg2 = dhv * (val * m - xp);
    const float dxp = dhv * (1.f - g2);
    const float dct = dc + dhv * g2 * m * sru_dact(c, val, p.act);
    const float du0 = dct * (1.f - g1);
    const float dg1 = dct * (cprev - u0);
    dc = dct * g1;
    const float du1 = dg1 * g1 * (1.f - g1), du2 = dg2 * g2 * (1.f - g2);
    dup[0] = du0; dup[1] = du1; dup[2] = du2;
    if (p.k == 3) p.dx[row * ncols + col] += dxp; else dup[3] = dxp;
    gbf += du1; gbr += du2;
  }
  p.dbias_part[(int64_t)b * 2 * ncols + col] = gbf;
  p.dbias_part[(int64_t)b * 2 * ncols + ncols + col] = gbr;
}

static int sru_check(const SruParams& p) {
  GANTTS_CHECK_ARG(p.B >= 1 && p.T >= 1 && p.d >= 1 && (p.k == 3 || p.k == 4), "sru: bad shape (B=%d T=%d d=%d k=%d)",
                   p.B, p.T, p.d, p.k);
  GANTTS_CHECK_ARG(p.act >= 0 && p.act <= 2, "sru: bad activation %d", p.act);
  GANTTS_CHECK_ARG(p.u && p.bias && (p.k == 4 || p.x), "sru: null pointer");
  return GANTTS_OK;
}

}  // namespace gantts

using namespace gantts;

extern "C" int gantts_sru_fwd(const float* u, const float* x, const float* bias, const float* mask_h, float* h, float* c,
                              int B, int T, int d, int k, int bidir, int act, void* stream) {
  SruParams p{};
  p.u = u; p.x = x; p.bias = bias; p.mask_h = mask_h; p.h = h; p.c = c;
  p.B = B; p.T = T; p.d = d; p.k = k; p.bidir = bidir ? 1 : 0; p.act = act;
  int rc = sru_check(p);
  if (rc) return rc;
  GANTTS_CHECK_ARG(h && c, "sru_fwd: null output");
  const int n = B * d * (bidir ? 2 : 1);
  sru_fwd_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(p);
  GANTTS_LAUNCH_CHECK("sru_fwd_kernel");
  return GANTTS_OK;
}

extern "C" int gantts_sru_bwd(const float* u, const float* x, const float* bias, const float* mask_h, const float* c,
                              const float* dh, float* du, float* dx, float* dbias_part, int B, int T, int d, int k,
                              int bidir, int act, void* stream) {
  SruParams p{};
  p.u = u; p.x = x; p.bias = bias; p.mask_h = mask_h; p.c = const_cast<float*>(c); p.dh = dh; p.du = du; p.dx = dx;
  p.dbias_part = dbias_part;
  p.B = B; p.T = T; p.d = d; p.k = k; p.bidir = bidir ? 1 : 0; p.act = act;
  int rc = sru_check(p);
  if (rc) return rc;
  GANTTS_CHECK_ARG(c && dh && du && dbias_part && (k == 4 || dx), "sru_bwd: null pointer");
  const int n = B * d * (bidir ? 2 : 1);
  sru_bwd_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(p);
  GANTTS_LAUNCH_CHECK("sru_bwd_kernel");
  return GANTTS_OK;
}
