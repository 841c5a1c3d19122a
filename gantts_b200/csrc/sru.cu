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

// The scan is sequential in t only through c; everything it READS (U, x, saved c, dh) is known up front.  A
// plain loop issues one dependent HBM round trip per step (the stores to c / h keep the compiler from hoisting
// the next step's loads): measured 1.3 ms per layer at B=32, T=1000, 1024 columns, 15x the HBM time.  So the
// steps are processed in chunks of SRU_UNR: all loads of a chunk are issued first (independent, SRU_UNR deep
// per thread), then the recurrence runs on registers.
constexpr int SRU_UNR = 8;

template <int K>
__global__ void __launch_bounds__(128) sru_fwd_kernel(const SruParams p) {
  const int ncols = p.d * (p.bidir ? 2 : 1);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * ncols) return;
  const int b = idx / ncols, col = idx - b * ncols;
  const bool rev = p.bidir && col >= p.d;
  const float bf = p.bias[col], br = p.bias[col + ncols];
  const float m = p.mask_h ? p.mask_h[idx] : 1.f;
  const int64_t base = (int64_t)b * p.T * ncols + col;       // element (b, t = 0, col)
  const int64_t tstep = rev ? -(int64_t)ncols : (int64_t)ncols;
  const int64_t first = rev ? base + (int64_t)(p.T - 1) * ncols : base;
  float c = 0.f;
  for (int s0 = 0; s0 < p.T; s0 += SRU_UNR) {
    float u0[SRU_UNR], u1[SRU_UNR], u2[SRU_UNR], xp[SRU_UNR];
#pragma unroll
    for (int j = 0; j < SRU_UNR; ++j) {
      u0[j] = u1[j] = u2[j] = xp[j] = 0.f;
      if (s0 + j < p.T) {
        const int64_t e = first + (int64_t)(s0 + j) * tstep;
        if (K == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p.u + e * 4));
          u0[j] = v.x; u1[j] = v.y; u2[j] = v.z; xp[j] = v.w;
        } else {
          const float* up = p.u + e * 3;
          u0[j] = __ldg(up); u1[j] = __ldg(up + 1); u2[j] = __ldg(up + 2);
          xp[j] = __ldg(p.x + e);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < SRU_UNR; ++j) {
      if (s0 + j < p.T) {
        const int64_t e = first + (int64_t)(s0 + j) * tstep;
        const float g1 = 1.f / (1.f + expf(-(u1[j] + bf)));
        const float g2 = 1.f / (1.f + expf(-(u2[j] + br)));
        c = (c - u0[j]) * g1 + u0[j];
        p.c[e] = c;
        const float val = sru_act(c, p.act);
        p.h[e] = (val * m - xp[j]) * g2 + xp[j];
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(128) sru_bwd_kernel(const SruParams p) {
  const int ncols = p.d * (p.bidir ? 2 : 1);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * ncols) return;
  const int b = idx / ncols, col = idx - b * ncols;
  const bool rev = p.bidir && col >= p.d;
  const float bf = p.bias[col], br = p.bias[col + ncols];
  const float m = p.mask_h ? p.mask_h[idx] : 1.f;
  const int64_t base = (int64_t)b * p.T * ncols + col;
  const int64_t tstep = rev ? -(int64_t)ncols : (int64_t)ncols;      // forward-scan direction
  const int64_t first = rev ? base + (int64_t)(p.T - 1) * ncols : base;
  float dc = 0.f, gbf = 0.f, gbr = 0.f;
  // scan steps s = T-1 .. 0 (reverse of the forward order); element of step s is first + s * tstep
  for (int s0 = p.T - 1; s0 >= 0; s0 -= SRU_UNR) {
    float u0[SRU_UNR], u1[SRU_UNR], u2[SRU_UNR], xp[SRU_UNR], cs[SRU_UNR + 1], dhv[SRU_UNR], dxo[SRU_UNR];
#pragma unroll
    for (int j = 0; j < SRU_UNR; ++j) {
      u0[j] = u1[j] = u2[j] = xp[j] = cs[j] = dhv[j] = dxo[j] = 0.f;
      const int sidx = s0 - j;
      if (sidx >= 0) {
        const int64_t e = first + (int64_t)sidx * tstep;
        if (K == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p.u + e * 4));
          u0[j] = v.x; u1[j] = v.y; u2[j] = v.z; xp[j] = v.w;
        } else {
          const float* up = p.u + e * 3;
          u0[j] = __ldg(up); u1[j] = __ldg(up + 1); u2[j] = __ldg(up + 2);
          xp[j] = __ldg(p.x + e);
          dxo[j] = p.dx[e];
        }
        cs[j] = p.c[e];
        dhv[j] = __ldg(p.dh + e);
      }
    }
    {
      const int sidx = s0 - SRU_UNR;                       // c of the step before the chunk's last one
      cs[SRU_UNR] = sidx >= 0 ? p.c[first + (int64_t)sidx * tstep] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < SRU_UNR; ++j) {
      const int sidx = s0 - j;
      if (sidx >= 0) {
        const int64_t e = first + (int64_t)sidx * tstep;
        const float g1 = 1.f / (1.f + expf(-(u1[j] + bf)));
        const float g2 = 1.f / (1.f + expf(-(u2[j] + br)));
        const float c = cs[j];
        const float cprev = sidx > 0 ? cs[j + 1] : 0.f;
        const float val = sru_act(c, p.act);
        const float dg2 = dhv[j] * (val * m - xp[j]);
        const float dxp = dhv[j] * (1.f - g2);
        const float dct = dc + dhv[j] * g2 * m * sru_dact(c, val, p.act);
        const float du0 = dct * (1.f - g1);
        const float dg1 = dct * (cprev - u0[j]);
        dc = dct * g1;
        const float du1 = dg1 * g1 * (1.f - g1), du2 = dg2 * g2 * (1.f - g2);
        if (K == 4) {
          *reinterpret_cast<float4*>(p.du + e * 4) = make_float4(du0, du1, du2, dxp);
        } else {
          float* dup = p.du + e * 3;
          dup[0] = du0; dup[1] = du1; dup[2] = du2;
          p.dx[e] = dxo[j] + dxp;
        }
        gbf += du1; gbr += du2;
      }
    }
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
  if (k == 4) sru_fwd_kernel<4><<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(p);
  else sru_fwd_kernel<3><<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(p);
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
  if (k == 4) sru_bwd_kernel<4><<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(p);
  else sru_bwd_kernel<3><<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(p);
  GANTTS_LAUNCH_CHECK("sru_bwd_kernel");
  return GANTTS_OK;
}
