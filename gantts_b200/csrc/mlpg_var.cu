// Inference-time parameter generation with real variances: nnmnkwii.paramgen.mlpg(mean_frames,
// variance_frames, windows) as called by the reference's evaluation scripts (evaluation_tts.py:70-72,
// 92-94; the package itself is not vendored, the maths is the published MLPG):
//     y = argmax N(W y; mu, Sigma)  <=>  (sum_w W_w^T diag(1/var_w) W_w) y = sum_w W_w^T diag(1/var_w) mu_w
// for every static dimension independently.  The system is banded SPD (half bandwidth hb = max_w(l_w + u_w),
// 2 for the reference windows).  One thread per (batch row, static dimension) assembles its band row by
// row, factors it (banded Cholesky) and substitutes forwards while walking t upwards, then substitutes
// backwards walking down; L and z live in a caller workspace laid out [t][slot][column] so that the
// threads of a warp (adjacent columns) touch adjacent addresses.  fp64 arithmetic like bandmat's.
// Latency-bound in T (sequential recurrence), parallel over B * sd columns; an inference-time op.
#include "common.cuh"

namespace gantts {

constexpr int MV_HB = GANTTS_MAX_WINDOW_TAPS - 1;   // largest half bandwidth of W^T W: l + u <= taps - 1

__global__ void __launch_bounds__(128)
mlpg_var_kernel(const float* __restrict__ mean, int64_t m_bs, int64_t m_ts, const float* __restrict__ var,
                int64_t v_bs, int64_t v_ts, float* __restrict__ out, int64_t o_bs, int64_t o_ts,
                const gantts_windows_t w, int B, int T, int sd, int hb, double* __restrict__ ws) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t ncols = (int64_t)B * sd;
  if (col >= ncols) return;
  const int b = (int)(col / sd), d = (int)(col - (int64_t)b * sd);
  const float* mu = mean + b * m_bs + d;
  const float* vr = var + b * v_bs + d;
  const int slots = hb + 2;                               // L[i][i-hb..i] (hb+1 values) and z[i]
  // ring of the last MV_HB rows of L: Lr[r][k] = L[i-r-1][i-r-1-k]  (k = 0 is the diagonal)
  double Lr[MV_HB][MV_HB + 1];
  for (int r = 0; r < MV_HB; ++r)
    for (int k = 0; k <= MV_HB; ++k) Lr[r][k] = 0.0;
  double zr[MV_HB];
  for (int r = 0; r < MV_HB; ++r) zr[r] = 0.0;

  for (int i = 0; i < T; ++i) {
    // band row i of P (columns i-hb..i) and right-hand side
    double prow[MV_HB + 1];
    for (int k = 0; k <= MV_HB; ++k) prow[k] = 0.0;
    double rhs = 0.0;
    for (int wi = 0; wi < w.n; ++wi) {
      // rows t' of W_w touching column i: i - u <= t' <= i + l
      for (int tp = i - w.u[wi]; tp <= i + w.l[wi]; ++tp) {
        if (tp < 0 || tp >= T) continue;
        const double ci = (double)w.coef[wi][i - tp + w.l[wi]];
        const double tau = 1.0 / (double)vr[(int64_t)tp * v_ts + (int64_t)wi * sd];
        rhs += ci * tau * (double)mu[(int64_t)tp * m_ts + (int64_t)wi * sd];
        for (int k = 0; k <= hb; ++k) {
          const int j = i - k;                            // column j <= i
          const int off = j - tp + w.l[wi];
          if (j < 0 || off < 0 || off > w.l[wi] + w.u[wi]) continue;
          prow[k] += ci * tau * (double)w.coef[wi][off];
        }
      }
    }
    // Cholesky row: L[i][j] for j = i-hb..i   (lrow[k] = L[i][i-k])
    double lrow[MV_HB + 1];
    for (int k = 0; k <= MV_HB; ++k) lrow[k] = 0.0;
    for (int k = hb; k >= 0; --k) {
      const int j = i - k;
      if (j < 0) continue;
      double s = prow[k];
      // sum over m < j, m >= i - hb:  L[i][m] * L[j][m];  L[i][m] = lrow[i-m], L[j][m] = row j's entry (j-m)
      for (int m = i - hb; m < j; ++m) {
        if (m < 0) continue;
        const double lim = lrow[i - m];
        const double ljm = (k == 0) ? lim : Lr[k - 1][j - m];   // row j = i-k is ring slot k-1
        if (k == 0 || j - m <= hb) s -= lim * ljm;
      }
      if (k == 0) lrow[0] = sqrt(s);
      else lrow[k] = s / Lr[k - 1][0];
    }
    // forward substitution z[i] = (rhs - sum_{k>=1} L[i][i-k] z[i-k]) / L[i][i]
    double zz = rhs;
    for (int k = 1; k <= hb; ++k)
      if (i - k >= 0) zz -= lrow[k] * zr[k - 1];
    zz /= lrow[0];
    double* wrow = ws + ((int64_t)i * slots) * ncols + col;
    for (int k = 0; k <= hb; ++k) wrow[(int64_t)k * ncols] = lrow[k];
    wrow[(int64_t)(hb + 1) * ncols] = zz;
    // rotate the rings
    for (int r = MV_HB - 1; r > 0; --r) {
      for (int k = 0; k <= MV_HB; ++k) Lr[r][k] = Lr[r - 1][k];
      zr[r] = zr[r - 1];
    }
    for (int k = 0; k <= MV_HB; ++k) Lr[0][k] = lrow[k];
    zr[0] = zz;
  }
  // backward substitution L^T y = z: y[i] = (z[i] - sum_{k=1..hb} L[i+k][i] y[i+k]) / L[i][i]
  double yr[MV_HB];
  for (int r = 0; r < MV_HB; ++r) yr[r] = 0.0;
  // ring of L rows above: Lup[r][k] = L[i+r+1][i+r+1-k]
  for (int r = 0; r < MV_HB; ++r)
    for (int k = 0; k <= MV_HB; ++k) Lr[r][k] = 0.0;
  float* o = out + b * o_bs + d;
  for (int i = T - 1; i >= 0; --i) {
    const double* wrow = ws + ((int64_t)i * slots) * ncols + col;
    double lrow[MV_HB + 1];
    for (int k = 0; k <= MV_HB; ++k) lrow[k] = k <= hb ? wrow[(int64_t)k * ncols] : 0.0;
    double yy = wrow[(int64_t)(hb + 1) * ncols];
    for (int k = 1; k <= hb; ++k)
      if (i + k < T) yy -= Lr[k - 1][k] * yr[k - 1];    // L[i+k][i] = row (i+k)'s entry k
    yy /= lrow[0];
    o[(int64_t)i * o_ts] = (float)yy;
    for (int r = MV_HB - 1; r > 0; --r) {
      for (int k = 0; k <= MV_HB; ++k) Lr[r][k] = Lr[r - 1][k];
      yr[r] = yr[r - 1];
    }
    for (int k = 0; k <= MV_HB; ++k) Lr[0][k] = lrow[k];
    yr[0] = yy;
  }
}

static int windows_half_band(const gantts_windows_t* w) {
  int hb = 0;
  for (int i = 0; i < w->n; ++i)
    if (w->l[i] + w->u[i] > hb) hb = w->l[i] + w->u[i];
  return hb;
}

}  // namespace gantts

using namespace gantts;

extern "C" size_t gantts_mlpg_var_workspace_bytes(const gantts_windows_t* windows, int B, int T, int sd) {
  if (!windows || B < 1 || T < 1 || sd < 1) return 0;
  return (size_t)T * (size_t)(windows_half_band(windows) + 2) * (size_t)B * sd * sizeof(double) + 256;
}

extern "C" int gantts_mlpg_var(const float* mean, int64_t m_bstride, int64_t m_tstride, const float* var,
                               int64_t v_bstride, int64_t v_tstride, float* out, int64_t o_bstride,
                               int64_t o_tstride, const gantts_windows_t* windows, int B, int T, int sd,
                               void* workspace, size_t workspace_bytes, void* stream) {
  GANTTS_CHECK_ARG(mean && var && out && windows, "mlpg_var: null argument");
  GANTTS_CHECK_ARG(B >= 1 && T >= 1 && sd >= 1, "mlpg_var: bad sizes");
  GANTTS_CHECK_ARG(windows->n >= 1 && windows->n <= GANTTS_MAX_WINDOWS, "mlpg_var: bad window count");
  for (int i = 0; i < windows->n; ++i)
    GANTTS_CHECK_ARG(windows->l[i] >= 0 && windows->u[i] >= 0 &&
                         windows->l[i] + windows->u[i] + 1 <= GANTTS_MAX_WINDOW_TAPS,
                     "mlpg_var: window too wide");
  const int hb = windows_half_band(windows);
  const size_t need = gantts_mlpg_var_workspace_bytes(windows, B, T, sd);
  if (!workspace || workspace_bytes < need) {
    set_error("mlpg_var: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GANTTS_E_WORKSPACE;
  }
  double* ws = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  const int64_t ncols = (int64_t)B * sd;
  mlpg_var_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, as_stream(stream)>>>(
      mean, m_bstride, m_tstride, var, v_bstride, v_tstride, out, o_bstride, o_tstride, *windows, B, T, sd, hb, ws);
  GANTTS_LAUNCH_CHECK("mlpg_var_kernel");
  return GANTTS_OK;
}
