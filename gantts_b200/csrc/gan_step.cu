// Fused GAN training step (include/gantts_b200.h: gantts_gan_step): the whole mini-batch of reference
// train.py:528-580 -- batch prologue, apply_generator (:336-355), update_discriminator (:245-279),
// update_generator (:282-320) with both clip_grad_norm_ + Adagrad steps -- enqueued on ONE stream by one
// C call, with no host synchronisation: every loss is a device scalar.
//
// Semantics kept from the reference:
//   * the fake-term gradient of loss_d reaches the generator (y_hat_static is not detached, one
//     zero_grad per step): it is accumulated into the SAME upstream buffer as the gradient of loss_g,
//     so the generator/MLPG backward runs ONCE on the summed gradient (gradients are linear; the
//     reference runs it twice and adds the results);
//   * three discriminator forwards with independent dropout masks; the discriminator is updated
//     BEFORE the third forward used by the adversarial loss;
//   * losses are normalised by the number of valid frames, BCE uses log(D + 1e-20) verbatim.
// Real and fake discriminator batches are stacked into one 2M-row batch (one GEMM per layer).
#include "common.cuh"

namespace gantts {

enum ScalarSlot {
  S_REAL = 0,      // [0..2]  real: loss sum, correct count, sum(mask)
  S_FAKE = 3,      // [3..5]
  S_ADV = 6,       // [6..8]
  S_MGE = 9,       // [9..10] sse, sum(mask)
  S_MSE = 11,      // [11..12]
  S_DSUMSQ = 13,
  S_GSUMSQ = 14,
  S_INV_T = 15,    // 1 / frames
  S_ADV_SCALE = 16,
  S_MGE_SCALE = 17,
  S_MSE_SCALE = 18,
  S_COUNT = 32
};

struct ColList {
  int n;
  int c[GANTTS_MAX_COLS];
};

__global__ void gather_cols_list_kernel(const float* __restrict__ in, int64_t in_rs, float* __restrict__ out,
                                        int64_t out_rs, ColList cols, int64_t rows) {
  __shared__ int sc[GANTTS_MAX_COLS];
  for (int i = threadIdx.x; i < cols.n; i += blockDim.x) sc[i] = cols.c[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps)
    for (int j = lane; j < cols.n; j += 32) out[r * out_rs + j] = in[r * in_rs + sc[j]];
}

__global__ void scatter_cols_list_add_kernel(const float* __restrict__ go, int64_t go_rs, float* __restrict__ gi,
                                             int64_t gi_rs, ColList cols, int64_t rows) {
  __shared__ int sc[GANTTS_MAX_COLS];
  for (int i = threadIdx.x; i < cols.n; i += blockDim.x) sc[i] = cols.c[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps)
    for (int j = lane; j < cols.n; j += 32) gi[r * gi_rs + sc[j]] += go[r * go_rs + j];
}

__global__ void set_scales_kernel(float* scal, float inv_frames, float adv_w, float mge_w, float mse_w,
                                  int zero_norms) {
  if (threadIdx.x == 0) {
    if (zero_norms) scal[S_DSUMSQ] = scal[S_GSUMSQ] = 0.f;
    scal[S_INV_T] = inv_frames;
    scal[S_ADV_SCALE] = adv_w * inv_frames;
    scal[S_MGE_SCALE] = mge_w * inv_frames;
    scal[S_MSE_SCALE] = mse_w * inv_frames;
  }
}

// losses[0..11] = loss_d, loss_fake_d, loss_real_d, loss_mse, loss_mge, loss_adv, loss_g,
//                 real_correct, fake_correct, frames(local sum of mask), d_grad_norm, g_grad_norm
__global__ void finalize_losses_kernel(const float* scal, float* losses, float adv_w, float mge_w, float mse_w,
                                       int has_d) {
  if (threadIdx.x != 0) return;
  const float invT = scal[S_INV_T];
  const float real = has_d ? scal[S_REAL] * invT : 0.f, fake = has_d ? scal[S_FAKE] * invT : 0.f;
  const float adv = (has_d && adv_w > 0.f) ? scal[S_ADV] * invT : 0.f;
  const float mge = scal[S_MGE] * invT, mse = scal[S_MSE] * invT;
  losses[0] = real + fake;
  losses[1] = fake;
  losses[2] = real;
  losses[3] = mse;
  losses[4] = mge;
  losses[5] = adv;
  losses[6] = (mse_w * mse + mge_w * mge) + adv_w * adv;
  losses[7] = has_d ? scal[S_REAL + 1] : 0.f;
  losses[8] = has_d ? scal[S_FAKE + 1] : 0.f;
  losses[9] = scal[S_MGE + 1];
  losses[10] = has_d ? sqrtf(scal[S_DSUMSQ]) : 0.f;
  losses[11] = sqrtf(scal[S_GSUMSQ]);
}

static inline int blocks_1d(int64_t work, int per_block) {
  int64_t b = (work + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

struct ParamList {
  int n;
  float* p[2 * GANTTS_MAX_LAYERS];
  float* g[2 * GANTTS_MAX_LAYERS];
  float* s[2 * GANTTS_MAX_LAYERS];
  int64_t sizes[2 * GANTTS_MAX_LAYERS];
  float* gW[GANTTS_MAX_LAYERS];
  float* gb[GANTTS_MAX_LAYERS];
  int64_t total;
};

static inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

struct StepLayout {
  float* scal;
  float* mask;            // [M]
  float* d_in;            // [2M][dD]  rows 0..M-1 real, M..2M-1 fake
  float* d_out;           // [2M]
  float* g_dout;          // [2M]
  float* g_din;           // [2M][dD]
  float* y_static;        // [M][n_static]
  float* g_static;        // [M][n_static]
  float* g_yhat;          // [M][d_out]
  float* g_grads;         // flat generator gradients
  float* d_grads;         // flat discriminator gradients
  char* g_tape;
  size_t g_tape_bytes;
  char* d_tape;
  size_t d_tape_bytes;
  char* mlp_ws;
  size_t mlp_ws_bytes;
  char* red_ws;
  size_t red_ws_bytes;
  size_t total;
};

static int64_t mlp_param_count(const gantts_mlp_t& m) {
  int64_t n = 0;
  for (int l = 0; l < m.num_layers; ++l) n += (int64_t)m.dims[l + 1] * m.dims[l] + m.dims[l + 1];
  return n;
}

static void layout(const gantts_gan_step_t* c, char* base, StepLayout* L) {
  const int64_t M = (int64_t)c->B * c->T;
  const int dD = c->d.dims[0];
  char* cur = base;
  auto take = [&](size_t bytes) { char* p = cur; cur += al256(bytes); return p; };
  L->scal = (float*)take(S_COUNT * sizeof(float));
  L->mask = (float*)take(M * sizeof(float));
  L->d_in = (float*)take((size_t)2 * M * dD * sizeof(float));
  L->d_out = (float*)take((size_t)2 * M * sizeof(float));
  L->g_dout = (float*)take((size_t)2 * M * sizeof(float));
  L->g_din = (float*)take((size_t)2 * M * dD * sizeof(float));
  L->y_static = (float*)take((size_t)M * c->n_static * sizeof(float));
  L->g_static = (float*)take((size_t)M * c->n_static * sizeof(float));
  L->g_yhat = (float*)take((size_t)M * c->g.dims[c->g.num_layers] * sizeof(float));
  L->g_grads = (float*)take(mlp_param_count(c->g) * sizeof(float));
  L->d_grads = (float*)take(mlp_param_count(c->d) * sizeof(float));
  L->g_tape_bytes = gantts_mlp_tape_bytes(&c->g, M);
  L->g_tape = take(L->g_tape_bytes);
  L->d_tape_bytes = gantts_mlp_tape_bytes(&c->d, 2 * M);
  L->d_tape = take(L->d_tape_bytes);
  size_t a = gantts_mlp_workspace_bytes(&c->g, M), b = gantts_mlp_workspace_bytes(&c->d, 2 * M);
  L->mlp_ws_bytes = a > b ? a : b;
  L->mlp_ws = take(L->mlp_ws_bytes);
  size_t r1 = gantts_masked_sse_workspace_bytes(), r2 = gantts_optim_workspace_bytes();
  L->red_ws_bytes = r1 > r2 ? r1 : r2;
  L->red_ws = take(L->red_ws_bytes);
  L->total = (size_t)(cur - base) + 256;
}

static void param_list(const gantts_mlp_t& m, float* const* sumW, float* const* sumb, float* flat, ParamList* pl) {
  pl->n = 0;
  pl->total = 0;
  float* cur = flat;
  for (int l = 0; l < m.num_layers; ++l) {
    const int64_t nw = (int64_t)m.dims[l + 1] * m.dims[l], nb = m.dims[l + 1];
    pl->gW[l] = cur;
    pl->p[pl->n] = const_cast<float*>(m.W[l]);
    pl->g[pl->n] = cur;
    pl->s[pl->n] = sumW[l];
    pl->sizes[pl->n++] = nw;
    cur += nw;
    pl->gb[l] = cur;
    pl->p[pl->n] = const_cast<float*>(m.b[l]);
    pl->g[pl->n] = cur;
    pl->s[pl->n] = sumb[l];
    pl->sizes[pl->n++] = nb;
    cur += nb;
  }
  pl->total = cur - flat;
}

static int check_step(const gantts_gan_step_t* c) {
  GANTTS_CHECK_ARG(c, "gan_step: null config");
  GANTTS_CHECK_ARG(c->B >= 1 && c->T >= 1, "gan_step: bad batch shape");
  GANTTS_CHECK_ARG(c->g.num_layers >= 1 && c->g.num_layers <= GANTTS_MAX_LAYERS, "gan_step: bad generator");
  GANTTS_CHECK_ARG(c->n_static >= 1 && c->n_static <= GANTTS_MAX_COLS, "gan_step: bad n_static");
  GANTTS_CHECK_ARG(c->n_static_cols == c->n_static, "gan_step: static column list must have n_static entries");
  if (c->w_d > 0.f) {
    GANTTS_CHECK_ARG(c->d.num_layers >= 1 && c->d.num_layers <= GANTTS_MAX_LAYERS, "gan_step: bad discriminator");
    const int cond_w = c->d_conditioned ? c->g.dims[0] : 0;
    GANTTS_CHECK_ARG(c->n_adv >= 1 && c->n_adv <= GANTTS_MAX_COLS && c->d.dims[0] == cond_w + c->n_adv,
                     "gan_step: discriminator input width %d != %d conditioning + %d adversarial columns",
                     c->d.dims[0], cond_w, c->n_adv);
    GANTTS_CHECK_ARG(c->d.dims[c->d.num_layers] == 1 && c->d.last_act == GANTTS_ACT_SIGMOID,
                     "gan_step: discriminator must end in a single sigmoid output");
  }
  GANTTS_CHECK_ARG(c->g.last_act == GANTTS_ACT_NONE, "gan_step: generator must have a linear output");
  GANTTS_CHECK_ARG(c->mlpg_table, "gan_step: null MLPG table");
  return GANTTS_OK;
}

}  // namespace gantts

using namespace gantts;

extern "C" uint64_t gantts_gan_step_seed(uint64_t seed, int which) { return seed * 4 + (uint64_t)which; }

extern "C" size_t gantts_gan_step_workspace_bytes(const gantts_gan_step_t* c) {
  if (check_step(c)) return 0;
  StepLayout L;
  layout(c, nullptr, &L);
  return L.total + 256;
}

extern "C" int gantts_gan_step_grad_buffer(const gantts_gan_step_t* c, void* workspace, int which, float** ptr,
                                           int64_t* count) {
  int rc = check_step(c);
  if (rc) return rc;
  GANTTS_CHECK_ARG(workspace && ptr && count, "gan_step_grad_buffer: null pointer");
  StepLayout L;
  layout(c, reinterpret_cast<char*>(al256(reinterpret_cast<uintptr_t>(workspace))), &L);
  *ptr = which == 0 ? L.g_grads : L.d_grads;
  *count = mlp_param_count(which == 0 ? c->g : c->d);
  return GANTTS_OK;
}

extern "C" int gantts_gan_step(const gantts_gan_step_t* c, int phases, const float* x, const float* y,
                               const int64_t* lengths_dev, float inv_frames, uint64_t seed, float* y_hat,
                               float* y_hat_static, float* losses_dev, void* workspace, size_t workspace_bytes,
                               void* stream) {
  int rc = check_step(c);
  if (rc) return rc;
  GANTTS_CHECK_ARG(x && y && lengths_dev && y_hat && y_hat_static && losses_dev, "gan_step: null pointer");
  size_t need = gantts_gan_step_workspace_bytes(c);
  if (!workspace || workspace_bytes < need) {
    set_error("gan_step: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GANTTS_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  StepLayout L;
  layout(c, reinterpret_cast<char*>(al256(reinterpret_cast<uintptr_t>(workspace))), &L);
  const int64_t M = (int64_t)c->B * c->T;
  const int Lg = c->g.num_layers, d_in = c->g.dims[0], d_out = c->g.dims[Lg];
  const int dD = c->d.dims[0], nS = c->n_static;
  const bool has_d = c->w_d > 0.f;
  const bool has_adv = has_d && c->adv_w > 0.f;
  // discriminator_linguistic_condition (train.py:254-256,302-303): D sees cat((x, y_adv), -1); the first
  // cond_w columns of both halves of d_in are copies of x, the gradient w.r.t. them is discarded.
  const int cond_w = (has_d && c->d_conditioned) ? d_in : 0;
  const int nA = dD - cond_w;
  ParamList pg, pd;
  param_list(c->g, c->g_sumW, c->g_sumb, L.g_grads, &pg);
  if (has_d) param_list(c->d, c->d_sumW, c->d_sumb, L.d_grads, &pd);
  ColList static_cols, adv_cols;
  static_cols.n = c->n_static_cols;
  for (int i = 0; i < c->n_static_cols; ++i) static_cols.c[i] = c->static_cols[i];
  adv_cols.n = has_d ? c->n_adv : 0;
  for (int i = 0; i < adv_cols.n; ++i) adv_cols.c[i] = c->adv_cols[i];
  gantts_mlp_t g = c->g, d = c->d;
  g.seed = gantts_gan_step_seed(seed, 0);

  if (phases & GANTTS_STEP_EVAL) {
    // ---- "test" phase of train.py:481-486 (model.eval(), phase != "train" at :273,:315): forwards and losses only
    GANTTS_CHECK_ARG(phases == GANTTS_STEP_EVAL, "gan_step: GANTTS_STEP_EVAL cannot be combined with training phases");
    g.dropout_p = 0.f;
    d.dropout_p = 0.f;
    if ((rc = gantts_sequence_mask(lengths_dev, L.mask, c->B, c->T, stream))) return rc;
    set_scales_kernel<<<1, 32, 0, st>>>(L.scal, inv_frames, has_adv ? c->adv_w : 0.f, c->mge_w, c->mse_w, 1);
    GANTTS_LAUNCH_CHECK("set_scales_kernel");
    gather_cols_list_kernel<<<blocks_1d(M * nS, 1024), 256, 0, st>>>(y, d_out, L.y_static, nS, static_cols, M);
    GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(y_static)");
    if ((rc = gantts_mlp_fwd(&g, x, d_in, M, y_hat, d_out, L.g_tape, L.g_tape_bytes, stream))) return rc;
    if ((rc = gantts_mlpg_fwd(y_hat, (int64_t)c->T * d_out, d_out, y_hat_static, (int64_t)c->T * nS, nS,
                              c->mlpg_table, &c->streams, &c->windows, c->B, c->T, stream)))
      return rc;
    if (has_d) {
      gather_cols_list_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(L.y_static, nS, L.d_in + cond_w, dD, adv_cols,
                                                                       M);
      GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(real)");
      gather_cols_list_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(y_hat_static, nS, L.d_in + M * dD + cond_w, dD,
                                                                       adv_cols, M);
      GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(fake)");
      if (cond_w) {
        GANTTS_CUDA(cudaMemcpy2DAsync(L.d_in, (size_t)dD * sizeof(float), x, (size_t)d_in * sizeof(float),
                                      (size_t)cond_w * sizeof(float), (size_t)M, cudaMemcpyDeviceToDevice, st));
        GANTTS_CUDA(cudaMemcpy2DAsync(L.d_in + M * dD, (size_t)dD * sizeof(float), x, (size_t)d_in * sizeof(float),
                                      (size_t)cond_w * sizeof(float), (size_t)M, cudaMemcpyDeviceToDevice, st));
      }
      if ((rc = gantts_mlp_fwd(&d, L.d_in, dD, 2 * M, L.d_out, 1, L.d_tape, L.d_tape_bytes, stream))) return rc;
      if ((rc = gantts_masked_bce_fwd(L.d_out, L.mask, M, 0, L.scal + S_REAL, L.red_ws, L.red_ws_bytes, stream))) return rc;
      if ((rc = gantts_masked_bce_fwd(L.d_out + M, L.mask, M, 1, L.scal + S_FAKE, L.red_ws, L.red_ws_bytes, stream))) return rc;
      if (has_adv &&
          (rc = gantts_masked_bce_fwd(L.d_out + M, L.mask, M, 0, L.scal + S_ADV, L.red_ws, L.red_ws_bytes, stream)))
        return rc;
    }
    if ((rc = gantts_masked_sse_fwd(y_hat_static, nS, L.y_static, nS, L.mask, M, nS, L.scal + S_MGE, L.red_ws,
                                    L.red_ws_bytes, stream)))
      return rc;
    if ((rc = gantts_masked_sse_fwd(y_hat, d_out, y, d_out, L.mask, M, d_out, L.scal + S_MSE, L.red_ws,
                                    L.red_ws_bytes, stream)))
      return rc;
    finalize_losses_kernel<<<1, 32, 0, st>>>(L.scal, losses_dev, has_adv ? c->adv_w : 0.f, c->mge_w, c->mse_w,
                                             has_d ? 1 : 0);
    GANTTS_LAUNCH_CHECK("finalize_losses_kernel");
    return GANTTS_OK;
  }

  if (phases & 1) {
    // ---- prologue: mask, scales, y_static (train.py:528-535)
    if ((rc = gantts_sequence_mask(lengths_dev, L.mask, c->B, c->T, stream))) return rc;
    set_scales_kernel<<<1, 32, 0, st>>>(L.scal, inv_frames, has_adv ? c->adv_w : 0.f, c->mge_w, c->mse_w, 0);
    GANTTS_LAUNCH_CHECK("set_scales_kernel");
    gather_cols_list_kernel<<<blocks_1d(M * nS, 1024), 256, 0, st>>>(y, d_out, L.y_static, nS, static_cols, M);
    GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(y_static)");
    GANTTS_CUDA(cudaMemsetAsync(L.g_static, 0, (size_t)M * nS * sizeof(float), st));
    // ---- apply_generator (train.py:336-355): G forward + MLPG
    if ((rc = gantts_mlp_fwd(&g, x, d_in, M, y_hat, d_out, L.g_tape, L.g_tape_bytes, stream))) return rc;
    if ((rc = gantts_mlpg_fwd(y_hat, (int64_t)c->T * d_out, d_out, y_hat_static, (int64_t)c->T * nS, nS,
                              c->mlpg_table, &c->streams, &c->windows, c->B, c->T, stream)))
      return rc;
    if (has_d) {
      // ---- update_discriminator (train.py:245-279): stacked real | fake batch of 2M rows
      gather_cols_list_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(L.y_static, nS, L.d_in + cond_w, dD, adv_cols,
                                                                       M);
      GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(real)");
      gather_cols_list_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(y_hat_static, nS, L.d_in + M * dD + cond_w, dD,
                                                                       adv_cols, M);
      GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(fake)");
      if (cond_w) {
        GANTTS_CUDA(cudaMemcpy2DAsync(L.d_in, (size_t)dD * sizeof(float), x, (size_t)d_in * sizeof(float),
                                      (size_t)cond_w * sizeof(float), (size_t)M, cudaMemcpyDeviceToDevice, st));
        GANTTS_CUDA(cudaMemcpy2DAsync(L.d_in + M * dD, (size_t)dD * sizeof(float), x, (size_t)d_in * sizeof(float),
                                      (size_t)cond_w * sizeof(float), (size_t)M, cudaMemcpyDeviceToDevice, st));
      }
      d.seed = gantts_gan_step_seed(seed, 1);
      if ((rc = gantts_mlp_fwd(&d, L.d_in, dD, 2 * M, L.d_out, 1, L.d_tape, L.d_tape_bytes, stream))) return rc;
      if ((rc = gantts_masked_bce_fwd(L.d_out, L.mask, M, 0, L.scal + S_REAL, L.red_ws, L.red_ws_bytes, stream))) return rc;
      if ((rc = gantts_masked_bce_fwd(L.d_out + M, L.mask, M, 1, L.scal + S_FAKE, L.red_ws, L.red_ws_bytes, stream))) return rc;
      if ((rc = gantts_masked_bce_bwd(L.d_out, L.mask, M, 0, L.scal + S_INV_T, L.g_dout, stream))) return rc;
      if ((rc = gantts_masked_bce_bwd(L.d_out + M, L.mask, M, 1, L.scal + S_INV_T, L.g_dout + M, stream))) return rc;
      // loss_d.backward(): D parameter gradients + gradient w.r.t. the (fake) D input
      // (input gradient for the fake half only: rows M..2M-1)
      if ((rc = mlp_bwd_impl(&d, L.g_dout, 1, L.d_out, 1, 2 * M, L.d_tape, L.d_tape_bytes, L.g_din, dD, M, pd.gW,
                             pd.gb, 0, L.mlp_ws, L.mlp_ws_bytes, stream)))
        return rc;
      scatter_cols_list_add_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(L.g_din + M * dD + cond_w, dD, L.g_static,
                                                                            nS, adv_cols, M);
      GANTTS_LAUNCH_CHECK("scatter_cols_list_add_kernel(fake)");
    }
  }
  if (phases & 2) {
    if (has_d) {
      // ---- clip_grad_norm_ + Adagrad on D (train.py:275-276)
      if ((rc = gantts_grad_sumsq(pd.g, pd.sizes, pd.n, L.scal + S_DSUMSQ, L.red_ws, L.red_ws_bytes, stream))) return rc;
      if ((rc = gantts_clip_adagrad_step(pd.p, pd.g, pd.s, pd.sizes, pd.n, L.scal + S_DSUMSQ, c->max_norm, c->lr_d,
                                         c->wd_d, c->eps, stream)))
        return rc;
    }
    // ---- update_generator (train.py:282-320)
    if ((rc = gantts_masked_sse_fwd(y_hat_static, nS, L.y_static, nS, L.mask, M, nS, L.scal + S_MGE, L.red_ws,
                                    L.red_ws_bytes, stream)))
      return rc;
    if ((rc = gantts_masked_sse_fwd(y_hat, d_out, y, d_out, L.mask, M, d_out, L.scal + S_MSE, L.red_ws,
                                    L.red_ws_bytes, stream)))
      return rc;
    if (has_adv) {
      // third D forward: updated weights, fresh dropout mask (train.py:307)
      d.seed = gantts_gan_step_seed(seed, 2);
      if ((rc = gantts_mlp_fwd(&d, L.d_in + M * dD, dD, M, L.d_out, 1, L.d_tape, L.d_tape_bytes, stream))) return rc;
      if ((rc = gantts_masked_bce_fwd(L.d_out, L.mask, M, 0, L.scal + S_ADV, L.red_ws, L.red_ws_bytes, stream))) return rc;
      if ((rc = gantts_masked_bce_bwd(L.d_out, L.mask, M, 0, L.scal + S_ADV_SCALE, L.g_dout, stream))) return rc;
      if ((rc = gantts_mlp_bwd(&d, L.g_dout, 1, L.d_out, 1, M, L.d_tape, L.d_tape_bytes, L.g_din, dD, nullptr,
                               nullptr, 0, L.mlp_ws, L.mlp_ws_bytes, stream)))
        return rc;
      scatter_cols_list_add_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(L.g_din + cond_w, dD, L.g_static, nS,
                                                                            adv_cols, M);
      GANTTS_LAUNCH_CHECK("scatter_cols_list_add_kernel(adv)");
    }
    if (c->mge_w != 0.f) {
      if ((rc = gantts_masked_sse_bwd(y_hat_static, nS, L.y_static, nS, L.mask, M, nS, L.scal + S_MGE_SCALE,
                                      L.g_static, nS, 1, stream)))
        return rc;
    }
    // ---- loss_g.backward(): MLPG backward + generator backward on the summed upstream gradient
    if ((rc = gantts_mlpg_bwd(L.g_static, (int64_t)c->T * nS, nS, L.g_yhat, (int64_t)c->T * d_out, d_out,
                              c->mlpg_table, &c->streams, &c->windows, c->B, c->T, 0, stream)))
      return rc;
    if (c->mse_w != 0.f) {
      if ((rc = gantts_masked_sse_bwd(y_hat, d_out, y, d_out, L.mask, M, d_out, L.scal + S_MSE_SCALE, L.g_yhat,
                                      d_out, 1, stream)))
        return rc;
    }
    if ((rc = gantts_mlp_bwd(&g, L.g_yhat, d_out, nullptr, 0, M, L.g_tape, L.g_tape_bytes, nullptr, 0, pg.gW, pg.gb,
                             0, L.mlp_ws, L.mlp_ws_bytes, stream)))
      return rc;
  }
  if (phases & 4) {
    // ---- clip_grad_norm_ + Adagrad on G (train.py:317-318), then the loss scalars
    if ((rc = gantts_grad_sumsq(pg.g, pg.sizes, pg.n, L.scal + S_GSUMSQ, L.red_ws, L.red_ws_bytes, stream))) return rc;
    if ((rc = gantts_clip_adagrad_step(pg.p, pg.g, pg.s, pg.sizes, pg.n, L.scal + S_GSUMSQ, c->max_norm, c->lr_g,
                                       c->wd_g, c->eps, stream)))
      return rc;
    finalize_losses_kernel<<<1, 32, 0, st>>>(L.scal, losses_dev, has_adv ? c->adv_w : 0.f, c->mge_w, c->mse_w,
                                             has_d ? 1 : 0);
    GANTTS_LAUNCH_CHECK("finalize_losses_kernel");
  }
  return GANTTS_OK;
}
