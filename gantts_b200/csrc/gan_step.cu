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
#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace gantts {

// NVTX range per phase of the step (header-only NVTX3: a no-op unless a profiler is attached; `ncu --nvtx` and nsys
// show the phases of one gantts_gan_step call on the timeline).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

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

// Column gather of up to two row blocks straight into bf16 hi/lo operand planes (the discriminator's input: rows
// [0, rows_a) = selected columns of `a`, rows [rows_a, rows_a + rows_b) = selected columns of `b`): replaces
// gather -> fp32 matrix -> split_planes.  One warp per row, a lane converts PAIRS of columns (4-byte stores).
__global__ void gather_planes_kernel(const float* __restrict__ a, int64_t a_rs, ColList ca, int64_t rows_a,
                                     const float* __restrict__ b, int64_t b_rs, ColList cb, int64_t rows_b,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t pitch) {
  pdl_entry();
  __shared__ int sa[GANTTS_MAX_COLS], sb[GANTTS_MAX_COLS];
  for (int i = threadIdx.x; i < ca.n; i += blockDim.x) sa[i] = ca.c[i];
  for (int i = threadIdx.x; i < cb.n; i += blockDim.x) sb[i] = cb.c[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows_a + rows_b; r += nwarps) {
    const bool first = r < rows_a;
    const float* src = first ? a + r * a_rs : b + (r - rows_a) * b_rs;
    const int* sc = first ? sa : sb;
    const int n = first ? ca.n : cb.n;
    uint32_t* hr = reinterpret_cast<uint32_t*>(hi + r * pitch);
    uint32_t* lr = reinterpret_cast<uint32_t*>(lo + r * pitch);
    for (int c = 2 * lane; c < n; c += 64) {
      const float v0 = src[sc[c]], v1 = (c + 1 < n) ? src[sc[c + 1]] : 0.f;
      const uint32_t hp = pack_bf16x2(v0, v1);
      hr[c >> 1] = hp;
      lr[c >> 1] = pack_bf16x2(v0 - __uint_as_float(hp << 16), v1 - __uint_as_float(hp & 0xffff0000u));
    }
  }
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

// inv_frames <= 0: derive the normaliser on the device, 1 / sum_b min(len_b, T) (= mask.sum() of train.py:258,286) --
// single-process use; a data-parallel caller passes 1 / (GLOBAL number of valid frames).
__global__ void set_scales_kernel(float* scal, float inv_frames, float adv_w, float mge_w, float mse_w,
                                  int zero_norms, const int64_t* __restrict__ lengths, int B, int T) {
  pdl_entry();
  if (threadIdx.x == 0) {
    if (zero_norms) scal[S_DSUMSQ] = scal[S_GSUMSQ] = 0.f;
    if (!(inv_frames > 0.f)) {
      int64_t n = 0;
      for (int b = 0; b < B; ++b) {
        const int64_t l = lengths[b];
        n += l < 0 ? 0 : (l > T ? T : l);
      }
      inv_frames = n > 0 ? 1.f / (float)n : 0.f;
    }
    scal[S_INV_T] = inv_frames;
    scal[S_ADV_SCALE] = adv_w * inv_frames;
    scal[S_MGE_SCALE] = mge_w * inv_frames;
    scal[S_MSE_SCALE] = mse_w * inv_frames;
  }
}

// Deferred reductions: every loss kernel of the step leaves per-block partial sums in its own slot; the single
// finalize kernel at the end of the step reduces all of them (deterministic: fixed block order) -- no per-loss
// "finish" launch on the way.
enum RedSlot { R_REAL = 0, R_FAKE = 1, R_ADV = 2, R_MGE = 3, R_MSE = 4, R_COUNT = 5 };

struct RedCounts {
  int n[R_COUNT];      // blocks that wrote partials into slot i (0 = slot unused this step)
};

// Adversarial BCE terms of train.py:262-270,307-308 for one or two halves of a stacked discriminator output, forward
// sums AND the gradient w.r.t. D in one pass: half 0 = rows [0, M) with kind0, half 1 = rows [M, 2M) with kind1
// (kind 0: -log(D + eps) * m, correct = D > 0.5; kind 1: -log(1 - D + eps) * m, correct = D < 0.5).  mask is [M] for
// both halves.  Blocks [0, nbh) serve half 0 and write ws0, blocks [nbh, 2 nbh) serve half 1 and write ws1.
__global__ void __launch_bounds__(RED_THREADS)
bce_fwd_bwd_kernel(const float* __restrict__ Dv, const float* __restrict__ mask, int64_t M, int kind0, int kind1,
                   int nbh, const float* __restrict__ scale, float* __restrict__ gD, RedWs* ws0, RedWs* ws1) {
  pdl_entry();
  __shared__ float sm[RED_NV * 32];
  const int half = blockIdx.x >= nbh ? 1 : 0;
  const int kind = half ? kind1 : kind0;
  const int blk = blockIdx.x - half * nbh;
  const float* d = Dv + (int64_t)half * M;
  float* g = gD ? gD + (int64_t)half * M : nullptr;
  const float s = scale[0];
  float v[RED_NV] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t i = (int64_t)blk * RED_THREADS + threadIdx.x; i < M; i += (int64_t)nbh * RED_THREADS) {
    const float dv = d[i], m = mask[i];
    const float arg = kind == 0 ? (dv + 1e-20f) : (1.f - dv + 1e-20f);
    v[0] -= logf(arg) * m;
    const bool hit = kind == 0 ? (dv > 0.5f) : (dv < 0.5f);
    v[1] += hit ? m : 0.f;
    v[2] += m;
    if (g) g[i] = kind == 0 ? (-s * m / (dv + 1e-20f)) : (s * m / (1.f - dv + 1e-20f));
  }
  block_sum<RED_NV>(v, sm);
  if (threadIdx.x == 0) {
    RedWs* ws = half ? ws1 : ws0;
#pragma unroll
    for (int k = 0; k < RED_NV; ++k) ws->partial[blk][k] = v[k];
  }
}

// MaskedMSELoss forward sums (gantts/seqloss.py:41-43) AND its gradient 2 * scale * (a m - b m) * m in one pass
// (ga == nullptr: forward only).  The gradient is STORED (not accumulated): this launch initialises the buffer.
// bmap.n > 0: column d of the target is column bmap.c[d] of `b` (the static features are read straight out of y:
// get_static_features of multistream.py:56-79 without materialising y_static).
__global__ void __launch_bounds__(RED_THREADS)
sse_fwd_bwd_kernel(const float* __restrict__ a, int64_t a_rs, const float* __restrict__ b, int64_t b_rs,
                   const float* __restrict__ mask, int64_t rows, int D, const float* __restrict__ scale,
                   float* __restrict__ ga, int64_t ga_rs, RedWs* ws, ColList bmap) {
  pdl_entry();
  __shared__ float sm[RED_NV * 32];
  __shared__ int sc[GANTTS_MAX_COLS];
  for (int i = threadIdx.x; i < bmap.n; i += RED_THREADS) sc[i] = bmap.c[i];
  __syncthreads();
  const bool mapped = bmap.n > 0;
  float v[RED_NV] = {0.f, 0.f, 0.f, 0.f};
  const float s2 = ga ? 2.f * scale[0] : 0.f;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * RED_THREADS + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * RED_THREADS) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float m = mask[r];
    const float* ar = a + r * a_rs;
    const float* br = b + r * b_rs;
#pragma unroll 4
    for (int d = lane; d < D; d += 32) {
      const float x = ar[d] * m - br[mapped ? sc[d] : d] * m;
      v[0] = fmaf(x, x, v[0]);
      if (ga) ga[r * ga_rs + d] = s2 * x * m;
    }
    if (lane == 0) v[1] += m;
  }
  block_sum<RED_NV>(v, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < RED_NV; ++k) ws->partial[blockIdx.x][k] = v[k];
  }
}

// clip_grad_norm_ + Adagrad with the sum of squares taken from the per-block partials of sumsq_partial_kernel:
// every block re-reduces the (<= 592) partials itself in the same fixed order, which removes the finish launch.
__global__ void __launch_bounds__(OPT_THREADS)
clip_adagrad_partials_kernel(TensorList tl, const float* __restrict__ partial, int npartial, float* __restrict__ sumsq_out,
                             float max_norm, float lr, float wd, float eps) {
  pdl_entry();
  __shared__ float sm[32];
  __shared__ float total_s;
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < npartial; i += OPT_THREADS) v[0] += partial[i];
  block_sum<1>(v, sm);
  if (threadIdx.x == 0) {
    total_s = v[0];
    if (blockIdx.x == 0) sumsq_out[0] = v[0];
  }
  __syncthreads();
  const float total_norm = sqrtf(total_s);
  float coef = max_norm / (total_norm + 1e-6f);
  coef = coef < 1.f ? coef : 1.f;
  const int64_t total = tl.off[tl.n];
  for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * OPT_THREADS) {
    int k = find_tensor(tl, i);
    int64_t j = i - tl.off[k];
    float g = tl.g[k][j] * coef;
    tl.g[k][j] = g;
    float p = tl.p[k][j];
    g = fmaf(wd, p, g);
    float s = fmaf(g, g, tl.s[k][j]);
    tl.s[k][j] = s;
    tl.p[k][j] = p - lr * g / (sqrtf(s) + eps);
  }
}

// The same with torch.optim.Adam (amsgrad off; reference hparams.py:125-130): tl.s = exp_avg, tl.s2 = exp_avg_sq;
// step_size = lr / (1 - beta1^t), inv_sqrt_bc2 = 1 / sqrt(1 - beta2^t) come from the host.
__global__ void __launch_bounds__(OPT_THREADS)
clip_adam_partials_kernel(TensorList tl, const float* __restrict__ partial, int npartial, float* __restrict__ sumsq_out,
                          float max_norm, float b1, float b2, float wd, float eps, float step_size, float inv_sqrt_bc2) {
  pdl_entry();
  __shared__ float sm[32];
  __shared__ float total_s;
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < npartial; i += OPT_THREADS) v[0] += partial[i];
  block_sum<1>(v, sm);
  if (threadIdx.x == 0) {
    total_s = v[0];
    if (blockIdx.x == 0) sumsq_out[0] = v[0];
  }
  __syncthreads();
  const float total_norm = sqrtf(total_s);
  float coef = max_norm / (total_norm + 1e-6f);
  coef = coef < 1.f ? coef : 1.f;
  const int64_t total = tl.off[tl.n];
  for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * OPT_THREADS) {
    int k = find_tensor(tl, i);
    int64_t j = i - tl.off[k];
    float g = tl.g[k][j] * coef;
    tl.g[k][j] = g;
    const float p = tl.p[k][j];
    g = fmaf(wd, p, g);
    const float m = b1 * tl.s[k][j] + (1.f - b1) * g;
    const float q = b2 * tl.s2[k][j] + (1.f - b2) * g * g;
    tl.s[k][j] = m;
    tl.s2[k][j] = q;
    tl.p[k][j] = p - step_size * m / (sqrtf(q) * inv_sqrt_bc2 + eps);
  }
}

// losses[0..11] = loss_d, loss_fake_d, loss_real_d, loss_mse, loss_mge, loss_adv, loss_g,
//                 real_correct, fake_correct, frames(local sum of mask), d_grad_norm, g_grad_norm
__global__ void __launch_bounds__(RED_THREADS)
finalize_losses_kernel(const float* scal, float* losses, const RedWs* red, RedCounts cnt, float adv_w, float mge_w,
                       float mse_w, int has_d) {
  pdl_entry();
  __shared__ float sm[RED_NV * 32];
  __shared__ float tot[R_COUNT][RED_NV];
  for (int sl = 0; sl < R_COUNT; ++sl) {
    float v[RED_NV] = {0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < cnt.n[sl]; i += RED_THREADS) {
#pragma unroll
      for (int k = 0; k < RED_NV; ++k) v[k] += red[sl].partial[i][k];
    }
    block_sum<RED_NV>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < RED_NV; ++k) tot[sl][k] = v[k];
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const float invT = scal[S_INV_T];
  const float real = has_d ? tot[R_REAL][0] * invT : 0.f, fake = has_d ? tot[R_FAKE][0] * invT : 0.f;
  const float adv = (has_d && adv_w > 0.f) ? tot[R_ADV][0] * invT : 0.f;
  const float mge = tot[R_MGE][0] * invT, mse = tot[R_MSE][0] * invT;
  losses[0] = real + fake;
  losses[1] = fake;
  losses[2] = real;
  losses[3] = mse;
  losses[4] = mge;
  losses[5] = adv;
  losses[6] = (mse_w * mse + mge_w * mge) + adv_w * adv;
  losses[7] = has_d ? tot[R_REAL][1] : 0.f;
  losses[8] = has_d ? tot[R_FAKE][1] : 0.f;
  losses[9] = tot[R_MGE][1];
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
  float* s2[2 * GANTTS_MAX_LAYERS];                     // Adam: exp_avg_sq (null for Adagrad)
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
  RedWs* red;             // [R_COUNT] deferred loss partials
  float* opt_partial;     // [OPT_MAX_BLOCKS] sum-of-squares partials of the model being stepped
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
  L->red = (RedWs*)take(R_COUNT * sizeof(RedWs));
  L->opt_partial = (float*)take(OPT_MAX_BLOCKS * sizeof(float));
  L->total = (size_t)(cur - base) + 256;
}

static void param_list(const gantts_mlp_t& m, float* const* sumW, float* const* sumb, float* const* sqW, float* const* sqb,
                       float* flat, ParamList* pl) {
  pl->n = 0;
  pl->total = 0;
  float* cur = flat;
  for (int l = 0; l < m.num_layers; ++l) {
    const int64_t nw = (int64_t)m.dims[l + 1] * m.dims[l], nb = m.dims[l + 1];
    pl->gW[l] = cur;
    pl->p[pl->n] = const_cast<float*>(m.W[l]);
    pl->g[pl->n] = cur;
    pl->s[pl->n] = sumW[l];
    pl->s2[pl->n] = sqW[l];
    pl->sizes[pl->n++] = nw;
    cur += nw;
    pl->gb[l] = cur;
    pl->p[pl->n] = const_cast<float*>(m.b[l]);
    pl->g[pl->n] = cur;
    pl->s[pl->n] = sumb[l];
    pl->s2[pl->n] = sqb[l];
    pl->sizes[pl->n++] = nb;
    cur += nb;
  }
  pl->total = cur - flat;
}

static inline int bce_blocks(int64_t rows) { return grid_for(rows, RED_THREADS); }
static inline int sse_blocks(int64_t rows, int D) { return grid_for(rows * D, RED_THREADS * 4); }

// BCE of a stacked discriminator output: halves = 2 -> rows [0,M) kind0 into slot0 and rows [M,2M) kind1 into slot1
static int launch_bce(const float* Dv, const float* mask, int64_t M, int halves, int kind0, int kind1, const float* scale,
                      float* gD, RedWs* ws0, RedWs* ws1, cudaStream_t st) {
  const int nbh = bce_blocks(M);
  GANTTS_PDL_LAUNCH((bce_fwd_bwd_kernel), nbh * halves, RED_THREADS, 0, st, Dv, mask, M, kind0, kind1, nbh, scale, gD, ws0, ws1);
  GANTTS_LAUNCH_CHECK("bce_fwd_bwd_kernel");
  return GANTTS_OK;
}

static int launch_sse(const float* a, int64_t a_rs, const float* b, int64_t b_rs, const float* mask, int64_t rows, int D,
                      const float* scale, float* ga, int64_t ga_rs, RedWs* ws, cudaStream_t st,
                      const ColList* bmap = nullptr) {
  ColList none;
  none.n = 0;
  GANTTS_PDL_LAUNCH((sse_fwd_bwd_kernel), sse_blocks(rows, D), RED_THREADS, 0, st, a, a_rs, b, b_rs, mask, rows, D, scale, ga, ga_rs, ws,
                                                                 bmap ? *bmap : none);
  GANTTS_LAUNCH_CHECK("sse_fwd_bwd_kernel");
  return GANTTS_OK;
}

// clip_grad_norm_ + optimiser step over one model's parameter list: two launches (partials, update), no finish kernel
static int clip_opt_model(const gantts_gan_step_t* c, const ParamList& pl, float* partial, float* sumsq_out, float lr, float wd,
                          cudaStream_t st) {
  TensorList tl;
  const bool adam = c->optimizer == GANTTS_OPT_ADAM;
  int rc = fill(tl, pl.p, pl.g, pl.s, adam ? pl.s2 : nullptr, pl.sizes, 0, pl.n);
  if (rc) return rc;
  const int nb = blocks_for(tl.off[tl.n], OPT_MAX_BLOCKS);
  GANTTS_PDL_LAUNCH((sumsq_partial_kernel), nb, OPT_THREADS, 0, st, tl, partial);
  GANTTS_LAUNCH_CHECK("sumsq_partial_kernel");
  if (adam) {
    for (int i = 0; i < pl.n; ++i) GANTTS_CHECK_ARG(pl.s[i] && pl.s2[i], "gan_step: Adam needs exp_avg and exp_avg_sq for every tensor");
    GANTTS_CHECK_ARG(c->opt_step >= 1, "gan_step: Adam needs opt_step >= 1 (the number of the step being taken)");
    const double t = (double)c->opt_step;
    const float step_size = (float)((double)lr / (1.0 - pow((double)c->beta1, t)));
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)c->beta2, t)));
    GANTTS_PDL_LAUNCH((clip_adam_partials_kernel), nb, OPT_THREADS, 0, st, tl, partial, nb, sumsq_out, c->max_norm, c->beta1, c->beta2,
                      wd, c->eps, step_size, inv_sqrt_bc2);
    GANTTS_LAUNCH_CHECK("clip_adam_partials_kernel");
  } else {
    GANTTS_PDL_LAUNCH((clip_adagrad_partials_kernel), nb, OPT_THREADS, 0, st, tl, partial, nb, sumsq_out, c->max_norm, lr, wd, c->eps);
    GANTTS_LAUNCH_CHECK("clip_adagrad_partials_kernel");
  }
  return GANTTS_OK;
}

static int check_step(const gantts_gan_step_t* c) {
  GANTTS_CHECK_ARG(c, "gan_step: null config");
  GANTTS_CHECK_ARG(c->B >= 1 && c->T >= 1, "gan_step: bad batch shape");
  GANTTS_CHECK_ARG(c->optimizer == GANTTS_OPT_ADAGRAD || c->optimizer == GANTTS_OPT_ADAM, "gan_step: unknown optimizer %d", c->optimizer);
  if (c->optimizer == GANTTS_OPT_ADAM)
    GANTTS_CHECK_ARG(c->beta1 >= 0.f && c->beta1 < 1.f && c->beta2 >= 0.f && c->beta2 < 1.f, "gan_step: Adam betas out of range");
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
  param_list(c->g, c->g_sumW, c->g_sumb, c->g_sqW, c->g_sqb, L.g_grads, &pg);
  if (has_d) param_list(c->d, c->d_sumW, c->d_sumb, c->d_sqW, c->d_sqb, L.d_grads, &pd);
  ColList static_cols, adv_cols;
  static_cols.n = c->n_static_cols;
  for (int i = 0; i < c->n_static_cols; ++i) static_cols.c[i] = c->static_cols[i];
  adv_cols.n = has_d ? c->n_adv : 0;
  for (int i = 0; i < adv_cols.n; ++i) adv_cols.c[i] = c->adv_cols[i];
  // the adversarial columns are one contiguous window of y_hat_static (mgc with the first coefficients masked, the
  // hparams case): the discriminator's input gradient can be accumulated in place
  bool adv_window = has_d && !cond_w && adv_cols.n >= 1;
  for (int i = 1; i < adv_cols.n; ++i) adv_window = adv_window && adv_cols.c[i] == adv_cols.c[0] + i;
  ColList real_cols;          // adversarial columns taken from y directly: static_cols o adv_cols
  real_cols.n = adv_cols.n;
  for (int i = 0; i < adv_cols.n; ++i) {
    GANTTS_CHECK_ARG(adv_cols.c[i] >= 0 && adv_cols.c[i] < c->n_static_cols, "gan_step: adversarial column out of range");
    real_cols.c[i] = static_cols.c[adv_cols.c[i]];
  }
  gantts_mlp_t g = c->g, d = c->d;
  g.seed = gantts_gan_step_seed(seed, 0);

  RedCounts cnt{};
  if (phases & GANTTS_STEP_EVAL) {
    NvtxRange r_eval("gantts_gan_step/eval");
    // ---- "test" phase of train.py:481-486 (model.eval(), phase != "train" at :273,:315): forwards and losses only
    GANTTS_CHECK_ARG(phases == GANTTS_STEP_EVAL, "gan_step: GANTTS_STEP_EVAL cannot be combined with training phases");
    g.dropout_p = 0.f;
    d.dropout_p = 0.f;
    if ((rc = gantts_sequence_mask(lengths_dev, L.mask, c->B, c->T, stream))) return rc;
    GANTTS_PDL_LAUNCH((set_scales_kernel), 1, 32, 0, st, L.scal, inv_frames, has_adv ? c->adv_w : 0.f, c->mge_w, c->mse_w, 1, lengths_dev, c->B, c->T);
    GANTTS_LAUNCH_CHECK("set_scales_kernel");
    if (has_d) {      // (the eval path keeps the two-step gather of the discriminator input)
      gather_cols_list_kernel<<<blocks_1d(M * nS, 1024), 256, 0, st>>>(y, d_out, L.y_static, nS, static_cols, M);
      GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(y_static)");
    }
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
      if ((rc = launch_bce(L.d_out, L.mask, M, 2, 0, 1, L.scal + S_INV_T, nullptr, &L.red[R_REAL], &L.red[R_FAKE], st)))
        return rc;
      cnt.n[R_REAL] = cnt.n[R_FAKE] = bce_blocks(M);
      if (has_adv) {
        if ((rc = launch_bce(L.d_out + M, L.mask, M, 1, 0, 0, L.scal + S_INV_T, nullptr, &L.red[R_ADV], nullptr, st)))
          return rc;
        cnt.n[R_ADV] = bce_blocks(M);
      }
    }
    if ((rc = launch_sse(y_hat_static, nS, y, d_out, L.mask, M, nS, L.scal + S_MGE_SCALE, nullptr, 0, &L.red[R_MGE], st,
                         &static_cols)))
      return rc;
    if ((rc = launch_sse(y_hat, d_out, y, d_out, L.mask, M, d_out, L.scal + S_MSE_SCALE, nullptr, 0, &L.red[R_MSE], st)))
      return rc;
    cnt.n[R_MGE] = sse_blocks(M, nS);
    cnt.n[R_MSE] = sse_blocks(M, d_out);
    GANTTS_PDL_LAUNCH((finalize_losses_kernel), 1, RED_THREADS, 0, st, L.scal, losses_dev, L.red, cnt, has_adv ? c->adv_w : 0.f, c->mge_w,
                                                      c->mse_w, has_d ? 1 : 0);
    GANTTS_LAUNCH_CHECK("finalize_losses_kernel");
    return GANTTS_OK;
  }

  if (phases & 1) {
    NvtxRange r1("gantts_gan_step/phase1: G fwd, MLPG, MGE, D fwd+bwd");
    // ---- prologue: mask, scales, y_static (train.py:528-535)
    if ((rc = gantts_sequence_mask(lengths_dev, L.mask, c->B, c->T, stream))) return rc;
    GANTTS_PDL_LAUNCH((set_scales_kernel), 1, 32, 0, st, L.scal, inv_frames, has_adv ? c->adv_w : 0.f, c->mge_w, c->mse_w, 0, lengths_dev, c->B, c->T);
    GANTTS_LAUNCH_CHECK("set_scales_kernel");
    if (has_d && cond_w) {      // only the conditioned-discriminator fallback still gathers from y_static
      gather_cols_list_kernel<<<blocks_1d(M * nS, 1024), 256, 0, st>>>(y, d_out, L.y_static, nS, static_cols, M);
      GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(y_static)");
    }
    // ---- apply_generator (train.py:336-355): G forward + MLPG
    if ((rc = gantts_mlp_fwd(&g, x, d_in, M, y_hat, d_out, L.g_tape, L.g_tape_bytes, stream))) return rc;
    if ((rc = gantts_mlpg_fwd(y_hat, (int64_t)c->T * d_out, d_out, y_hat_static, (int64_t)c->T * nS, nS,
                              c->mlpg_table, &c->streams, &c->windows, c->B, c->T, stream)))
      return rc;
    // MGE loss (train.py:291) and its gradient in one pass; the gradient INITIALISES g_static, the two discriminator
    // passes then accumulate their input gradients on top of it
    if ((rc = launch_sse(y_hat_static, nS, y, d_out, L.mask, M, nS, L.scal + S_MGE_SCALE, L.g_static, nS,
                         &L.red[R_MGE], st, &static_cols)))
      return rc;
    if (has_d) {
      // ---- update_discriminator (train.py:245-279): stacked real | fake batch of 2M rows
      d.seed = gantts_gan_step_seed(seed, 1);
      if (!cond_w) {
        // selected columns of y (real) and y_hat_static (fake) straight into the discriminator's input planes
        Planes din;
        if ((rc = mlp_tape_input_planes(&d, 2 * M, L.d_tape, L.d_tape_bytes, &din))) return rc;
        GANTTS_PDL_LAUNCH((gather_planes_kernel), blocks_1d(2 * M * nA, 1024), 256, 0, st, y, d_out, real_cols, M, y_hat_static, nS,
                                                                          adv_cols, M, din.hi, din.lo, din.pitch);
        GANTTS_LAUNCH_CHECK("gather_planes_kernel(real|fake)");
        if ((rc = mlp_fwd_impl(&d, nullptr, 0, 2 * M, L.d_out, 1, L.d_tape, L.d_tape_bytes, stream, true))) return rc;
      } else {
        gather_cols_list_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(L.y_static, nS, L.d_in + cond_w, dD, adv_cols,
                                                                         M);
        GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(real)");
        gather_cols_list_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(y_hat_static, nS, L.d_in + M * dD + cond_w, dD,
                                                                         adv_cols, M);
        GANTTS_LAUNCH_CHECK("gather_cols_list_kernel(fake)");
        GANTTS_CUDA(cudaMemcpy2DAsync(L.d_in, (size_t)dD * sizeof(float), x, (size_t)d_in * sizeof(float),
                                      (size_t)cond_w * sizeof(float), (size_t)M, cudaMemcpyDeviceToDevice, st));
        GANTTS_CUDA(cudaMemcpy2DAsync(L.d_in + M * dD, (size_t)dD * sizeof(float), x, (size_t)d_in * sizeof(float),
                                      (size_t)cond_w * sizeof(float), (size_t)M, cudaMemcpyDeviceToDevice, st));
        if ((rc = gantts_mlp_fwd(&d, L.d_in, dD, 2 * M, L.d_out, 1, L.d_tape, L.d_tape_bytes, stream))) return rc;
      }
      // real and fake BCE terms, counts and dL/dD of both halves (train.py:262-270) in one launch
      if ((rc = launch_bce(L.d_out, L.mask, M, 2, 0, 1, L.scal + S_INV_T, L.g_dout, &L.red[R_REAL], &L.red[R_FAKE], st)))
        return rc;
      // loss_d.backward(): D parameter gradients + gradient w.r.t. the (fake) D input (rows M..2M-1 only).  When the
      // adversarial columns form one window of y_hat_static the last GEMM adds its result straight into g_static
      // (the scatter of the column gather's backward); otherwise it goes to g_din and a scatter kernel follows.
      if (adv_window) {
        float* win = L.g_static + adv_cols.c[0] - M * (int64_t)nS;      // row r of the stacked batch -> g_static[r - M]
        if ((rc = mlp_bwd_impl(&d, L.g_dout, 1, L.d_out, 1, 2 * M, L.d_tape, L.d_tape_bytes, win, nS, M, pd.gW, pd.gb, 0,
                               L.mlp_ws, L.mlp_ws_bytes, stream, 1)))
          return rc;
      } else {
        if ((rc = mlp_bwd_impl(&d, L.g_dout, 1, L.d_out, 1, 2 * M, L.d_tape, L.d_tape_bytes, L.g_din, dD, M, pd.gW,
                               pd.gb, 0, L.mlp_ws, L.mlp_ws_bytes, stream)))
          return rc;
        scatter_cols_list_add_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(L.g_din + M * dD + cond_w, dD, L.g_static,
                                                                              nS, adv_cols, M);
        GANTTS_LAUNCH_CHECK("scatter_cols_list_add_kernel(fake)");
      }
    }
  }
  if (phases & 2) {
    NvtxRange r2("gantts_gan_step/phase2: D step, adv D fwd+bwd, MLPG bwd, G bwd");
    if (has_d) {
      // ---- clip_grad_norm_ + Adagrad on D (train.py:275-276)
      if ((rc = clip_opt_model(c, pd, L.opt_partial, L.scal + S_DSUMSQ, c->lr_d, c->wd_d, st)))
        return rc;
    }
    // ---- update_generator (train.py:282-320); the MGE term was evaluated in phase 1
    if (has_adv) {
      // third D forward: updated weights, fresh dropout mask (train.py:307)
      d.seed = gantts_gan_step_seed(seed, 2);
      if (!cond_w) {
        Planes din;
        if ((rc = mlp_tape_input_planes(&d, M, L.d_tape, L.d_tape_bytes, &din))) return rc;
        ColList none;
        none.n = 0;
        GANTTS_PDL_LAUNCH((gather_planes_kernel), blocks_1d(M * nA, 1024), 256, 0, st, y_hat_static, nS, adv_cols, M, nullptr, 0, none, 0,
                                                                      din.hi, din.lo, din.pitch);
        GANTTS_LAUNCH_CHECK("gather_planes_kernel(adv)");
        if ((rc = mlp_fwd_impl(&d, nullptr, 0, M, L.d_out, 1, L.d_tape, L.d_tape_bytes, stream, true))) return rc;
      } else if ((rc = gantts_mlp_fwd(&d, L.d_in + M * dD, dD, M, L.d_out, 1, L.d_tape, L.d_tape_bytes, stream))) {
        return rc;
      }
      if ((rc = launch_bce(L.d_out, L.mask, M, 1, 0, 0, L.scal + S_ADV_SCALE, L.g_dout, &L.red[R_ADV], nullptr, st)))
        return rc;
      if (adv_window) {
        if ((rc = mlp_bwd_impl(&d, L.g_dout, 1, L.d_out, 1, M, L.d_tape, L.d_tape_bytes, L.g_static + adv_cols.c[0], nS, 0,
                               nullptr, nullptr, 0, L.mlp_ws, L.mlp_ws_bytes, stream, 1)))
          return rc;
      } else {
        if ((rc = gantts_mlp_bwd(&d, L.g_dout, 1, L.d_out, 1, M, L.d_tape, L.d_tape_bytes, L.g_din, dD, nullptr,
                                 nullptr, 0, L.mlp_ws, L.mlp_ws_bytes, stream)))
          return rc;
        scatter_cols_list_add_kernel<<<blocks_1d(M * nA, 1024), 256, 0, st>>>(L.g_din + cond_w, dD, L.g_static, nS,
                                                                              adv_cols, M);
        GANTTS_LAUNCH_CHECK("scatter_cols_list_add_kernel(adv)");
      }
    }
    // ---- loss_g.backward(): MSE term (train.py:294) + MLPG backward + generator backward on the summed gradient.
    // With mse_w != 0 the MSE pass stores its gradient into g_yhat and the MLPG backward accumulates on top of it.
    if ((rc = launch_sse(y_hat, d_out, y, d_out, L.mask, M, d_out, L.scal + S_MSE_SCALE, c->mse_w != 0.f ? L.g_yhat : nullptr,
                         d_out, &L.red[R_MSE], st)))
      return rc;
    // mse_w == 0 (the CLI default, train.py:15): nothing else adds to dL/dy_hat, so the MLPG backward writes the
    // operand planes of the generator's backward GEMMs directly (no fp32 matrix, no conversion pass)
    bool direct = false;
    if (c->mse_w == 0.f) {
      Planes gp;
      if ((rc = mlp_bwd_gy_planes(&g, M, L.mlp_ws, L.mlp_ws_bytes, &gp))) return rc;
      rc = mlpg_bwd_planes(L.g_static, (int64_t)c->T * nS, nS, gp.hi, gp.lo, gp.pitch, c->mlpg_table, &c->streams,
                           &c->windows, c->B, c->T, stream);
      if (rc == GANTTS_OK) direct = true;
      else if (rc != GANTTS_E_UNSUPPORTED) return rc;
    }
    if (!direct &&
        (rc = gantts_mlpg_bwd(L.g_static, (int64_t)c->T * nS, nS, L.g_yhat, (int64_t)c->T * d_out, d_out, c->mlpg_table,
                              &c->streams, &c->windows, c->B, c->T, c->mse_w != 0.f ? 1 : 0, stream)))
      return rc;
    if ((rc = mlp_bwd_impl(&g, direct ? nullptr : L.g_yhat, d_out, nullptr, 0, M, L.g_tape, L.g_tape_bytes, nullptr, 0, 0,
                           pg.gW, pg.gb, 0, L.mlp_ws, L.mlp_ws_bytes, stream, -1, direct)))
      return rc;
  }
  if (phases & 4) {
    NvtxRange r4("gantts_gan_step/phase4: G step, losses");
    // ---- clip_grad_norm_ + Adagrad on G (train.py:317-318), then the loss scalars
    if ((rc = clip_opt_model(c, pg, L.opt_partial, L.scal + S_GSUMSQ, c->lr_g, c->wd_g, st)))
      return rc;
    if (has_d) cnt.n[R_REAL] = cnt.n[R_FAKE] = bce_blocks(M);
    if (has_adv) cnt.n[R_ADV] = bce_blocks(M);
    cnt.n[R_MGE] = sse_blocks(M, nS);
    cnt.n[R_MSE] = sse_blocks(M, d_out);
    GANTTS_PDL_LAUNCH((finalize_losses_kernel), 1, RED_THREADS, 0, st, L.scal, losses_dev, L.red, cnt, has_adv ? c->adv_w : 0.f, c->mge_w,
                                                      c->mse_w, has_d ? 1 : 0);
    GANTTS_LAUNCH_CHECK("finalize_losses_kernel");
  }
  return GANTTS_OK;
}
