/*
 * gantts_b200.h -- C ABI of libgantts_b200.so: the B200 (sm_100a) GAN-step hot path of r9y9/gantts.
 *
 * Every entry point takes plain device/host pointers and sizes; no torch types.  All work is
 * enqueued on the caller's `stream` (a cudaStream_t passed as void*), nothing synchronises
 * internally unless stated.  Return value: 0 on success, otherwise a GANTTS_E_* code; the message
 * for the calling thread's last failure is available from gantts_last_error_string().  The library
 * owns no device memory: every buffer, including workspaces, is the caller's.
 *
 * Each declaration cites the reference interface (r9y9/gantts @ fb1e75f, file:line) it replaces.
 * The reference-side binding is shown in INTEGRATION.md (ctypes, since the reference is Python).
 */
#ifndef GANTTS_B200_H_
#define GANTTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GANTTS_OK 0
#define GANTTS_E_BADARG 1   /* shape / pointer / enum out of contract */
#define GANTTS_E_CUDA 2     /* a CUDA runtime or driver call failed */
#define GANTTS_E_UNSUPPORTED 3
#define GANTTS_E_WORKSPACE 4 /* caller workspace too small */

#define GANTTS_MAX_STREAMS 8
#define GANTTS_MAX_WINDOWS 4
#define GANTTS_MAX_WINDOW_TAPS 5 /* l, u <= 2 */
#define GANTTS_MLPG_HALF_TAPS 24 /* FIR half width K: P^-1 decays to 2.6e-10 at lag 24 */
/* floats per row of the MLPG coefficient table: [0, 2K+1) rows of P^-1 (FIR form), [52, 56) = {1/L_tt, L[t][t-1], L[t][t-2], 0}
 * and [56, 60) = {L[t+1][t], L[t+2][t], 0, 0}: rows of the banded Cholesky factor of P for the substitution kernels */
#define GANTTS_MLPG_TABLE_COLS 60
#define GANTTS_MAX_LAYERS 8
#define GANTTS_MAX_COLS 256 /* static / adversarial column lists of the fused step */

int gantts_version(void);                       /* 100 * major + minor */
const char* gantts_last_error_string(void);     /* thread-local, never NULL */
/* 1 when the current device is compute capability 10.x (the kernels are sm_100a only). */
int gantts_device_supported(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long gantts_launch_count(void);
/* Measurement hooks: when enabled, the tensor-core GEMM and MLPG launches are bracketed by CUDA
 * events on the launching stream.  gantts_profile_collect synchronises those events and ADDS, per
 * kind (0 = GEMM K-major, 1 = GEMM MN-major, 2 = MLPG fwd, 3 = MLPG bwd), the elapsed milliseconds,
 * the algorithmic work (flops for GEMMs, bytes for MLPG) and the launch count into arrays of 8. */
int gantts_profile_enable(int on);
int gantts_profile_collect(double* ms, double* work, long long* launches);

/* ---------------------------------------------------------------------------------------------
 * Feature-stream layout (reference hparams.py:196-206: stream_sizes / has_dynamic_features).
 * Stream s occupies input columns [in_start, in_start + width) laid out window-major
 * [static sd | delta sd | delta-delta sd] when dyn != 0 (width = num_windows * sd), or sd plain
 * columns when dyn == 0; its static part lands in output columns [out_start, out_start + sd).
 * Disabled streams are simply left out of the table.
 */
typedef struct {
  int n;
  int in_start[GANTTS_MAX_STREAMS];
  int sd[GANTTS_MAX_STREAMS];
  int dyn[GANTTS_MAX_STREAMS];
  int out_start[GANTTS_MAX_STREAMS];
} gantts_streams_t;

/* Delta windows (reference hparams.py:22-26,183-187): window w has taps coef[w][0..l+u]. */
typedef struct {
  int n;
  int l[GANTTS_MAX_WINDOWS];
  int u[GANTTS_MAX_WINDOWS];
  float coef[GANTTS_MAX_WINDOWS][GANTTS_MAX_WINDOW_TAPS];
} gantts_windows_t;

/* ---------------------------------------------------------------------------------------------
 * MLPG (replaces nnmnkwii.paramgen.unit_variance_mlpg_matrix + nnmnkwii.autograd.unit_variance_mlpg
 * as called at reference train.py:510-513, gantts/multistream.py:82-123, gantts/models.py:66,115).
 *
 * The reference multiplies by the dense (T x nw*T) matrix R = (W^T W)^-1 W^T.  Here
 *   y = P^-1 (sum_w W_w^T mu_w),  P = sum_w W_w^T W_w (banded SPD),
 * is evaluated as a 3-tap stencil followed by a (2K+1)-tap row-variant FIR with the rows of P^-1
 * (K = GANTTS_MLPG_HALF_TAPS), over the PADDED length T for every batch row, exactly like the
 * reference (SURVEY.md 8a note iv).
 *
 * gantts_mlpg_table: HOST function.  Fills table_host[T * GANTTS_MLPG_TABLE_COLS] (row stride GANTTS_MLPG_TABLE_COLS) with
 *   table[t][j] = (P^-1)[t, t + j - K]   (0 outside [0,T)), computed in float64 by banded Cholesky,
 * stored as float32.  The caller uploads it once per (windows, T) and passes the device copy below.
 * Returns GANTTS_E_UNSUPPORTED when P^-1 has not decayed below 1e-8 of its diagonal at lag K.
 */
int gantts_mlpg_table(const gantts_windows_t* windows, int T, float* table_host);

/* out[b,t,out_col] = MLPG(in[b,:,stream cols]) for dynamic streams, copy for static ones.
 * in:  float32 [B][T][*] with element strides (in_bstride, in_tstride), unit column stride.
 * out: float32 [B][T][*] with element strides (out_bstride, out_tstride). */
int gantts_mlpg_fwd(const float* in, int64_t in_bstride, int64_t in_tstride,
                    float* out, int64_t out_bstride, int64_t out_tstride,
                    const float* table_dev, const gantts_streams_t* streams,
                    const gantts_windows_t* windows, int B, int T, void* stream);

/* grad_in[b,t,stream cols] (+)= W_w P^-1 grad_out (backward of the above; reference backward is
 * R^T g).  Columns of grad_in belonging to no listed stream are NOT written.  accumulate != 0
 * adds into grad_in instead of overwriting. */
int gantts_mlpg_bwd(const float* grad_out, int64_t go_bstride, int64_t go_tstride,
                    float* grad_in, int64_t gi_bstride, int64_t gi_tstride,
                    const float* table_dev, const gantts_streams_t* streams,
                    const gantts_windows_t* windows, int B, int T, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stream column gathers (reference gantts/multistream.py:33-43 select_streams, :56-79
 * get_static_features, train.py:232-242 get_selected_static_stream).  Pure copies => bit-exact.
 * out[r, j] = in[r, cols[j]] for r < rows.  cols_dev: int32[ncols] on the device.
 */
int gantts_gather_cols(const float* in, int64_t in_rstride, float* out, int64_t out_rstride,
                       const int32_t* cols_dev, int ncols, int64_t rows, void* stream);
/* Scatter-add of the backward: gin[r, cols[j]] += gout[r, j]. */
int gantts_scatter_cols_add(const float* gout, int64_t go_rstride, float* gin, int64_t gi_rstride,
                            const int32_t* cols_dev, int ncols, int64_t rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sequence mask + masked MSE (reference gantts/seqloss.py:9-20 sequence_mask, :27-43
 * MaskedMSELoss.forward).
 */
/* mask[b,t] = (t < lengths[b]) ? 1.f : 0.f ; lengths_dev int64[B]. */
int gantts_sequence_mask(const int64_t* lengths_dev, float* mask, int B, int T, void* stream);

/* sums_dev[0] = sum_{b,t,d} ((a - b) * m[b,t])^2 ; sums_dev[1] = sum_{b,t} m[b,t].
 * a, b: float32 [rows][D] with row strides; mask: float32[rows].  Deterministic two-pass reduction;
 * workspace must hold gantts_masked_sse_workspace_bytes() bytes. The loss is sums[0] / sums[1]. */
size_t gantts_masked_sse_workspace_bytes(void);
int gantts_masked_sse_fwd(const float* a, int64_t a_rstride, const float* b, int64_t b_rstride,
                          const float* mask, int64_t rows, int D, float* sums_dev,
                          void* workspace, size_t workspace_bytes, void* stream);
/* grad_a[r,d] (+)= scale_dev[0] * 2 * (a - b) * m^2 ; scale is read on the device
 * (= upstream_grad / sum(mask)), so no host sync is needed. */
int gantts_masked_sse_bwd(const float* a, int64_t a_rstride, const float* b, int64_t b_rstride,
                          const float* mask, int64_t rows, int D, const float* scale_dev,
                          float* grad_a, int64_t ga_rstride, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Masked adversarial BCE terms (written inline in reference train.py:258-271,286,307-310):
 *   kind 0 ("real"/"adv"):  -(log(D + 1e-20) * m).sum()      count: sum((D > 0.5) * m)
 *   kind 1 ("fake"):        -(log(1 - D + 1e-20) * m).sum()   count: sum((D < 0.5) * m)
 * D: float32[rows] discriminator outputs (after sigmoid).  out_dev[0] = un-normalised loss sum,
 * out_dev[1] = count, out_dev[2] = sum(mask).  Same workspace contract as masked_sse.
 */
int gantts_masked_bce_fwd(const float* D, const float* mask, int64_t rows, int kind,
                          float* out_dev, void* workspace, size_t workspace_bytes, void* stream);
/* grad_D[r] = scale_dev[0] * d/dD of the un-normalised sum above. */
int gantts_masked_bce_bwd(const float* D, const float* mask, int64_t rows, int kind,
                          const float* scale_dev, float* grad_D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Linear -> LeakyReLU(0.01) -> Dropout layer (reference gantts/models.py:137-141 MLP.forward,
 * :63-65 In2OutHighwayNet.forward; NOTE the reference order is Linear -> LeakyReLU -> Dropout).
 *
 * act: 0 = none (last_linear), 1 = LeakyReLU(slope) then dropout(p), 2 = sigmoid.
 * Dropout keeps an element with probability 1-p and scales it by 1/(1-p); the keep decision is a
 * counter-based hash of (seed, row * N + col), so no mask is stored: the backward recovers
 * "dropped" / "negative" from the saved OUTPUT y (y == 0 <=> dropped, sign(y) = sign(pre-act)).
 *
 * engine: GANTTS_ENGINE_SIMT  = exact fp32 FFMA tiles (validation / odd shapes),
 *         GANTTS_ENGINE_TC    = tcgen05 tensor cores, bf16x3 split operands with fp32 accumulation
 *                               in TMEM (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, ~2^-16 per product).
 */
#define GANTTS_ENGINE_SIMT 0
#define GANTTS_ENGINE_TC 1
#define GANTTS_ACT_NONE 0
#define GANTTS_ACT_LEAKY_DROPOUT 1
#define GANTTS_ACT_SIGMOID 2

/* y[M][N] = act(x[M][K] W[N][K]^T + bias[N]); x, y row-major with row strides; W row-major [N][K]
 * (the nn.Linear state_dict layout). */
int gantts_linear_fwd(const float* x, int64_t x_rstride, const float* W, const float* bias,
                      float* y, int64_t y_rstride, int64_t M, int N, int K, int act, float slope,
                      float p, uint64_t seed, int engine, void* workspace, size_t workspace_bytes,
                      void* stream);
size_t gantts_linear_workspace_bytes(int64_t M, int N, int K, int engine);

/* Backward of the fused layer.  gy: upstream gradient w.r.t. the layer OUTPUT y [M][N].
 * Computes gz = gy * act'(y) (in place into gz_scratch [M][N], may alias gy when gy is dead),
 *   gx[M][K]  (=) gz W          (skipped when gx == NULL),
 *   gW[N][K] (+)= gz^T x,  gb[N] (+)= column sums of gz   (accumulate != 0 adds; skipped if NULL).
 */
int gantts_linear_bwd(const float* gy, int64_t gy_rstride, const float* y, int64_t y_rstride,
                      const float* x, int64_t x_rstride, const float* W,
                      float* gz_scratch, float* gx, int64_t gx_rstride, float* gW, float* gb,
                      int64_t M, int N, int K, int act, float slope, float p, int accumulate,
                      int engine, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole MLP on the tensor-core engine (reference gantts/models.py:121-141 MLP, used as generator and
 * discriminator): hidden layers Linear -> LeakyReLU(slope) -> Dropout(p), then last_linear with
 * last_act (NONE or SIGMOID).  Activations stay resident as bf16 hi/lo planes between layers; the
 * caller-owned `tape` keeps them (plus the split weights) for the backward.  dropout_p = 0 in eval
 * mode; hidden layer l draws its mask from seed + golden-ratio * (l+1).
 */
typedef struct {
  int num_layers;                     /* linear layers including last_linear, 1..GANTTS_MAX_LAYERS */
  int dims[GANTTS_MAX_LAYERS + 1];    /* dims[0] = in ... dims[num_layers] = out */
  const float* W[GANTTS_MAX_LAYERS];  /* W[l]: [dims[l+1]][dims[l]] row-major (nn.Linear layout) */
  const float* b[GANTTS_MAX_LAYERS];  /* b[l]: [dims[l+1]], 16-byte aligned */
  float slope;
  float dropout_p;
  int last_act;
  uint64_t seed;
} gantts_mlp_t;

size_t gantts_mlp_tape_bytes(const gantts_mlp_t* mlp, int64_t M);
size_t gantts_mlp_workspace_bytes(const gantts_mlp_t* mlp, int64_t M);
/* y[M][dims[L]] = MLP(x[M][dims[0]]); fills `tape`. */
int gantts_mlp_fwd(const gantts_mlp_t* mlp, const float* x, int64_t x_rstride, int64_t M, float* y,
                   int64_t y_rstride, void* tape, size_t tape_bytes, void* stream);
/* Backward from gy = dL/dy.  y is the forward output (needed for SIGMOID, may be NULL otherwise).
 * gW[l] / gb[l] (host arrays of device pointers, entries may be NULL) receive (+= when accumulate)
 * the parameter gradients; gx (may be NULL) receives dL/dx. */
int gantts_mlp_bwd(const gantts_mlp_t* mlp, const float* gy, int64_t gy_rstride, const float* y,
                   int64_t y_rstride, int64_t M, const void* tape, size_t tape_bytes, float* gx,
                   int64_t gx_rstride, float* const* gW, float* const* gb, int accumulate,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gradient clipping + Adagrad (torch.nn.utils.clip_grad_norm_ + torch.optim.Adagrad as used at
 * reference train.py:275-276,317-318 with hparams.py:223-227,240-244).  Operates on a list of
 * parameter tensors given as device pointer arrays.
 */
/* Tensor lists are HOST arrays of device pointers (any count: lists longer than 32 tensors are processed in
 * chunks of 32 that share the same sum of squares); sizes are element counts.
 * sumsq_dev[0] = sum over all tensors of g^2 (deterministic two-stage reduction). */
size_t gantts_optim_workspace_bytes(void);
int gantts_grad_sumsq(float* const* grads, const int64_t* sizes_host, int ntensors, float* sumsq_dev,
                      void* workspace, size_t workspace_bytes, void* stream);
/* coef = min(1, max_norm / (sqrt(sumsq) + 1e-6)); g *= coef (in place, like clip_grad_norm_);
 * g' = g + wd * p; s += g'*g'; p -= lr * g' / (sqrt(s) + eps).  sumsq_dev is read on the device. */
int gantts_clip_adagrad_step(float* const* params, float* const* grads, float* const* state_sums,
                             const int64_t* sizes_host, int ntensors, const float* sumsq_dev,
                             float max_norm, float lr, float weight_decay, float eps, void* stream);
/* clip_grad_norm_ + torch.optim.Adam.step() (reference hparams.py:125-130: the duration model's optimiser,
 * lr 1e-3, betas (0.5, 0.9), weight_decay 0, eps 1e-8, amsgrad off):  g *= coef; g' = g + wd * p;
 * m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),
 * t = step (1 for the first call). */
int gantts_clip_adam_step(float* const* params, float* const* grads, float* const* exp_avg,
                          float* const* exp_avg_sq, const int64_t* sizes_host, int ntensors,
                          const float* sumsq_dev, float max_norm, float lr, float beta1, float beta2,
                          float weight_decay, float eps, int64_t step, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LSTM layer with packed-sequence semantics (reference gantts/models.py:84-85 In2OutRNNHighwayNet,
 * :175-176 GRURNN -- an nn.LSTM --, :198-199 LSTMRNN; pack_padded_sequence / pad_packed_sequence at
 * :101-112,182-187,205-210).  Gate order i,f,g,o and weight layout of torch.nn.LSTM.
 *
 * xproj  [B][T][ndir*4H] = x W_ih^T + b_ih + b_hh for every time step (one tensor-core GEMM, e.g.
 *        gantts_linear_fwd with the direction-stacked W_ih), direction-major columns.
 * W_hh   [ndir][4H][H].   lengths_dev int64[B] (any order).   ndir = 1 or 2 (bidirectional).
 * h_out  [B][T][ndir*H]: hidden states, ZERO for t >= lengths[b]; the reverse direction starts at
 *        t = lengths[b]-1.   gates [ndir][B][T][4H] / cells [ndir][B][T][H]: saved for the backward.
 * One cooperative launch runs all T steps of both directions (persistent CTAs, W_hh slices resident in
 * shared memory, one grid barrier per step).  workspace: gantts_lstm_workspace_bytes() bytes.
 */
size_t gantts_lstm_workspace_bytes(void);
int gantts_lstm_layer_fwd(const float* xproj, const float* W_hh, const int64_t* lengths_dev, float* h_out,
                          float* gates, float* cells, int B, int T, int H, int ndir, void* workspace,
                          size_t workspace_bytes, void* stream);
/* dxproj [B][T][ndir*4H] = dL/d(xproj) by back-propagation through time from dh_out = dL/dh_out. */
int gantts_lstm_layer_bwd(const float* dh_out, const float* W_hh, const int64_t* lengths_dev,
                          const float* gates, const float* cells, float* dxproj, int B, int T, int H,
                          int ndir, void* workspace, size_t workspace_bytes, void* stream);
/* hprev[b][t][:] = h_out[b][t-1 (dir 0) | t+1 (dir 1)][dir*H:(dir+1)*H], zero at the first step of each
 * sequence and beyond its length: the right operand of dW_hh[dir] = dxproj_dir^T hprev. */
int gantts_lstm_hprev(const float* h_out, const int64_t* lengths_dev, float* hprev, int B, int T, int H,
                      int ndir, int dir, void* stream);
/* y = keep ? x/(1-p) : 0 with the counter-hash mask (inter-layer dropout of nn.LSTM(dropout=p)); applying
 * the same call to the gradient is the backward. */
int gantts_dropout(const float* x, float* y, int64_t rows, int cols, float p, uint64_t seed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SRU v1 scan (third-party cuda_functional.SRU imported by reference gantts/models.py:144-167 SRURNN;
 * github.com/taolei87/sru is not vendored and untested by the reference: parity unpinned).
 * u [B][T][ncols*k] = x W (tensor-core GEMM by the caller), ncols = d * (bidir ? 2 : 1), k = 3 or 4 values per
 * column (k fastest): candidate, forget pre-activation, reset pre-activation[, highway input];
 * x [B][T][ncols] is the highway input when k == 3; bias [2*ncols] = forget | reset; mask_h [B][ncols]
 * optional (already scaled) output dropout mask shared over time; act: 0 identity, 1 tanh, 2 relu.
 * Outputs h, c [B][T][ncols].  Backward: du [B][T][ncols*k], dx += (k == 3), dbias_part [B][2*ncols]
 * (sum over B gives the bias gradient).
 */
int gantts_sru_fwd(const float* u, const float* x, const float* bias, const float* mask_h, float* h, float* c,
                   int B, int T, int d, int k, int bidir, int act, void* stream);
int gantts_sru_bwd(const float* u, const float* x, const float* bias, const float* mask_h, const float* c,
                   const float* dh, float* du, float* dx, float* dbias_part, int B, int T, int d, int k,
                   int bidir, int act, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused GAN training step: one call enqueues the whole mini-batch of reference train.py:528-580
 * (batch prologue :528-535, apply_generator :336-355, update_discriminator :245-279,
 * update_generator :282-320, both clip_grad_norm_ + Adagrad steps) on `stream`, no host sync.
 *
 * MLP generator (g, linear output) and MLP discriminator (d, one sigmoid output, input = the
 * `adv_cols` columns of the static features).  Weights (g.W/g.b, d.W/d.b) and the Adagrad
 * accumulators are updated IN PLACE.  phases is a bit mask so that a data-parallel caller can
 * all-reduce the gradient buffers (gantts_gan_step_grad_buffer) between the pieces:
 *   1 = prologue, G forward, MLPG, D forward on [real | fake], loss_d backward  -> D gradients ready
 *   2 = D clip+Adagrad, MGE/MSE/ADV losses, third D forward, loss_g backward    -> G gradients ready
 *   4 = G clip+Adagrad, loss scalars
 * inv_frames = 1 / (GLOBAL number of valid frames) (reference normaliser T = mask.sum()); a value <= 0 makes the
 * step derive it on the device from lengths_dev (single-process use).
 * losses_dev[12] = loss_d, loss_fake_d, loss_real_d, loss_mse, loss_mge, loss_adv, loss_g,
 *                  real_correct, fake_correct, local frames, d_grad_norm, g_grad_norm.
 */
typedef struct {
  int B, T;
  gantts_mlp_t g;                          /* generator: dims[0] = linguistic width, dims[L] = acoustic width */
  gantts_mlp_t d;                          /* discriminator: dims[0] = n_adv, dims[L] = 1, last_act = SIGMOID */
  float* g_sumW[GANTTS_MAX_LAYERS];        /* Adagrad state_sum per parameter tensor */
  float* g_sumb[GANTTS_MAX_LAYERS];
  float* d_sumW[GANTTS_MAX_LAYERS];
  float* d_sumb[GANTTS_MAX_LAYERS];
  gantts_streams_t streams;                /* MLPG stream layout of the generator output */
  gantts_windows_t windows;
  const float* mlpg_table;                 /* device copy of gantts_mlpg_table(windows, T) */
  int n_static;                            /* width of y_hat_static */
  int n_static_cols;                       /* == n_static */
  int static_cols[GANTTS_MAX_COLS];        /* columns of y forming y_static (get_static_features) */
  int n_adv;
  int adv_cols[GANTTS_MAX_COLS];           /* columns of y_(hat_)static fed to D (select + mask_nth) */
  int d_conditioned;                       /* hp.discriminator_linguistic_condition (train.py:254-256): D input =
                                              cat((x, y_adv), -1), d.dims[0] = g.dims[0] + n_adv */
  float lr_g, lr_d, wd_g, wd_d, eps, max_norm;
  float w_d, mse_w, mge_w, adv_w;
  /* Optimiser of both models (reference train.py:784-789 getattr(optim, hp.optimizer_g)(...)): 0 = Adagrad
   * (hparams.py:201-206; *_sum* = state_sum), 1 = Adam (hparams.py:125-130, the duration model: lr 1e-3,
   * betas (0.5, 0.9), weight_decay 0, eps 1e-8, amsgrad off; *_sum* = exp_avg, *_sq* = exp_avg_sq, opt_step =
   * number of the step being taken, 1 for the first -- the bias corrections are computed on the host). */
  int optimizer;
  float beta1, beta2;
  int64_t opt_step;
  float* g_sqW[GANTTS_MAX_LAYERS];
  float* g_sqb[GANTTS_MAX_LAYERS];
  float* d_sqW[GANTTS_MAX_LAYERS];
  float* d_sqb[GANTTS_MAX_LAYERS];
} gantts_gan_step_t;
#define GANTTS_OPT_ADAGRAD 0
#define GANTTS_OPT_ADAM 1

/* Phase bits of gantts_gan_step.  GANTTS_STEP_EVAL = the "test" phase of reference train.py:481-486,
 * :273,:315 (model.eval(), phase != "train"): forwards and losses only -- dropout off, no backward, no
 * optimiser step, parameters and Adagrad state untouched; the adversarial loss re-uses the fake half of the
 * discriminator forward (with dropout off and no discriminator step in between, the reference's third forward
 * returns exactly those values). */
#define GANTTS_STEP_D 1
#define GANTTS_STEP_G 2
#define GANTTS_STEP_FINISH 4
#define GANTTS_STEP_EVAL 8

/* Dropout seeds of the fused step, so a test can regenerate every keep mask with gantts_dropout():
 * forward `which` (0 generator, 1 stacked real|fake discriminator batch, 2 adversarial discriminator forward)
 * of a step called with `seed` runs its MLP with gantts_gan_step_seed(seed, which); hidden layer l of an MLP
 * run with seed s draws its mask as gantts_dropout(ones[rows][dims[l+1]], p, gantts_mlp_layer_seed(s, l)). */
uint64_t gantts_gan_step_seed(uint64_t seed, int which);
uint64_t gantts_mlp_layer_seed(uint64_t seed, int layer);

size_t gantts_gan_step_workspace_bytes(const gantts_gan_step_t* cfg);
/* Flat gradient buffer inside `workspace` (which: 0 = generator, 1 = discriminator). */
int gantts_gan_step_grad_buffer(const gantts_gan_step_t* cfg, void* workspace, int which, float** ptr,
                                int64_t* count);
int gantts_gan_step(const gantts_gan_step_t* cfg, int phases, const float* x, const float* y,
                    const int64_t* lengths_dev, float inv_frames, uint64_t seed, float* y_hat,
                    float* y_hat_static, float* losses_dev, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Inference-time MLPG with real variances (replaces nnmnkwii.paramgen.mlpg as called by reference
 * evaluation_tts.py:70-72,92-94):  solves, per static dimension d and batch row b,
 *   (sum_w W_w^T diag(1/var_w) W_w) y = sum_w W_w^T diag(1/var_w) mean_w
 * by banded Cholesky in float64.  mean/var: float32 [B][T][nw*sd], window-major feature blocks
 * ([static sd, delta sd, delta-delta sd]), element strides (bstride, tstride); a time-invariant variance
 * vector (the reference's use) is passed with v_tstride = 0 (and v_bstride = 0).  out: float32 [B][T][sd].
 */
size_t gantts_mlpg_var_workspace_bytes(const gantts_windows_t* windows, int B, int T, int sd);
int gantts_mlpg_var(const float* mean, int64_t m_bstride, int64_t m_tstride, const float* var,
                    int64_t v_bstride, int64_t v_tstride, float* out, int64_t o_bstride, int64_t o_tstride,
                    const gantts_windows_t* windows, int B, int T, int sd, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Objective distortions of the training loop (reference train.py:399-432 compute_distortions, :383-396
 * split_streams, :358-380 inv_scale; nnmnkwii.metrics.{melcd, lf0_mean_squared_error, vuv_error,
 * mean_squared_error}).  y, y_hat: float32 [B][T][D] static-domain features (normalised); mean_dev /
 * std_dev: float32[D] de-normalisation per static column (the caller maps the reference's
 * static+dynamic-domain indices).  Column groups (count 0 / col -1 disables a term):
 *   mcd  : [mcd_start, +mcd_count)   sum over valid frames of ||delta||_2   (reference passes mgc[:, :, 1:])
 *   bap  : [bap_start, +bap_count)   the same for band aperiodicity
 *   lf0_col, vuv_col: F0 squared error (after exp when lf0_linear) over frames voiced in both; V/UV is
 *          binarised at 0.5 after de-normalisation (train.py:374-377)
 *   mse  : [mse_start, +mse_count)   plain squared error sum (duration model / VC)
 * out8_dev: { sum ||d mcd||, sum ||d bap||, sum f0 err^2, #frames voiced in both, #V/UV mismatches,
 *             #valid frames, sum sq err of the mse group, 0 }.  One device-to-host read of 32 bytes
 * gives every metric: mcd = 10/ln10*sqrt(2) * out[0]/out[5], f0_rmse = sqrt(out[2]/out[3]), ...
 */
typedef struct {
  int mcd_start, mcd_count;
  int bap_start, bap_count;
  int lf0_col, vuv_col;
  int lf0_linear;
  int mse_start, mse_count;
} gantts_distortion_cols_t;

size_t gantts_distortions_workspace_bytes(void);
int gantts_distortions(const float* y, int64_t y_bstride, int64_t y_tstride, const float* y_hat,
                       int64_t yh_bstride, int64_t yh_tstride, const int64_t* lengths_dev, int B, int T, int D,
                       const float* mean_dev, const float* std_dev, const gantts_distortion_cols_t* cols,
                       float* out8_dev, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GANTTS_B200_H_ */
