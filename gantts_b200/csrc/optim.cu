// clip_grad_norm_ + Adagrad over a short list of parameter tensors (include/gantts_b200.h).
// Reference: torch.nn.utils.clip_grad_norm_(params, 1.0) then torch.optim.Adagrad.step(), as called at
// train.py:275-276,317-318 with lr 0.01, weight_decay 1e-7 (hparams.py:223-227,240-244).
#include "common.cuh"

namespace gantts {

constexpr int OPT_MAX_TENSORS = 32;
constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_BLOCKS = 148 * 4;

struct TensorList {
  int n;
  float* p[OPT_MAX_TENSORS];
  float* g[OPT_MAX_TENSORS];
  float* s[OPT_MAX_TENSORS];
  int64_t off[OPT_MAX_TENSORS + 1];   // prefix sums of sizes
};

__device__ __forceinline__ int find_tensor(const TensorList& tl, int64_t i) {
  int k = 0;
  while (k + 1 < tl.n && i >= tl.off[k + 1]) ++k;
  return k;
}

__global__ void __launch_bounds__(OPT_THREADS)
sumsq_partial_kernel(TensorList tl, float* partial) {
  __shared__ float sm[32];
  float v[1] = {0.f};
  const int64_t total = tl.off[tl.n];
  for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * OPT_THREADS) {
    int k = find_tensor(tl, i);
    float g = tl.g[k][i - tl.off[k]];
    v[0] = fmaf(g, g, v[0]);
  }
  block_sum<1>(v, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = v[0];
}

__global__ void __launch_bounds__(OPT_THREADS)
sumsq_finish_kernel(const float* partial, int n, float* out) {
  __shared__ float sm[32];
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < n; i += OPT_THREADS) v[0] += partial[i];
  block_sum<1>(v, sm);
  if (threadIdx.x == 0) out[0] = v[0];
}

__global__ void __launch_bounds__(OPT_THREADS)
clip_adagrad_kernel(TensorList tl, const float* __restrict__ sumsq, float max_norm, float lr, float wd,
                    float eps) {
  const float total_norm = sqrtf(sumsq[0]);
  float coef = max_norm / (total_norm + 1e-6f);
  coef = coef < 1.f ? coef : 1.f;
  const int64_t total = tl.off[tl.n];
  for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * OPT_THREADS) {
    int k = find_tensor(tl, i);
    int64_t j = i - tl.off[k];
    float g = tl.g[k][j] * coef;
    tl.g[k][j] = g;                       // clip_grad_norm_ scales .grad in place
    float p = tl.p[k][j];
    g = fmaf(wd, p, g);
    float s = fmaf(g, g, tl.s[k][j]);
    tl.s[k][j] = s;
    tl.p[k][j] = p - lr * g / (sqrtf(s) + eps);
  }
}

static int fill(TensorList& tl, float* const* params, float* const* grads, float* const* sums,
                const int64_t* sizes, int n) {
  GANTTS_CHECK_ARG(n >= 1 && n <= OPT_MAX_TENSORS, "optim: tensor count %d out of [1,%d]", n, OPT_MAX_TENSORS);
  tl.n = n;
  tl.off[0] = 0;
  for (int i = 0; i < n; ++i) {
    GANTTS_CHECK_ARG(sizes[i] >= 1 && grads[i], "optim: bad tensor %d", i);
    tl.p[i] = params ? params[i] : nullptr;
    tl.g[i] = grads[i];
    tl.s[i] = sums ? sums[i] : nullptr;
    tl.off[i + 1] = tl.off[i] + sizes[i];
  }
  return GANTTS_OK;
}

static int blocks_for(int64_t total) {
  int64_t b = (total + OPT_THREADS * 4 - 1) / (OPT_THREADS * 4);
  if (b < 1) b = 1;
  if (b > OPT_MAX_BLOCKS) b = OPT_MAX_BLOCKS;
  return (int)b;
}

}  // namespace gantts

using namespace gantts;

extern "C" size_t gantts_optim_workspace_bytes(void) { return OPT_MAX_BLOCKS * sizeof(float); }

extern "C" int gantts_grad_sumsq(float* const* grads, const int64_t* sizes_host, int ntensors,
                                 float* sumsq_dev, void* workspace, size_t workspace_bytes, void* stream) {
  TensorList tl;
  int rc = fill(tl, nullptr, grads, nullptr, sizes_host, ntensors);
  if (rc) return rc;
  if (!workspace || workspace_bytes < OPT_MAX_BLOCKS * sizeof(float)) {
    set_error("grad_sumsq: workspace too small");
    return GANTTS_E_WORKSPACE;
  }
  int nb = blocks_for(tl.off[tl.n]);
  float* partial = static_cast<float*>(workspace);
  sumsq_partial_kernel<<<nb, OPT_THREADS, 0, as_stream(stream)>>>(tl, partial);
  GANTTS_LAUNCH_CHECK("sumsq_partial_kernel");
  sumsq_finish_kernel<<<1, OPT_THREADS, 0, as_stream(stream)>>>(partial, nb, sumsq_dev);
  GANTTS_LAUNCH_CHECK("sumsq_finish_kernel");
  return GANTTS_OK;
}

extern "C" int gantts_clip_adagrad_step(float* const* params, float* const* grads, float* const* state_sums,
                                        const int64_t* sizes_host, int ntensors, const float* sumsq_dev,
                                        float max_norm, float lr, float weight_decay, float eps,
                                        void* stream) {
  GANTTS_CHECK_ARG(params && state_sums && sumsq_dev, "clip_adagrad_step: null pointer");
  TensorList tl;
  int rc = fill(tl, params, grads, state_sums, sizes_host, ntensors);
  if (rc) return rc;
  for (int i = 0; i < ntensors; ++i) GANTTS_CHECK_ARG(tl.p[i] && tl.s[i], "clip_adagrad_step: null tensor %d", i);
  int nb = blocks_for(tl.off[tl.n]);
  clip_adagrad_kernel<<<nb, OPT_THREADS, 0, as_stream(stream)>>>(tl, sumsq_dev, max_norm, lr, weight_decay, eps);
  GANTTS_LAUNCH_CHECK("clip_adagrad_kernel");
  return GANTTS_OK;
}
