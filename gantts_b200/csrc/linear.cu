// C-ABI dispatch for the fused Linear -> LeakyReLU -> Dropout layer (include/gantts_b200.h).
#include "common.cuh"

namespace gantts {
size_t simt_workspace_bytes(int64_t M, int N, int K);
int simt_linear_fwd(const float* x, int64_t x_rs, const float* W, const float* bias, float* y,
                    int64_t y_rs, int64_t M, int N, int K, int act, float slope, float p,
                    uint64_t seed, cudaStream_t st);
int act_bwd(const float* gy, int64_t gy_rs, const float* y, int64_t y_rs, float* gz, int64_t M, int N,
            int act, float slope, float p, cudaStream_t st);
int colsum(const float* gz, int64_t M, int N, float* gb, int accumulate, float* partial, cudaStream_t st);
int simt_linear_bwd_gemms(const float* gz, const float* x, int64_t x_rs, const float* W, float* gx,
                          int64_t gx_rs, float* gW, int64_t M, int N, int K, int accumulate,
                          float* ws, cudaStream_t st);
size_t tc_linear_workspace_bytes(int64_t M, int N, int K);
int tc_linear_fwd(const float* x, int64_t x_rs, const float* W, const float* bias, float* y,
                  int64_t y_rs, int64_t M, int N, int K, int act, float slope, float p, uint64_t seed,
                  void* ws, size_t ws_bytes, cudaStream_t st);
int tc_linear_bwd_gemms(const float* gz, const float* x, int64_t x_rs, const float* W, float* gx,
                        int64_t gx_rs, float* gW, int64_t M, int N, int K, int accumulate, void* ws,
                        size_t ws_bytes, cudaStream_t st);
}  // namespace gantts

using namespace gantts;

static int check_common(int64_t M, int N, int K, int act, float p, int engine) {
  GANTTS_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, "linear: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  GANTTS_CHECK_ARG(act >= 0 && act <= 2, "linear: bad act %d", act);
  GANTTS_CHECK_ARG(p >= 0.f && p < 1.f, "linear: dropout p=%f out of [0,1)", p);
  GANTTS_CHECK_ARG(engine == GANTTS_ENGINE_SIMT || engine == GANTTS_ENGINE_TC, "linear: bad engine %d", engine);
  return GANTTS_OK;
}

// Workspace layout: [column-sum partials: 64*N floats, 256B aligned][engine area].
static size_t colsum_area_bytes(int N) { return (((size_t)64 * N * sizeof(float)) + 255) / 256 * 256; }

extern "C" size_t gantts_linear_workspace_bytes(int64_t M, int N, int K, int engine) {
  size_t a = simt_workspace_bytes(M, N, K);
  if (engine == GANTTS_ENGINE_TC) {
    size_t b = tc_linear_workspace_bytes(M, N, K);
    a = a > b ? a : b;
  }
  return colsum_area_bytes(N) + a;
}

extern "C" int gantts_linear_fwd(const float* x, int64_t x_rs, const float* W, const float* bias,
                                 float* y, int64_t y_rs, int64_t M, int N, int K, int act, float slope,
                                 float p, uint64_t seed, int engine, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  int rc = check_common(M, N, K, act, p, engine);
  if (rc) return rc;
  GANTTS_CHECK_ARG(x && W && y && x_rs >= K && y_rs >= N, "linear_fwd: bad pointers/strides");
  if (engine == GANTTS_ENGINE_TC)
    return tc_linear_fwd(x, x_rs, W, bias, y, y_rs, M, N, K, act, slope, p, seed, workspace,
                         workspace_bytes, as_stream(stream));
  return simt_linear_fwd(x, x_rs, W, bias, y, y_rs, M, N, K, act, slope, p, seed, as_stream(stream));
}

extern "C" int gantts_linear_bwd(const float* gy, int64_t gy_rs, const float* y, int64_t y_rs,
                                 const float* x, int64_t x_rs, const float* W, float* gz_scratch,
                                 float* gx, int64_t gx_rs, float* gW, float* gb, int64_t M, int N, int K,
                                 int act, float slope, float p, int accumulate, int engine,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(M, N, K, act, p, engine);
  if (rc) return rc;
  GANTTS_CHECK_ARG(gy && y && W && gz_scratch, "linear_bwd: null pointer");
  GANTTS_CHECK_ARG(!gW || x, "linear_bwd: gW requested without x");
  size_t need = gantts_linear_workspace_bytes(M, N, K, engine);
  if (!workspace || workspace_bytes < need) {
    set_error("linear_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GANTTS_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  rc = act_bwd(gy, gy_rs, y, y_rs, gz_scratch, M, N, act, slope, p, st);
  if (rc) return rc;
  if (gb) {
    rc = colsum(gz_scratch, M, N, gb, accumulate, static_cast<float*>(workspace), st);
    if (rc) return rc;
  }
  char* area = static_cast<char*>(workspace) + colsum_area_bytes(N);
  size_t area_bytes = workspace_bytes - colsum_area_bytes(N);
  if (engine == GANTTS_ENGINE_TC)
    return tc_linear_bwd_gemms(gz_scratch, x, x_rs, W, gx, gx_rs, gW, M, N, K, accumulate, area,
                               area_bytes, st);
  return simt_linear_bwd_gemms(gz_scratch, x, x_rs, W, gx, gx_rs, gW, M, N, K, accumulate,
                               reinterpret_cast<float*>(area), st);
}
