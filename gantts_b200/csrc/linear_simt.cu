// Exact-fp32 (FFMA) engine for the fused Linear -> LeakyReLU -> Dropout layer: the on-device
// validation path for the tcgen05 engine and the fallback for shapes the tensor path does not take
// (see include/gantts_b200.h, GANTTS_ENGINE_SIMT).  Also hosts the engine-independent elementwise
// pieces of the layer backward (activation derivative, bias-gradient column sums).
#include "common.cuh"

namespace gantts {

constexpr int BM = 64, BN = 64, BK = 16, SIMT_THREADS = 256;

struct Epilogue {
  const float* bias;   // [N] or null
  int act;             // GANTTS_ACT_*
  float slope;
  float keep_scale;    // 1/(1-p)
  uint32_t thresh;     // round(p * 65536), 0 => no dropout
  uint64_t seed;
  int accumulate;      // C += result (no bias/act)
};

// C[i][j] = epi( sum_k A(i,k) * B(k,j) ), A(i,k) = A[i*a_si + k*a_sk], B(k,j) = B[k*b_sk + j*b_sj].
// Split along k over gridDim.z: slice z handles k in [z*kchunk, min(K,(z+1)*kchunk)) and writes to
// C + z*c_zstride (partials reduced by splitk_reduce_kernel).
__global__ void __launch_bounds__(SIMT_THREADS)
sgemm_kernel(const float* __restrict__ A, int64_t a_si, int64_t a_sk, const float* __restrict__ Bm,
             int64_t b_sk, int64_t b_sj, float* __restrict__ C, int64_t c_si, int64_t c_zstride,
             int64_t Mi, int Nj, int64_t Kk, int64_t kchunk, Epilogue ep) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * BM;
  const int j0 = blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = kbeg + kchunk < Kk ? kbeg + kchunk : Kk;
  const int ty = tid / 16, tx = tid % 16;          // 16x16 threads, 4x4 outputs each
  float acc[4][4] = {};
  // Loader mapping: make the unit-stride dimension the fastest-varying across threads.
  const bool a_kfast = (a_sk == 1), b_jfast = (b_sj == 1);
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int e = 0; e < (BM * BK) / SIMT_THREADS; ++e) {
      int idx = tid + e * SIMT_THREADS;
      int ii, kk;
      if (a_kfast) { kk = idx % BK; ii = idx / BK; } else { ii = idx % BM; kk = idx / BM; }
      int64_t gi = i0 + ii, gk = k0 + kk;
      As[kk][ii] = (gi < Mi && gk < kend) ? A[gi * a_si + gk * a_sk] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < (BN * BK) / SIMT_THREADS; ++e) {
      int idx = tid + e * SIMT_THREADS;
      int jj, kk;
      if (b_jfast) { jj = idx % BN; kk = idx / BN; } else { kk = idx % BK; jj = idx / BK; }
      int64_t gk = k0 + kk;
      int gj = j0 + jj;
      Bs[kk][jj] = (gj < Nj && gk < kend) ? Bm[gk * b_sk + (int64_t)gj * b_sj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = As[kk][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = Bs[kk][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
  }
  float* Cz = C + (int64_t)blockIdx.z * c_zstride;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int64_t gi = i0 + ty * 4 + r;
    if (gi >= Mi) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int gj = j0 + tx * 4 + c;
      if (gj >= Nj) continue;
      float v = acc[r][c];
      float* p = Cz + gi * c_si + gj;
      if (ep.accumulate) { *p += v; continue; }
      if (ep.bias) v += ep.bias[gj];
      if (ep.act == GANTTS_ACT_LEAKY_DROPOUT) {
        v = v > 0.f ? v : v * ep.slope;
        if (ep.thresh) {
          bool keep = dropout_keep(ep.seed, (uint32_t)gi, (uint32_t)Nj, (uint32_t)gj, ep.thresh);
          v = keep ? v * ep.keep_scale : 0.f;
        }
      } else if (ep.act == GANTTS_ACT_SIGMOID) {
        v = 1.f / (1.f + expf(-v));
      }
      *p = v;
    }
  }
}

// out[i] (+)= sum_z partial[z][i]   (deterministic: fixed summation order over z)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int nsplit, int64_t n,
                                     float* __restrict__ out, int accumulate) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(partial) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int z = 0; z < nsplit; ++z) {
      const float4 v = *reinterpret_cast<const float4*>(partial + (int64_t)z * n + i4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float4* o = reinterpret_cast<float4*>(out + i4);
    if (accumulate) {
      const float4 c = *o;
      s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
    }
    *o = s;
  } else {
    for (int64_t i = i4; i < i4 + 4 && i < n; ++i) {
      float s = 0.f;
      for (int z = 0; z < nsplit; ++z) s += partial[(int64_t)z * n + i];
      out[i] = accumulate ? out[i] + s : s;
    }
  }
}

// gz = gy * act'(y), derivative recovered from the saved layer OUTPUT y.
__global__ void act_bwd_kernel(const float* __restrict__ gy, int64_t gy_rs, const float* __restrict__ y,
                               int64_t y_rs, float* __restrict__ gz, int64_t M, int N, int act,
                               float slope, float keep_scale, int has_dropout) {
  const int64_t total = M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / N;
    int c = (int)(i - r * N);
    float g = gy[r * gy_rs + c], yy = y[r * y_rs + c];
    float o;
    if (act == GANTTS_ACT_LEAKY_DROPOUT) {
      if (has_dropout) o = yy > 0.f ? g * keep_scale : (yy < 0.f ? g * slope * keep_scale : 0.f);
      else o = yy > 0.f ? g : g * slope;
    } else if (act == GANTTS_ACT_SIGMOID) {
      o = g * yy * (1.f - yy);
    } else {
      o = g;
    }
    gz[i] = o;
  }
}

// Column sums of gz [M][N] (contiguous): partial[chunk][col], then finish.
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ gz, int64_t M, int N, int64_t rows_per_chunk,
                      float* __restrict__ partial) {
  __shared__ float sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const int64_t rbeg = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t rend = rbeg + rows_per_chunk < M ? rbeg + rows_per_chunk : M;
  float s = 0.f;
  if (col < N)
    for (int64_t r = rbeg + ry; r < rend; r += 8) s += gz[r * N + col];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][cx];
    partial[(int64_t)blockIdx.y * N + col] = t;
  }
}

}  // namespace gantts

using namespace gantts;

namespace gantts {

constexpr int COLSUM_CHUNKS = 64;

static int64_t splits_for(int64_t Mi, int Nj, int64_t Kk) {
  int64_t tiles = ((Mi + BM - 1) / BM) * ((Nj + BN - 1) / BN);
  if (tiles >= 296 || Kk < 2048) return 1;
  int64_t s = (592 + tiles - 1) / tiles;
  int64_t maxs = (Kk + 511) / 512;
  if (s > maxs) s = maxs;
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

size_t simt_workspace_bytes(int64_t M, int N, int K) {
  // split-K partials for gW (N x K outputs, reduction over M).
  int64_t s = splits_for(N, K, M);
  return (size_t)(s * (int64_t)N * K) * sizeof(float) + 256;
}

int simt_linear_fwd(const float* x, int64_t x_rs, const float* W, const float* bias, float* y,
                    int64_t y_rs, int64_t M, int N, int K, int act, float slope, float p,
                    uint64_t seed, cudaStream_t st) {
  Epilogue ep{};
  ep.bias = bias;
  ep.act = act;
  ep.slope = slope;
  ep.keep_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  ep.thresh = (act == GANTTS_ACT_LEAKY_DROPOUT && p > 0.f) ? (uint32_t)(p * 65536.f + 0.5f) : 0u;
  ep.seed = seed;
  dim3 grid((N + BN - 1) / BN, (unsigned)((M + BM - 1) / BM), 1);
  sgemm_kernel<<<grid, SIMT_THREADS, 0, st>>>(x, x_rs, 1, W, 1, K, y, y_rs, 0, M, N, K, K, ep);
  GANTTS_LAUNCH_CHECK("sgemm_kernel(fwd)");
  return GANTTS_OK;
}

int act_bwd(const float* gy, int64_t gy_rs, const float* y, int64_t y_rs, float* gz, int64_t M, int N,
            int act, float slope, float p, cudaStream_t st) {
  int64_t total = M * N;
  int nb = (int)((total + 1023) / 1024);
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
  act_bwd_kernel<<<nb, 256, 0, st>>>(gy, gy_rs, y, y_rs, gz, M, N, act, slope,
                                     p > 0.f ? 1.f / (1.f - p) : 1.f, p > 0.f ? 1 : 0);
  GANTTS_LAUNCH_CHECK("act_bwd_kernel");
  return GANTTS_OK;
}

int colsum(const float* gz, int64_t M, int N, float* gb, int accumulate, float* partial, cudaStream_t st) {
  int chunks = COLSUM_CHUNKS;
  int64_t rpc = (M + chunks - 1) / chunks;
  if (rpc < 1) rpc = 1;
  chunks = (int)((M + rpc - 1) / rpc);
  dim3 grid((N + 31) / 32, chunks);
  colsum_partial_kernel<<<grid, 256, 0, st>>>(gz, M, N, rpc, partial);
  GANTTS_LAUNCH_CHECK("colsum_partial_kernel");
  splitk_reduce_kernel<<<(N + 1023) / 1024, 256, 0, st>>>(partial, chunks, N, gb, accumulate);
  GANTTS_LAUNCH_CHECK("splitk_reduce_kernel(colsum)");
  return GANTTS_OK;
}

// gx = gz W ; gW (+)= gz^T x ; gz is contiguous [M][N].
int simt_linear_bwd_gemms(const float* gz, const float* x, int64_t x_rs, const float* W, float* gx,
                          int64_t gx_rs, float* gW, int64_t M, int N, int K, int accumulate,
                          float* ws, cudaStream_t st) {
  Epilogue ep{};
  if (gx) {
    dim3 grid((K + BN - 1) / BN, (unsigned)((M + BM - 1) / BM), 1);
    // C[m][k] = sum_n gz[m][n] * W[n][k]
    sgemm_kernel<<<grid, SIMT_THREADS, 0, st>>>(gz, N, 1, W, K, 1, gx, gx_rs, 0, M, K, N, N, ep);
    GANTTS_LAUNCH_CHECK("sgemm_kernel(gx)");
  }
  if (gW) {
    // C[n][k] = sum_m gz[m][n] * x[m][k]
    int64_t s = splits_for(N, K, M);
    int64_t kchunk = ((M + s - 1) / s + BK - 1) / BK * BK;
    s = (M + kchunk - 1) / kchunk;
    dim3 grid((K + BN - 1) / BN, (N + BM - 1) / BM, (unsigned)s);
    if (s == 1) {
      ep.accumulate = accumulate;
      sgemm_kernel<<<grid, SIMT_THREADS, 0, st>>>(gz, 1, N, x, x_rs, 1, gW, K, 0, N, K, M, kchunk, ep);
      GANTTS_LAUNCH_CHECK("sgemm_kernel(gW)");
    } else {
      sgemm_kernel<<<grid, SIMT_THREADS, 0, st>>>(gz, 1, N, x, x_rs, 1, ws, K, (int64_t)N * K, N, K, M,
                                                  kchunk, ep);
      GANTTS_LAUNCH_CHECK("sgemm_kernel(gW split)");
      int64_t n = (int64_t)N * K;
      splitk_reduce_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, st>>>(ws, (int)s, n, gW, accumulate);
      GANTTS_LAUNCH_CHECK("splitk_reduce_kernel(gW)");
    }
  }
  return GANTTS_OK;
}

}  // namespace gantts
