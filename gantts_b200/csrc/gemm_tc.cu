// tcgen05 tensor-core engine (GANTTS_ENGINE_TC): bf16x3 split GEMM with fp32 accumulation in TMEM.
//
// Every fp32 operand v is carried as two bf16 planes (hi = bf16(v), lo = bf16(v - hi)); a product
// a*b is evaluated as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi on the 5th-generation tensor cores
// (three tcgen05.mma per K=16 step into the same TMEM accumulator), which keeps ~2^-16 relative
// error per product (measured 8e-6 of output scale through the 4-layer generator) -- fp32-grade
// parity at 1.5x the cost of a single TF32 pass, where plain TF32/BF16 (2e-3) would miss the 1e-4 bar.
//
// One persistent warp-specialised kernel, two operand layouts:
//   MN = false : C[M][N] = A[M][K] * B[N][K]^T   both operands K-major   (layer forward, gx = gz W)
//   MN = true  : C[N][K] = A[M][N]^T * B[M][K]   both operands MN-major  (gW = gz^T x, split over M)
// warp 0: TMA producer   warp 1: TMEM allocator + single-thread MMA issuer   warps 2-5: epilogue
// (tcgen05.ld -> bias / LeakyReLU / dropout / sigmoid -> global).  smem ring of `num_stages`
// {A_hi, A_lo, B_hi, B_lo} tiles (SWIZZLE_128B), two TMEM accumulators so the epilogue of tile i
// overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gantts {

constexpr int TC_THREADS = 192;
constexpr int TC_BM = 128;          // MMA M (TMEM lanes)
constexpr int TC_BK = 64;           // reduction elements per stage (one 128B swizzle atom of bf16)
constexpr int TC_MAX_STAGES = 4;
constexpr uint32_t TC_A_PLANE = TC_BM * TC_BK * 2;   // 16 KB

struct GemmParams {
  int64_t rows_a;       // output rows   (extent of A's MN dimension)
  int cols_b;           // output cols   (extent of B's MN dimension)
  int64_t red;          // reduction extent
  int64_t red_chunk;    // reduction elements per z-slice (multiple of TC_BK)
  int num_a, num_b, num_z;
  int bn;               // MMA N, multiple of 16, <= 256
  int num_stages;
  uint32_t stage_bytes, b_plane_bytes, tx_bytes;
  uint32_t tmem_cols;
  float* C;
  int64_t ldc, c_zstride;
  int vec_ok;
  // epilogue
  const float* bias;
  int act;
  float slope, keep_scale;
  uint32_t thresh;
  uint64_t seed;
};

template <bool MN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + p.num_stages * p.stage_bytes;
  const uint32_t full0 = bar_base, empty0 = bar_base + 8 * TC_MAX_STAGES;
  const uint32_t tfull0 = bar_base + 16 * TC_MAX_STAGES, tempty0 = tfull0 + 16;
  const uint32_t tmem_slot = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull0 + 8 * a, 1);
      ptx::mbar_init(tempty0 + 8 * a, 4);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmAh);
    ptx::prefetch_tensormap(&tmAl);
    ptx::prefetch_tensormap(&tmBh);
    ptx::prefetch_tensormap(&tmBl);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int total_tiles = p.num_a * p.num_b * p.num_z;
  const int tiles_ab = p.num_a * p.num_b;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int z = tile / tiles_ab, rem = tile - z * tiles_ab;
        const int ta = rem / p.num_b, tb = rem - ta * p.num_b;
        const int64_t r_beg = (int64_t)z * p.red_chunk;
        const int64_t r_end = r_beg + p.red_chunk < p.red ? r_beg + p.red_chunk : p.red;
        const int a0 = ta * TC_BM, b0 = tb * p.bn;
        for (int64_t r0 = r_beg; r0 < r_end; r0 += TC_BK) {
          ptx::mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t fb = full0 + 8 * s;
          ptx::mbar_expect_tx(fb, p.tx_bytes);
          const uint32_t sa_hi = base + s * p.stage_bytes, sa_lo = sa_hi + TC_A_PLANE;
          const uint32_t sb_hi = sa_lo + TC_A_PLANE, sb_lo = sb_hi + p.b_plane_bytes;
          if (!MN) {
            ptx::tma_load_2d(sa_hi, &tmAh, fb, (int32_t)r0, a0);
            ptx::tma_load_2d(sa_lo, &tmAl, fb, (int32_t)r0, a0);
            ptx::tma_load_2d(sb_hi, &tmBh, fb, (int32_t)r0, b0);
            ptx::tma_load_2d(sb_lo, &tmBl, fb, (int32_t)r0, b0);
          } else {
            // 64-wide MN atoms, each [TC_BK reduction rows][128 B]
            for (int j = 0; j < TC_BM / 64; ++j) {
              ptx::tma_load_2d(sa_hi + j * 8192, &tmAh, fb, a0 + 64 * j, (int32_t)r0);
              ptx::tma_load_2d(sa_lo + j * 8192, &tmAl, fb, a0 + 64 * j, (int32_t)r0);
            }
            const int nb_atoms = (p.bn + 63) / 64;
            for (int j = 0; j < nb_atoms; ++j) {
              ptx::tma_load_2d(sb_hi + j * 8192, &tmBh, fb, b0 + 64 * j, (int32_t)r0);
              ptx::tma_load_2d(sb_lo + j * 8192, &tmBl, fb, b0 + 64 * j, (int32_t)r0);
            }
          }
          if (++s == (uint32_t)p.num_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer (one thread)
      const uint32_t idesc = ptx::make_idesc_bf16(TC_BM, p.bn, MN ? 1 : 0, MN ? 1 : 0);
      const uint32_t lbo = MN ? 8192u : 0u;
      const uint32_t kstep = MN ? 2048u : 32u;
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int z = tile / tiles_ab;
        const int64_t r_beg = (int64_t)z * p.red_chunk;
        const int64_t r_end = r_beg + p.red_chunk < p.red ? r_beg + p.red_chunk : p.red;
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        ptx::mbar_wait(tempty0 + 8 * acc, aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t first = 0;
        for (int64_t r0 = r_beg; r0 < r_end; r0 += TC_BK) {
          ptx::mbar_wait(full0 + 8 * s, ph);
          ptx::tc_fence_after();
          const uint32_t sa_hi = base + s * p.stage_bytes, sa_lo = sa_hi + TC_A_PLANE;
          const uint32_t sb_hi = sa_lo + TC_A_PLANE, sb_lo = sb_hi + p.b_plane_bytes;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t da_hi = ptx::make_smem_desc(sa_hi + k * kstep, lbo, 1024);
            const uint64_t da_lo = ptx::make_smem_desc(sa_lo + k * kstep, lbo, 1024);
            const uint64_t db_hi = ptx::make_smem_desc(sb_hi + k * kstep, lbo, 1024);
            const uint64_t db_lo = ptx::make_smem_desc(sb_lo + k * kstep, lbo, 1024);
            ptx::mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, first);
            first = 1;
            ptx::mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1);
            ptx::mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1);
          }
          ptx::mma_commit(empty0 + 8 * s);        // frees the smem stage once these MMAs retire
          if (++s == (uint32_t)p.num_stages) { s = 0; ph ^= 1; }
        }
        ptx::mma_commit(tfull0 + 8 * acc);        // accumulator ready for the epilogue
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue warps (2..5)
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int z = tile / tiles_ab, rem = tile - z * tiles_ab;
      const int ta = rem / p.num_b, tb = rem - ta * p.num_b;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      ptx::mbar_wait(tfull0 + 8 * acc, aph);
      ptx::tc_fence_after();
      const int64_t row = (int64_t)ta * TC_BM + q * 32 + lane;
      const int col0 = tb * p.bn;
      float* crow = p.C + (int64_t)z * p.c_zstride + row * p.ldc;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      for (int c = 0; c < p.bn; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(taddr + c, r);
        ptx::tmem_ld_wait();
        if (row < p.rows_a) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = col0 + c + j;
            float x = __uint_as_float(r[j]);
            if (p.bias && col < p.cols_b) x += p.bias[col];
            if (p.act == GANTTS_ACT_LEAKY_DROPOUT) {
              x = x > 0.f ? x : x * p.slope;
              if (p.thresh) {
                bool keep = dropout_keep(p.seed, (uint64_t)row * (uint64_t)p.cols_b + (uint64_t)col, p.thresh);
                x = keep ? x * p.keep_scale : 0.f;
              }
            } else if (p.act == GANTTS_ACT_SIGMOID) {
              x = 1.f / (1.f + expf(-x));
            }
            v[j] = x;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const int col = col0 + c + j;
            if (p.vec_ok && col + 3 < p.cols_b) {
              *reinterpret_cast<float4*>(crow + col) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col + e < p.cols_b) crow[col + e] = v[j + e];
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty0 + 8 * acc);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------- operand planes
// fp32 [rows][cols] (row stride rs) -> bf16 hi/lo planes [rows][pitch]; transpose: out[c][r] = in[r][c].
__global__ void split_planes_kernel(const float* __restrict__ src, int64_t rs, int64_t rows, int cols,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    int64_t pitch, int transpose) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols;
    int c = (int)(i - r * cols);
    float v = src[r * rs + c];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    int64_t o = transpose ? ((int64_t)c * pitch + r) : (r * pitch + c);
    hi[o] = h;
    lo[o] = l;
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int nsplit, int64_t n,
                                     float* __restrict__ out, int accumulate);

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D bf16 tensor [rows][cols] with row pitch `pitch` elements; box = {64 cols, box_rows}, SWIZZLE_128B.
static int make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t pitch, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("tc: cuTensorMapEncodeTiled entry point unavailable");
    return GANTTS_E_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld pitch=%lld box_rows=%d", (int)r,
              (long long)rows, (long long)cols, (long long)pitch, box_rows);
    return GANTTS_E_CUDA;
  }
  return GANTTS_OK;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline int64_t pitch_for(int64_t cols) { return (cols + 7) / 8 * 8; }
static inline size_t plane_bytes(int64_t rows, int64_t cols) {
  return ((size_t)rows * pitch_for(cols) * 2 + 255) / 256 * 256;
}

struct Planes {
  __nv_bfloat16 *hi, *lo;
  int64_t rows, cols, pitch;
};

static Planes carve_planes(char*& cur, int64_t rows, int64_t cols) {
  Planes pl;
  pl.rows = rows;
  pl.cols = cols;
  pl.pitch = pitch_for(cols);
  pl.hi = reinterpret_cast<__nv_bfloat16*>(cur);
  cur += plane_bytes(rows, cols);
  pl.lo = reinterpret_cast<__nv_bfloat16*>(cur);
  cur += plane_bytes(rows, cols);
  return pl;
}

static int launch_split(const float* src, int64_t rs, int64_t rows, int cols, const Planes& pl, int transpose,
                        cudaStream_t st) {
  int64_t total = rows * cols;
  int nb = (int)((total + 1023) / 1024);
  if (nb > num_sms() * 8) nb = num_sms() * 8;
  if (nb < 1) nb = 1;
  split_planes_kernel<<<nb, 256, 0, st>>>(src, rs, rows, cols, pl.hi, pl.lo, pl.pitch, transpose);
  GANTTS_LAUNCH_CHECK("split_planes_kernel");
  return GANTTS_OK;
}

static int pick_bn(int n) {
  int bn = (n + 15) / 16 * 16;
  if (bn <= 256) return bn;
  int tiles = (n + 255) / 256;
  bn = ((n + tiles - 1) / tiles + 15) / 16 * 16;
  return bn;
}

struct EpiArgs {
  const float* bias = nullptr;
  int act = GANTTS_ACT_NONE;
  float slope = 0.f, p = 0.f;
  uint64_t seed = 0;
};

// C[rows_a][cols_b] = A * B^T (K-major planes A [rows_a][red], B [cols_b][red]).
static int launch_gemm_kk(const Planes& A, const Planes& B, float* C, int64_t ldc, const EpiArgs& e,
                          cudaStream_t st) {
  GemmParams p{};
  p.rows_a = A.rows;
  p.cols_b = (int)B.rows;
  p.red = A.cols;
  p.red_chunk = (p.red + TC_BK - 1) / TC_BK * TC_BK;
  p.bn = pick_bn(p.cols_b);
  p.num_a = (int)((p.rows_a + TC_BM - 1) / TC_BM);
  p.num_b = (p.cols_b + p.bn - 1) / p.bn;
  p.num_z = 1;
  p.b_plane_bytes = ((uint32_t)p.bn * 128 + 1023) / 1024 * 1024;
  p.stage_bytes = 2 * TC_A_PLANE + 2 * p.b_plane_bytes;
  p.tx_bytes = 2 * TC_A_PLANE + 2 * (uint32_t)p.bn * 128;
  p.num_stages = (int)((220 * 1024) / p.stage_bytes);
  if (p.num_stages > TC_MAX_STAGES) p.num_stages = TC_MAX_STAGES;
  p.tmem_cols = 512;
  p.C = C;
  p.ldc = ldc;
  p.c_zstride = 0;
  p.vec_ok = ((ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (p.bn % 4) == 0) ? 1 : 0;
  p.bias = e.bias;
  p.act = e.act;
  p.slope = e.slope;
  p.keep_scale = e.p > 0.f ? 1.f / (1.f - e.p) : 1.f;
  p.thresh = (e.act == GANTTS_ACT_LEAKY_DROPOUT && e.p > 0.f) ? (uint32_t)(e.p * 65536.f + 0.5f) : 0u;
  p.seed = e.seed;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, A.hi, A.rows, A.cols, A.pitch, TC_BM))) return rc;
  if ((rc = make_map(&mAl, A.lo, A.rows, A.cols, A.pitch, TC_BM))) return rc;
  if ((rc = make_map(&mBh, B.hi, B.rows, B.cols, B.pitch, p.bn))) return rc;
  if ((rc = make_map(&mBl, B.lo, B.rows, B.cols, B.pitch, p.bn))) return rc;
  const size_t smem = (size_t)p.num_stages * p.stage_bytes + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    GANTTS_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    attr = true;
  }
  int total = p.num_a * p.num_b;
  int grid = total < num_sms() ? total : num_sms();
  prof_begin(PROF_GEMM_KK, 2.0 * (double)p.rows_a * p.cols_b * (double)p.red, st);
  gemm_bf16x3_kernel<false><<<grid, TC_THREADS, smem, st>>>(mAh, mAl, mBh, mBl, p);
  prof_end(st);
  GANTTS_LAUNCH_CHECK("gemm_bf16x3_kernel<K-major>");
  return GANTTS_OK;
}

// C[n][k] (+)= sum_m A[m][n] * B[m][k]  (MN-major planes A [red][rows_a], B [red][cols_b]);
// split over the reduction, partials in `partial`, reduced deterministically into C (ld = cols_b).
static size_t mn_partial_bytes(int64_t red, int rows_a, int cols_b, int* splits_out, int64_t* chunk_out) {
  int bn = pick_bn(cols_b);
  int tiles = ((rows_a + TC_BM - 1) / TC_BM) * ((cols_b + bn - 1) / bn);
  int64_t blocks = (red + TC_BK - 1) / TC_BK;
  int64_t splits = num_sms() / tiles;
  if (splits < 1) splits = 1;
  if (splits > blocks) splits = blocks;
  int64_t chunk = (blocks + splits - 1) / splits * TC_BK;
  splits = (red + chunk - 1) / chunk;
  if (splits_out) *splits_out = (int)splits;
  if (chunk_out) *chunk_out = chunk;
  return (size_t)splits * rows_a * cols_b * sizeof(float);
}

static int launch_gemm_mn(const Planes& A, const Planes& B, float* C, int accumulate, float* partial,
                          cudaStream_t st) {
  GemmParams p{};
  p.rows_a = A.cols;
  p.cols_b = (int)B.cols;
  p.red = A.rows;
  int splits;
  int64_t chunk;
  mn_partial_bytes(p.red, (int)p.rows_a, p.cols_b, &splits, &chunk);
  p.red_chunk = chunk;
  p.num_z = splits;
  p.bn = pick_bn(p.cols_b);
  p.num_a = (int)((p.rows_a + TC_BM - 1) / TC_BM);
  p.num_b = (p.cols_b + p.bn - 1) / p.bn;
  const int nb_atoms = (p.bn + 63) / 64;
  p.b_plane_bytes = (uint32_t)nb_atoms * 8192;
  p.stage_bytes = 2 * TC_A_PLANE + 2 * p.b_plane_bytes;
  p.tx_bytes = p.stage_bytes;
  p.num_stages = (int)((220 * 1024) / p.stage_bytes);
  if (p.num_stages > TC_MAX_STAGES) p.num_stages = TC_MAX_STAGES;
  p.tmem_cols = 512;
  const bool direct = (splits == 1 && !accumulate);
  p.C = direct ? C : partial;
  p.ldc = p.cols_b;
  p.c_zstride = (int64_t)p.rows_a * p.cols_b;
  p.vec_ok = ((p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (p.bn % 4) == 0) ? 1 : 0;
  p.act = GANTTS_ACT_NONE;
  p.keep_scale = 1.f;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, A.hi, A.rows, A.cols, A.pitch, TC_BK))) return rc;
  if ((rc = make_map(&mAl, A.lo, A.rows, A.cols, A.pitch, TC_BK))) return rc;
  if ((rc = make_map(&mBh, B.hi, B.rows, B.cols, B.pitch, TC_BK))) return rc;
  if ((rc = make_map(&mBl, B.lo, B.rows, B.cols, B.pitch, TC_BK))) return rc;
  const size_t smem = (size_t)p.num_stages * p.stage_bytes + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    GANTTS_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    attr = true;
  }
  int total = p.num_a * p.num_b * p.num_z;
  int grid = total < num_sms() ? total : num_sms();
  prof_begin(PROF_GEMM_MN, 2.0 * (double)p.rows_a * p.cols_b * (double)p.red, st);
  gemm_bf16x3_kernel<true><<<grid, TC_THREADS, smem, st>>>(mAh, mAl, mBh, mBl, p);
  prof_end(st);
  GANTTS_LAUNCH_CHECK("gemm_bf16x3_kernel<MN-major>");
  if (!direct) {
    int64_t n = (int64_t)p.rows_a * p.cols_b;
    splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, splits, n, C, accumulate);
    GANTTS_LAUNCH_CHECK("splitk_reduce_kernel(tc gW)");
  }
  return GANTTS_OK;
}

size_t tc_linear_workspace_bytes(int64_t M, int N, int K) {
  // forward: x planes + W planes ; backward: gz planes + x planes + W^T planes + gW partials
  size_t fwd = 2 * plane_bytes(M, K) + 2 * plane_bytes(N, K);
  size_t bwd = 2 * plane_bytes(M, N) + 2 * plane_bytes(M, K) + 2 * plane_bytes(K, N) +
               mn_partial_bytes(M, N, K, nullptr, nullptr) + 256;
  return (fwd > bwd ? fwd : bwd) + 1024;
}

int tc_linear_fwd(const float* x, int64_t x_rs, const float* W, const float* bias, float* y, int64_t y_rs,
                  int64_t M, int N, int K, int act, float slope, float p, uint64_t seed, void* ws,
                  size_t ws_bytes, cudaStream_t st) {
  size_t need = tc_linear_workspace_bytes(M, N, K);
  if (!ws || ws_bytes < need) {
    set_error("tc_linear_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
    return GANTTS_E_WORKSPACE;
  }
  char* cur = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  Planes X = carve_planes(cur, M, K), Wp = carve_planes(cur, N, K);
  int rc;
  if ((rc = launch_split(x, x_rs, M, K, X, 0, st))) return rc;
  if ((rc = launch_split(W, K, N, K, Wp, 0, st))) return rc;
  EpiArgs e;
  e.bias = bias;
  e.act = act;
  e.slope = slope;
  e.p = p;
  e.seed = seed;
  return launch_gemm_kk(X, Wp, y, y_rs, e, st);
}

int tc_linear_bwd_gemms(const float* gz, const float* x, int64_t x_rs, const float* W, float* gx,
                        int64_t gx_rs, float* gW, int64_t M, int N, int K, int accumulate, void* ws,
                        size_t ws_bytes, cudaStream_t st) {
  size_t need = tc_linear_workspace_bytes(M, N, K);
  if (!ws || ws_bytes < need) {
    set_error("tc_linear_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
    return GANTTS_E_WORKSPACE;
  }
  char* cur = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  Planes G = carve_planes(cur, M, N);
  Planes X = carve_planes(cur, M, K);
  Planes Wt = carve_planes(cur, K, N);
  float* partial = reinterpret_cast<float*>(cur);
  int rc;
  if ((rc = launch_split(gz, N, M, N, G, 0, st))) return rc;
  if (gx) {
    if ((rc = launch_split(W, K, N, K, Wt, 1, st))) return rc;      // Wt[k][n] = W[n][k]
    EpiArgs e;
    if ((rc = launch_gemm_kk(G, Wt, gx, gx_rs, e, st))) return rc;  // gx[m][k] = sum_n gz[m][n] W[n][k]
  }
  if (gW) {
    if ((rc = launch_split(x, x_rs, M, K, X, 0, st))) return rc;
    if ((rc = launch_gemm_mn(G, X, gW, accumulate, partial, st))) return rc;
  }
  return GANTTS_OK;
}

}  // namespace gantts
