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
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gantts {

constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int TC_BM = 128;          // MMA M (TMEM lanes)
constexpr int TC_BK = 64;           // default reduction elements per stage (p.bk: 64, or 32 for deeper pipelines)
constexpr int TC_MAX_STAGES = 8;
constexpr uint32_t TC_BIAS_SMEM = 4096;              // staged bias vector (<= 1024 columns)
constexpr uint32_t TC_F32_STAGE_SMEM = 16 * 2048;    // opt-in transpose scratch of the fp32 epilogue (2 KB per warp)

// Epilogue flavours (template parameter of the kernel).
constexpr int EPI_F32 = 0;          // bias + {none | leaky+dropout | sigmoid} -> fp32 C (optionally +=)
constexpr int EPI_PLANES_FWD = 1;   // bias + leaky + dropout -> bf16 hi/lo planes (next layer's operand)
constexpr int EPI_PLANES_BWD = 2;   // acc * act'(saved output hi plane) -> bf16 hi/lo planes (gz)

struct GemmParams {
  int64_t rows_a;       // output rows   (extent of A's MN dimension)
  int cols_b;           // output cols   (extent of B's MN dimension)
  int64_t red;          // reduction extent
  int64_t red_chunk;    // reduction elements per z-slice (multiple of TC_BK)
  int num_a, num_b, num_z;
  int bn;               // MMA N, multiple of 64, <= 256
  int num_stages;
  uint32_t stage_bytes, b_plane_bytes, tx_bytes;
  int64_t row0;         // global index of the first output row (dropout keys use global rows when a launch covers a row window)
  uint32_t bres_bytes;  // pair kernel, BRES: bytes of the resident B operand in front of the stage ring
  int bk;               // reduction elements per smem stage: 64 (SWIZZLE_128B K-major rows) or 32 (SWIZZLE_64B)
  uint32_t a_plane;     // bytes of one A plane tile in a stage
  uint32_t atom_bytes;  // MN-major: bytes of one 64-wide atom ([bk rows][128 B]) = its LBO
  uint32_t sbo, kstep, desc_layout;
  uint32_t tmem_cols;
  // EPI_F32 output
  float* C;
  int64_t ldc, c_zstride;
  int vec_ok, accumulate;
  // planes output (EPI_PLANES_*)
  __nv_bfloat16 *out_hi, *out_lo;
  int64_t out_pitch;
  // activation-derivative code plane: 2 bits per element (bit0 = zero/dropped, bit1 = negative), one
  // uint32 per (row, 16 columns).  Written by EPI_PLANES_FWD, read (prefetched) by EPI_PLANES_BWD.
  uint32_t* code;
  int64_t code_pitch;   // words per row
  uint32_t bias_off;    // byte offset (from the aligned smem base) of the staged bias vector, 0 = none
  uint32_t f32_stage_off;   // EPI_F32 with an unaligned row stride: byte offset of the per-warp transpose scratch
                            // (TC_EPI_WARPS x 2 KB), 0 = store straight from registers
  uint32_t dbg;         // experiment switches (env GANTTS_B200_DBG): 1 no plane stores, 2 no dropout, 4 no code
  // MN-major only: column sums of A (= bias gradient) via an extra N=16 MMA against a tile of ones
  float* db;            // [num_z][rows_a] partial sums, or null
  uint32_t ones_off;    // byte offset of the 8 KB all-ones bf16 tile from the aligned smem base
  // epilogue math
  const float* bias;
  int act;
  float slope, keep_scale;
  uint32_t thresh;
  uint64_t seed;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// Split 8 fp32 values into packed bf16 hi/lo words (4 words each).
__device__ __forceinline__ void split8(const float* v, uint32_t* h, uint32_t* l) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = v[2 * i], b = v[2 * i + 1];
    const uint32_t hp = pack_bf16x2(a, b);
    const float ah = __uint_as_float(hp << 16), bh = __uint_as_float(hp & 0xffff0000u);
    h[i] = hp;
    l[i] = pack_bf16x2(a - ah, b - bh);
  }
}

// 16 fp32 values of one row -> hi/lo planes.  Fast path: one 256-bit store (a full 32 B sector) per
// plane (sm_100 STG.256); `pitch` is a multiple of 16 elements and col a multiple of 16, so the
// address is 32-byte aligned.  Tail: 16-byte stores up to the pitch.
__device__ __forceinline__ void store_planes16(const float* v, __nv_bfloat16* hi, __nv_bfloat16* lo, int col,
                                               int64_t pitch) {
  uint32_t h[8], l[8];
  split8(v, h, l);
  split8(v + 8, h + 4, l + 4);
  if (col + 16 <= pitch) {
    st_global_256(hi, h);
    st_global_256(lo, l);
  } else if (col + 8 <= pitch) {
    *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// Dropout of the 16 values of one row at columns [col, col+16), col % 16 == 0: v = keep ? v * scale : 0
// (common.cuh: one hash chain per four columns, one unsigned compare per element).
__device__ __forceinline__ void dropout16(float (&v)[16], uint64_t seed, uint32_t row, int n_cols, int col, uint32_t thresh,
                                          float scale) {
  const uint32_t quarter_n = (uint32_t)(n_cols + 3) >> 2, limit = drop_limit(thresh);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const DropBits d = dropout_quad_bits(seed, row, quarter_n, ((uint32_t)col >> 2) + q);
    v[4 * q] = (d.a << 16) > limit ? v[4 * q] * scale : 0.f;
    v[4 * q + 1] = d.a > limit ? v[4 * q + 1] * scale : 0.f;
    v[4 * q + 2] = (d.b << 16) > limit ? v[4 * q + 2] * scale : 0.f;
    v[4 * q + 3] = d.b > limit ? v[4 * q + 3] * scale : 0.f;
  }
}

// EPI_F32 for outputs whose row stride is not a multiple of 4 floats (y_hat: ld 187, the discriminator input
// gradient: ld 58, weight-gradient partials of 425- and 58-wide layers).  Straight from registers a thread can
// only issue 16 scalar stores per chunk and a warp store touches 32 rows = 32 sectors (the 512->187 layer takes
// 58 us against a 14 us floor, profiles/r01_gemm_per_launch.md).  Here the warp's 32 x 16 tile goes through a
// private 2 KB shared-memory scratch (float4 writes, XOR-swizzled: conflict-free) and is written back with lanes
// 0-15 / 16-31 covering two whole 64-byte row segments per store instruction.  Same arithmetic, same order.
// Default on (GANTTS_B200_F32_STAGE=0 disables): measured on B200 at cfg2, the step drops 1.381 -> 1.327 ms and the
// K-major GEMM family 0.795 -> 0.701 ms (profiles/r02_gemm_experiments.md); the whole -m gpu suite passes either way.
__device__ __forceinline__ void epilogue_f32_staged(const GemmParams& p, const uint32_t (&r)[16], int64_t row0,
                                                    int lane, int col, int z, const float* __restrict__ bias_s,
                                                    float* scr) {
  const int64_t row = row0 + lane;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  if (p.bias) {
    if (bias_s) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += bias_s[col + j];
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col + j < p.cols_b) v[j] += __ldg(p.bias + col + j);
    }
  }
  if (p.act == GANTTS_ACT_LEAKY_DROPOUT) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * p.slope);
    if (p.thresh) dropout16(v, p.seed, (uint32_t)(row + p.row0), p.cols_b, col, p.thresh, p.keep_scale);
  } else if (p.act == GANTTS_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
  }
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    *reinterpret_cast<float4*>(scr + lane * 16 + 4 * (k ^ sw)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
  float* cbase = p.C + (int64_t)z * p.c_zstride;
  const int j = lane & 15, hr = lane >> 4;
  const bool col_ok = col + j < p.cols_b;
  float* q0 = cbase + (row0 + hr) * p.ldc + col + j;
  // accumulate (the discriminator's input gradient added into its window of g_static): all 16 loads are issued before
  // the first store -- interleaved `*q = *q + val` serialises 16 load -> store round trips per warp (the compiler must
  // assume the store aliases the next load): 31.6 us per launch against a 6 us HBM floor (profiles/r02_gemm_per_launch.md)
  // (two halves of 8: 16 live loads on top of the tile cost spills under the 96-register cap)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float old[8];
    if (p.accumulate) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        old[i] = (col_ok && row0 + 2 * (8 * h + i) + hr < p.rows_a) ? __ldcg(q0 + (int64_t)(2 * (8 * h + i)) * p.ldc) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = 2 * (8 * h + i) + hr;
      const float val = scr[rr * 16 + 4 * ((j >> 2) ^ ((rr >> 1) & 3)) + (j & 3)];
      if (col_ok && row0 + rr < p.rows_a) q0[(int64_t)(2 * (8 * h + i)) * p.ldc] = p.accumulate ? old[i] + val : val;
    }
  }
  __syncwarp();
}

// One 16-column chunk of one output row: registers (fp32 accumulators) -> global.
// Returns the derivative code word of the chunk (EPI_PLANES_FWD); `code_in` is the saved word (BWD).
template <int EPI>
__device__ __forceinline__ uint32_t epilogue_chunk16(const GemmParams& p, const uint32_t (&r)[16], int64_t row,
                                                     int col, int z, const float* __restrict__ bias_s,
                                                     uint32_t code_in) {
  uint32_t code = 0;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  if (EPI == EPI_F32) {
    float* crow = p.C + (int64_t)z * p.c_zstride + row * p.ldc;
    if (p.bias) {
      if (bias_s) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += bias_s[col + j];
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col + j < p.cols_b) v[j] += __ldg(p.bias + col + j);
      }
    }
    if (p.act == GANTTS_ACT_LEAKY_DROPOUT) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * p.slope);
      if (p.thresh) dropout16(v, p.seed, (uint32_t)(row + p.row0), p.cols_b, col, p.thresh, p.keep_scale);
    } else if (p.act == GANTTS_ACT_SIGMOID) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
    }
    if (p.vec_ok == 2 && col + 15 < p.cols_b) {
      uint32_t w0[8], w1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { w0[j] = __float_as_uint(v[j]); w1[j] = __float_as_uint(v[8 + j]); }
      st_global_256(crow + col, w0);
      st_global_256(crow + col + 8, w1);
    } else
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const int c = col + j;
      if (p.vec_ok && c + 3 < p.cols_b) {
        *reinterpret_cast<float4*>(crow + c) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e < p.cols_b) crow[c + e] = p.accumulate ? crow[c + e] + v[j + e] : v[j + e];
      }
    }
  } else if (EPI == EPI_PLANES_FWD) {
    // reference gantts/models.py:137-139: Dropout(LeakyReLU(Linear(x)))
    if (bias_s) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col + j);   // broadcast LDS.128
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col + j < p.cols_b) v[j] += __ldg(p.bias + col + j);
    }
    // LeakyReLU as compare + select (torch's x > 0 ? x : x * slope), dropout keep as one unsigned compare per element
    // (common.cuh); the 2-bit derivative code (bit 0: derivative 0, bit 1: negative side) is OR-ed in under the very
    // predicates those compares produce: 2 instructions per element where `v == 0` / `v < 0` tests on the result cost 6
    // (the epilogue is issue-bound: profiles/r02_gemm_experiments.md).  With dropout on, bit 0 = "dropped"; a kept element
    // whose pre-activation is exactly 0 decodes as the positive side (measure zero).  Without dropout bit 0 = (v == 0).
    if (p.thresh) {
      const uint32_t quarter_n = (uint32_t)(p.cols_b + 3) >> 2, limit = drop_limit(p.thresh);
      const uint32_t grow = (uint32_t)(row + p.row0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const DropBits d = dropout_quad_bits(p.seed, grow, quarter_n, ((uint32_t)col >> 2) + q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * q + e;
          const uint32_t w = e < 2 ? d.a : d.b;
          const bool keep = ((e & 1) ? w : (w << 16)) > limit;
          const bool neg = v[j] < 0.f;
          const float a = neg ? v[j] * p.slope : v[j];
          v[j] = keep ? a * p.keep_scale : 0.f;
          if (!keep) code |= 1u << (2 * j);
          if (neg) code |= 2u << (2 * j);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool neg = v[j] < 0.f;
        v[j] = neg ? v[j] * p.slope : v[j];
        if (v[j] == 0.f) code |= 1u << (2 * j);
        if (neg) code |= 2u << (2 * j);
      }
    }
    __nv_bfloat16* oh = p.out_hi + row * p.out_pitch + col;
    __nv_bfloat16* ol = p.out_lo + row * p.out_pitch + col;
    store_planes16(v, oh, ol, col, p.out_pitch);
  } else {  // EPI_PLANES_BWD: gz = g * act'(h), derivative class from the saved 2-bit code
    const float dpos = p.keep_scale, dneg = p.slope * p.keep_scale;
    const float dzero = p.thresh ? 0.f : p.slope;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t cj = (code_in >> (2 * j)) & 3u;
      v[j] *= (cj & 1u) ? dzero : ((cj & 2u) ? dneg : dpos);
    }
    __nv_bfloat16* oh = p.out_hi + row * p.out_pitch + col;
    __nv_bfloat16* ol = p.out_lo + row * p.out_pitch + col;
    store_planes16(v, oh, ol, col, p.out_pitch);
  }
  return code;
}

// CL = thread-block cluster size along the output rows (1 or 2).  With CL = 2 the two CTAs of a cluster
// work on adjacent row tiles of the SAME column tile: each loads its own A tile and HALF of the B tile,
// multicast to both (halves the dominant L2 -> SM traffic, the re-fetch of the weights per row tile).
template <bool MN, int EPI, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t ones_base = base + p.ones_off;                       // MN only (8 KB), else unused
  const uint32_t bar_base = base + p.num_stages * p.stage_bytes + (MN ? 8192u : 0u);
  const uint32_t full0 = bar_base, empty0 = bar_base + 8 * TC_MAX_STAGES;
  const uint32_t tfull0 = bar_base + 16 * TC_MAX_STAGES, tempty0 = tfull0 + 16;
  const uint32_t tmem_slot = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, CL);          // released by the MMA issuer of every CTA in the cluster
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull0 + 8 * a, 1);
      ptx::mbar_init(tempty0 + 8 * a, TC_EPI_WARPS);
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
  // PDL: let the next GEMM of the stream set itself up while this one drains; nothing produced by the
  // previous kernel is read before the dependency wait.
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const float* bias_s = nullptr;
  if (p.bias_off) {
    float* bs = reinterpret_cast<float*>(smem_raw + (base + p.bias_off - raw));
    const int nb = p.num_b * p.bn;                       // zero-padded to whole tiles
    for (int i = threadIdx.x; i < nb; i += TC_THREADS) bs[i] = i < p.cols_b ? p.bias[i] : 0.f;
    bias_s = bs;
  }
  if (MN && p.db) {
    // all-ones bf16 tile (any swizzle of a constant tile is the same tile)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_raw + (ones_base - raw));
    for (int i = threadIdx.x; i < 8192 / 4; i += TC_THREADS) ones[i] = 0x3F803F80u;
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync();                 // peers' barriers are initialised before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  // Work decomposition.  CL = 1: CTA b takes tiles b, b+grid, ...  CL = 2: cluster c takes PAIR tiles
  // c, c+nclusters, ...; a pair tile is two adjacent row tiles of one column tile, CTA rank r gets
  // row tile 2*pa + r.  Both CTAs of a cluster walk the same sequence in lockstep.
  const int crank = CL > 1 ? (int)ptx::cluster_ctarank() : 0;
  const int num_a_units = CL > 1 ? (p.num_a + CL - 1) / CL : p.num_a;
  const int tiles_ab = num_a_units * p.num_b;
  const int total_tiles = tiles_ab * p.num_z;
  const int first_tile = CL > 1 ? (int)blockIdx.x / CL : (int)blockIdx.x;
  const int tile_step = CL > 1 ? (int)gridDim.x / CL : (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      uint32_t s = 0, ph = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int z = tile / tiles_ab, rem = tile - z * tiles_ab;
        const int ta = (rem / p.num_b) * CL + crank, tb = rem % p.num_b;
        const int64_t r_beg = (int64_t)z * p.red_chunk;
        const int64_t r_end = r_beg + p.red_chunk < p.red ? r_beg + p.red_chunk : p.red;
        const int a0 = ta * TC_BM, b0 = tb * p.bn;
        for (int64_t r0 = r_beg; r0 < r_end; r0 += p.bk) {
          ptx::mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t fb = full0 + 8 * s;
          ptx::mbar_expect_tx(fb, p.tx_bytes);
          const uint32_t sa_hi = base + s * p.stage_bytes, sa_lo = sa_hi + p.a_plane;
          const uint32_t sb_hi = sa_lo + p.a_plane, sb_lo = sb_hi + p.b_plane_bytes;
          if (!MN && CL > 1) {
            // own A tile; own half of the B tile (box = bn/2 rows) multicast to both CTAs
            const uint32_t hoff = (uint32_t)crank * (uint32_t)(p.bn / 2) * (uint32_t)(p.bk * 2);
            const int bh0 = b0 + crank * (p.bn / 2);
            ptx::tma_load_2d(sa_hi, &tmAh, fb, (int32_t)r0, a0);
            ptx::tma_load_2d_mcast(sb_hi + hoff, &tmBh, fb, (int32_t)r0, bh0, (uint16_t)0x3);
            ptx::tma_load_2d(sa_lo, &tmAl, fb, (int32_t)r0, a0);
            ptx::tma_load_2d_mcast(sb_lo + hoff, &tmBl, fb, (int32_t)r0, bh0, (uint16_t)0x3);
          } else if (!MN) {
            ptx::tma_load_2d(sa_hi, &tmAh, fb, (int32_t)r0, a0);
            ptx::tma_load_2d(sb_hi, &tmBh, fb, (int32_t)r0, b0);
            ptx::tma_load_2d(sa_lo, &tmAl, fb, (int32_t)r0, a0);
            ptx::tma_load_2d(sb_lo, &tmBl, fb, (int32_t)r0, b0);
          } else {
            // 64-wide MN atoms, each [bk reduction rows][128 B]
            for (int j = 0; j < TC_BM / 64; ++j) {
              ptx::tma_load_2d(sa_hi + j * p.atom_bytes, &tmAh, fb, a0 + 64 * j, (int32_t)r0);
              ptx::tma_load_2d(sa_lo + j * p.atom_bytes, &tmAl, fb, a0 + 64 * j, (int32_t)r0);
            }
            const int nb_atoms = (p.bn + 63) / 64;
            for (int j = 0; j < nb_atoms; ++j) {
              ptx::tma_load_2d(sb_hi + j * p.atom_bytes, &tmBh, fb, b0 + 64 * j, (int32_t)r0);
              ptx::tma_load_2d(sb_lo + j * p.atom_bytes, &tmBl, fb, b0 + 64 * j, (int32_t)r0);
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
      const uint32_t lbo = MN ? p.atom_bytes : 0u;
      const uint32_t kstep = p.kstep, sbo = p.sbo, lay = p.desc_layout;
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
        const int z = tile / tiles_ab;
        const int64_t r_beg = (int64_t)z * p.red_chunk;
        const int64_t r_end = r_beg + p.red_chunk < p.red ? r_beg + p.red_chunk : p.red;
        // K-major: two accumulators (epilogue of tile i overlaps MMAs of tile i+1); MN-major: one
        // accumulator at columns [0,256) plus the bias-gradient accumulator at columns [256,272).
        const int acc = MN ? 0 : (it & 1);
        const uint32_t aph = MN ? (it & 1) : ((it >> 1) & 1);
        ptx::mbar_wait(tempty0 + 8 * acc, aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const int rem_i = tile - z * tiles_ab;   // (MN kernels run with CL = 1)
        const bool do_db = MN && p.db != nullptr && (rem_i % p.num_b) == 0;
        const uint32_t idesc_db = ptx::make_idesc_bf16(TC_BM, 16, 1, 1);
        const uint64_t d_ones = ptx::make_smem_desc(ones_base, 8192u, 1024);
        uint32_t first = 0;
        for (int64_t r0 = r_beg; r0 < r_end; r0 += p.bk) {
          ptx::mbar_wait(full0 + 8 * s, ph);
          ptx::tc_fence_after();
          const uint32_t sa_hi = base + s * p.stage_bytes, sa_lo = sa_hi + p.a_plane;
          const uint32_t sb_hi = sa_lo + p.a_plane, sb_lo = sb_hi + p.b_plane_bytes;
          for (int k = 0; k < ((p.dbg & 16) ? 0 : p.bk / 16); ++k) {
            const uint64_t da_hi = ptx::make_smem_desc(sa_hi + k * kstep, lbo, sbo, lay);
            const uint64_t da_lo = ptx::make_smem_desc(sa_lo + k * kstep, lbo, sbo, lay);
            const uint64_t db_hi = ptx::make_smem_desc(sb_hi + k * kstep, lbo, sbo, lay);
            const uint64_t db_lo = ptx::make_smem_desc(sb_lo + k * kstep, lbo, sbo, lay);
            ptx::mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, first);
            if (do_db) {
              ptx::mma_bf16_ss(tmem_base + 256, da_hi, d_ones, idesc_db, first);
              ptx::mma_bf16_ss(tmem_base + 256, da_lo, d_ones, idesc_db, 1);
            }
            first = 1;
            ptx::mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1);
            ptx::mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1);
          }
          // frees the smem stage (in every CTA of the cluster: the peer multicasts into it too)
          if (CL > 1) ptx::mma_commit_mcast(empty0 + 8 * s, (uint16_t)0x3);
          else ptx::mma_commit(empty0 + 8 * s);
          if (++s == (uint32_t)p.num_stages) { s = 0; ph ^= 1; }
        }
        ptx::mma_commit(tfull0 + 8 * acc);        // accumulator ready for the epilogue
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue warps (2..17)
    // warp w may access TMEM lanes 32*(w%4)..+31.  Sixteen warps: four per lane quarter, each taking
    // one quarter of the tile's columns, i.e. four warps per SM sub-partition to hide the TMEM-load
    // and shared-memory latencies of the element-wise epilogue by thread-level parallelism.
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;                  // column quarter 0..3
    const int cw = p.bn >> 2;                           // columns per warp (multiple of 16)
    const int cbeg = chalf * cw, cend = cbeg + cw;
    float* stage_scr = (EPI == EPI_F32 && p.f32_stage_off)
                           ? reinterpret_cast<float*>(smem_raw + (base + p.f32_stage_off - raw)) + (warp - 2) * 512
                           : nullptr;
    uint32_t code_next[4] = {0u, 0u, 0u, 0u};
    if (EPI == EPI_PLANES_BWD && first_tile < total_tiles) {
      const int nrem = first_tile % tiles_ab;
      const int64_t nrow = (int64_t)((nrem / p.num_b) * CL + crank) * TC_BM + q * 32 + lane;
      const int ncol0 = (nrem % p.num_b) * p.bn;
      const uint32_t* cp = p.code + nrow * p.code_pitch + ((ncol0 + cbeg) >> 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        code_next[i] = (nrow < p.rows_a && cbeg + 16 * i < cend && ncol0 + cbeg + 16 * i < p.cols_b)
                           ? __ldg(cp + i) : 0u;
    }
    int it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
      const int z = tile / tiles_ab, rem = tile - z * tiles_ab;
      const int ta = (rem / p.num_b) * CL + crank, tb = rem % p.num_b;
      const int acc = MN ? 0 : (it & 1);
      const uint32_t aph = MN ? (it & 1) : ((it >> 1) & 1);
      const int64_t row = (int64_t)ta * TC_BM + q * 32 + lane;
      const int col0 = tb * p.bn;
      const bool row_ok = row < p.rows_a;
      // EPI_PLANES_BWD: this thread's derivative codes (4 words) were prefetched one tile ahead.
      uint32_t codes[4];
      if (EPI == EPI_PLANES_BWD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) codes[i] = code_next[i];
        const int ntile = tile + tile_step;
        if (ntile < total_tiles) {
          const int nrem = ntile % tiles_ab;
          const int64_t nrow = (int64_t)((nrem / p.num_b) * CL + crank) * TC_BM + q * 32 + lane;
          const int ncol0 = (nrem % p.num_b) * p.bn;
          const uint32_t* cp = p.code + nrow * p.code_pitch + ((ncol0 + cbeg) >> 4);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            code_next[i] = (nrow < p.rows_a && cbeg + 16 * i < cend && ncol0 + cbeg + 16 * i < p.cols_b)
                               ? __ldg(cp + i) : 0u;
        }
      }
      ptx::mbar_wait(tfull0 + 8 * acc, aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      // The warp's span is up to four 16-column chunks.  The TMEM load of chunk k+1 is in flight while chunk k goes
      // through the element-wise epilogue (two register sets, alternating; tcgen05.wait::ld before the set is read):
      // with load -> wait -> compute about two of the four warps of a scheduler sat on the TMEM load at any time
      // (ncu: long scoreboard 1.9 per issue, issue slots 50 % busy on the discriminator's forward launches).
      uint32_t code_out[4] = {0u, 0u, 0u, 0u};
      const int nck = (p.dbg & 8) ? 0 : (cend - cbeg) >> 4;
      // (a macro, not a lambda: k must stay a literal so that codes[] / code_out[] live in registers)
#define run_chunk(r, k)                                                                                          \
  do {                                                                                                           \
    const int c = cbeg + 16 * (k);                                                                               \
    if (EPI == EPI_F32 && stage_scr != nullptr) { /* whole warp takes part; rows beyond rows_a masked at the store */ \
      if (col0 + c < p.cols_b) epilogue_f32_staged(p, r, row - lane, lane, col0 + c, z, bias_s, stage_scr);      \
    } else if (row_ok && col0 + c < p.cols_b) {                                                                  \
      code_out[k] = epilogue_chunk16<EPI>(p, r, row, col0 + c, z, bias_s, codes[k]);                             \
    }                                                                                                            \
  } while (0)
      if (EPI == EPI_F32) {
        // fp32 outputs: one chunk at a time (the staged store path needs the registers a second set would take:
        // 196 bytes of spills under the 96-register cap of a 576-thread block)
        uint32_t ra[16];
        if (nck > 0) { ptx::tmem_ld16(taddr + cbeg, ra); ptx::tmem_ld_wait(); run_chunk(ra, 0); }
        if (nck > 1) { ptx::tmem_ld16(taddr + cbeg + 16, ra); ptx::tmem_ld_wait(); run_chunk(ra, 1); }
        if (nck > 2) { ptx::tmem_ld16(taddr + cbeg + 32, ra); ptx::tmem_ld_wait(); run_chunk(ra, 2); }
        if (nck > 3) { ptx::tmem_ld16(taddr + cbeg + 48, ra); ptx::tmem_ld_wait(); run_chunk(ra, 3); }
      } else {
        uint32_t ra[16], rb[16];
        if (nck > 0) {
          ptx::tmem_ld16(taddr + cbeg, ra);
          ptx::tmem_ld_wait();
          if (nck > 1) ptx::tmem_ld16(taddr + cbeg + 16, rb);
          run_chunk(ra, 0);
        }
        if (nck > 1) {
          ptx::tmem_ld_wait();
          if (nck > 2) ptx::tmem_ld16(taddr + cbeg + 32, ra);
          run_chunk(rb, 1);
        }
        if (nck > 2) {
          ptx::tmem_ld_wait();
          if (nck > 3) ptx::tmem_ld16(taddr + cbeg + 48, rb);
          run_chunk(ra, 2);
        }
        if (nck > 3) {
          ptx::tmem_ld_wait();
          run_chunk(rb, 3);
        }
      }
#undef run_chunk
      if (EPI == EPI_PLANES_FWD && p.code != nullptr && row_ok) {
        uint32_t* cp = p.code + row * p.code_pitch + ((col0 + cbeg) >> 4);
        if (cw == 64 && (p.code_pitch & 3) == 0) {
          // whole 16-byte group inside the (4-word padded) row: one vector store
          *reinterpret_cast<uint4*>(cp) = make_uint4(code_out[0], code_out[1], code_out[2], code_out[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (cbeg + 16 * i < cend && col0 + cbeg + 16 * i < p.cols_b) cp[i] = code_out[i];
        }
      }
      if (MN && p.db != nullptr && tb == 0 && chalf == 0) {
        uint32_t r0[16];
        ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 256, r0);
        ptx::tmem_ld_wait();
        if (row_ok) p.db[(int64_t)z * p.rows_a + row] = __uint_as_float(r0[0]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty0 + 8 * acc);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync();                 // no CTA leaves while its peer may still signal / multicast into it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------- CTA-pair kernel
// K-major GEMM on CTA pairs (cluster of 2, tcgen05 cta_group::2): one MMA of M = 256 spans both SMs.  Each
// CTA loads its own 128 rows of A and only HALF of the B tile, so the per-SM operand ingest per unit of MMA
// work drops by a third against the single-CTA kernel (the measured limiter, profiles/r01_gemm_experiments.md)
// and the smem stage shrinks to 64 KB (3 stages).  Leader = cluster rank 0: it alone issues the MMAs; its
// `full` barriers collect the TMA bytes of BOTH CTAs; commits are multicast to both CTAs' barriers; the
// peer's epilogue warps release the accumulator on the leader's barrier with a remote arrive.
// BRES = true: the WHOLE B operand of the column tile (N <= 256 rows, K <= 256: the discriminator's weight matrices,
// 128 KB of hi/lo planes per CTA of the pair) is loaded ONCE and stays in shared memory for every row tile of the
// persistent CTA; only A streams (32-wide K stages, SWIZZLE_64B): the per-SM operand ingest per unit of MMA work drops
// from (128 + N/2) to 128 rows, which is what bounds the streaming kernels (profiles/r01_gemm_experiments.md).
template <int EPI, bool BRES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                 const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t ring0 = base + (BRES ? p.bres_bytes : 0u);          // resident B blocks first, then the stage ring
  const uint32_t bar_base = ring0 + p.num_stages * p.stage_bytes;
  const uint32_t full0 = bar_base, empty0 = bar_base + 8 * TC_MAX_STAGES;
  const uint32_t tfull0 = bar_base + 16 * TC_MAX_STAGES, tempty0 = tfull0 + 16;
  const uint32_t tmem_slot = tempty0 + 16;
  const uint32_t bfull = tmem_slot + 8;                              // BRES: the resident B operand has landed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull0 + 8 * a, 1);
      ptx::mbar_init(tempty0 + 8 * a, 2 * TC_EPI_WARPS);      // both CTAs' epilogue warps (leader's copy is used)
    }
    if (BRES) ptx::mbar_init(bfull, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmAh);
    ptx::prefetch_tensormap(&tmAl);
    ptx::prefetch_tensormap(&tmBh);
    ptx::prefetch_tensormap(&tmBl);
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish_pair();
  }
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const float* bias_s = nullptr;
  if (p.bias_off) {
    float* bs = reinterpret_cast<float*>(smem_raw + (base + p.bias_off - raw));
    const int nb = p.num_b * p.bn;
    for (int i = threadIdx.x; i < nb; i += TC_THREADS) bs[i] = i < p.cols_b ? p.bias[i] : 0.f;
    bias_s = bs;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int num_pairs = (p.num_a + 1) / 2;
  const int total_tiles = num_pairs * p.num_b;
  const int first_tile = (int)blockIdx.x / 2, tile_step = (int)gridDim.x / 2;
  const int half_bn = p.bn / 2;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      uint32_t s = 0, ph = 0;
      if (BRES) {
        // the CTA's half of the B rows, every 64-wide K block, once
        const int nkb = (int)((p.red + 63) / 64);
        if (leader) ptx::mbar_expect_tx(bfull, 2u * (uint32_t)nkb * 2u * p.b_plane_bytes);
        const uint32_t bb = ptx::mapa(bfull, 0);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::tma_load_2d_pair(base + kb * 2 * p.b_plane_bytes, &tmBh, bb, kb * 64, (int)crank * half_bn);
          ptx::tma_load_2d_pair(base + kb * 2 * p.b_plane_bytes + p.b_plane_bytes, &tmBl, bb, kb * 64, (int)crank * half_bn);
        }
      }
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int ta = (tile / p.num_b) * 2 + (int)crank, tb = tile % p.num_b;
        const int a0 = ta * TC_BM, b0 = tb * p.bn + (int)crank * half_bn;
        for (int64_t r0 = 0; r0 < p.red; r0 += p.bk) {
          ptx::mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t fb_local = full0 + 8 * s;
          if (leader) ptx::mbar_expect_tx(fb_local, 2 * p.tx_bytes);   // bytes of both CTAs land on this barrier
          const uint32_t fb = ptx::mapa(fb_local, 0);
          const uint32_t sa_hi = ring0 + s * p.stage_bytes, sa_lo = sa_hi + p.a_plane;
          ptx::tma_load_2d_pair(sa_hi, &tmAh, fb, (int32_t)r0, a0);
          ptx::tma_load_2d_pair(sa_lo, &tmAl, fb, (int32_t)r0, a0);
          if (!BRES) {
            const uint32_t sb_hi = sa_lo + p.a_plane, sb_lo = sb_hi + p.b_plane_bytes;
            ptx::tma_load_2d_pair(sb_hi, &tmBh, fb, (int32_t)r0, b0);
            ptx::tma_load_2d_pair(sb_lo, &tmBl, fb, (int32_t)r0, b0);
          }
          if (++s == (uint32_t)p.num_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ------------------------------------------------------------ MMA issuer (leader CTA, one thread)
      const uint32_t idesc = ptx::make_idesc_bf16(2 * TC_BM, p.bn, 0, 0);
      uint32_t s = 0, ph = 0;
      int it = 0;
      if (BRES) ptx::mbar_wait(bfull, 0);
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        ptx::mbar_wait(tempty0 + 8 * acc, aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t first = 0;
        for (int64_t r0 = 0; r0 < p.red; r0 += p.bk) {
          ptx::mbar_wait(full0 + 8 * s, ph);
          ptx::tc_fence_after();
          const uint32_t sa_hi = ring0 + s * p.stage_bytes, sa_lo = sa_hi + p.a_plane;
          if (BRES) {
            // A: 32-wide K stage (SWIZZLE_64B); B: resident 64-wide K block kb (SWIZZLE_128B), half (r0/32)&1 of it
            const uint32_t kb = (uint32_t)(r0 >> 6), hb = (uint32_t)((r0 >> 5) & 1);
            const uint32_t sb_hi = base + kb * 2 * p.b_plane_bytes + hb * 64u, sb_lo = sb_hi + p.b_plane_bytes;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t da_hi = ptx::make_smem_desc(sa_hi + k * 32u, 0u, 512u, 4u);
              const uint64_t da_lo = ptx::make_smem_desc(sa_lo + k * 32u, 0u, 512u, 4u);
              const uint64_t db_hi = ptx::make_smem_desc(sb_hi + k * 32u, 0u, 1024u, 2u);
              const uint64_t db_lo = ptx::make_smem_desc(sb_lo + k * 32u, 0u, 1024u, 2u);
              ptx::mma_bf16_ss_pair(d_tmem, da_hi, db_hi, idesc, first);
              first = 1;
              ptx::mma_bf16_ss_pair(d_tmem, da_hi, db_lo, idesc, 1);
              ptx::mma_bf16_ss_pair(d_tmem, da_lo, db_hi, idesc, 1);
            }
          } else {
            const uint32_t sb_hi = sa_lo + p.a_plane, sb_lo = sb_hi + p.b_plane_bytes;
            for (int k = 0; k < p.bk / 16; ++k) {
              const uint64_t da_hi = ptx::make_smem_desc(sa_hi + k * 32u, 0u, p.sbo, p.desc_layout);
              const uint64_t da_lo = ptx::make_smem_desc(sa_lo + k * 32u, 0u, p.sbo, p.desc_layout);
              const uint64_t db_hi = ptx::make_smem_desc(sb_hi + k * 32u, 0u, p.sbo, p.desc_layout);
              const uint64_t db_lo = ptx::make_smem_desc(sb_lo + k * 32u, 0u, p.sbo, p.desc_layout);
              ptx::mma_bf16_ss_pair(d_tmem, da_hi, db_hi, idesc, first);
              first = 1;
              ptx::mma_bf16_ss_pair(d_tmem, da_hi, db_lo, idesc, 1);
              ptx::mma_bf16_ss_pair(d_tmem, da_lo, db_hi, idesc, 1);
            }
          }
          ptx::mma_commit_pair(empty0 + 8 * s, (uint16_t)0x3);     // frees the stage in both CTAs
          if (++s == (uint32_t)p.num_stages) { s = 0; ph ^= 1; }
        }
        ptx::mma_commit_pair(tfull0 + 8 * acc, (uint16_t)0x3);      // accumulators ready in both CTAs
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue warps (both CTAs, own 128 rows)
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int cw = p.bn >> 2;
    const int cbeg = chalf * cw, cend = cbeg + cw;
    uint32_t code_next[4] = {0u, 0u, 0u, 0u};
    if (EPI == EPI_PLANES_BWD && first_tile < total_tiles) {
      const int64_t nrow = (int64_t)((first_tile / p.num_b) * 2 + (int)crank) * TC_BM + q * 32 + lane;
      const int ncol0 = (first_tile % p.num_b) * p.bn;
      const uint32_t* cp = p.code + nrow * p.code_pitch + ((ncol0 + cbeg) >> 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        code_next[i] = (nrow < p.rows_a && cbeg + 16 * i < cend && ncol0 + cbeg + 16 * i < p.cols_b)
                           ? __ldg(cp + i) : 0u;
    }
    int it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
      const int ta = (tile / p.num_b) * 2 + (int)crank, tb = tile % p.num_b;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int64_t row = (int64_t)ta * TC_BM + q * 32 + lane;
      const int col0 = tb * p.bn;
      const bool row_ok = row < p.rows_a;
      uint32_t codes[4];
      if (EPI == EPI_PLANES_BWD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) codes[i] = code_next[i];
        const int ntile = tile + tile_step;
        if (ntile < total_tiles) {
          const int64_t nrow = (int64_t)((ntile / p.num_b) * 2 + (int)crank) * TC_BM + q * 32 + lane;
          const int ncol0 = (ntile % p.num_b) * p.bn;
          const uint32_t* cp = p.code + nrow * p.code_pitch + ((ncol0 + cbeg) >> 4);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            code_next[i] = (nrow < p.rows_a && cbeg + 16 * i < cend && ncol0 + cbeg + 16 * i < p.cols_b)
                               ? __ldg(cp + i) : 0u;
        }
      }
      ptx::mbar_wait(tfull0 + 8 * acc, aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      // TMEM load of chunk k+1 in flight under the epilogue of chunk k (see gemm_bf16x3_kernel)
      uint32_t code_out[4] = {0u, 0u, 0u, 0u};
      const int nck = (cend - cbeg) >> 4;
#define run_chunk(r, k)                                                                                          \
  do {                                                                                                           \
    const int c = cbeg + 16 * (k);                                                                               \
    if (row_ok && col0 + c < p.cols_b) code_out[k] = epilogue_chunk16<EPI>(p, r, row, col0 + c, 0, bias_s, codes[k]); \
  } while (0)
      if (EPI == EPI_F32) {
        // fp32 outputs: one chunk at a time (the staged store path needs the registers a second set would take:
        // 196 bytes of spills under the 96-register cap of a 576-thread block)
        uint32_t ra[16];
        if (nck > 0) { ptx::tmem_ld16(taddr + cbeg, ra); ptx::tmem_ld_wait(); run_chunk(ra, 0); }
        if (nck > 1) { ptx::tmem_ld16(taddr + cbeg + 16, ra); ptx::tmem_ld_wait(); run_chunk(ra, 1); }
        if (nck > 2) { ptx::tmem_ld16(taddr + cbeg + 32, ra); ptx::tmem_ld_wait(); run_chunk(ra, 2); }
        if (nck > 3) { ptx::tmem_ld16(taddr + cbeg + 48, ra); ptx::tmem_ld_wait(); run_chunk(ra, 3); }
      } else {
        uint32_t ra[16], rb[16];
        if (nck > 0) {
          ptx::tmem_ld16(taddr + cbeg, ra);
          ptx::tmem_ld_wait();
          if (nck > 1) ptx::tmem_ld16(taddr + cbeg + 16, rb);
          run_chunk(ra, 0);
        }
        if (nck > 1) {
          ptx::tmem_ld_wait();
          if (nck > 2) ptx::tmem_ld16(taddr + cbeg + 32, ra);
          run_chunk(rb, 1);
        }
        if (nck > 2) {
          ptx::tmem_ld_wait();
          if (nck > 3) ptx::tmem_ld16(taddr + cbeg + 48, rb);
          run_chunk(ra, 2);
        }
        if (nck > 3) {
          ptx::tmem_ld_wait();
          run_chunk(rb, 3);
        }
      }
#undef run_chunk
      if (EPI == EPI_PLANES_FWD && p.code != nullptr && row_ok) {
        uint32_t* cp = p.code + row * p.code_pitch + ((col0 + cbeg) >> 4);
        if (cw == 64 && (p.code_pitch & 3) == 0) {
          *reinterpret_cast<uint4*>(cp) = make_uint4(code_out[0], code_out[1], code_out[2], code_out[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (cbeg + 16 * i < cend && col0 + cbeg + 16 * i < p.cols_b) cp[i] = code_out[i];
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // the leader's MMA issuer owns the accumulator hand-back of BOTH CTAs
        if (leader) ptx::mbar_arrive(tempty0 + 8 * acc);
        else ptx::mbar_arrive_cluster(ptx::mapa(tempty0 + 8 * acc, 0));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, p.tmem_cols);
  }
}


// ---------------------------------------------------------------------------- CTA-pair kernel, MN-major
// Weight-gradient GEMM C[n][k] = sum_m A[m][n] B[m][k] on CTA pairs: the pair's MMA is 256 (n) x bn (k), each
// CTA stages its own 128 columns of A and HALF of the B columns per reduction block, so the per-SM operand
// ingest per unit of MMA work drops from (128 + bn) to (128 + bn/2) rows -- the single-CTA MN kernel is
// ingest-bound (profiles/r01_gemm_experiments.md).  Requires bn % 128 == 0 (each half is whole 64-wide atoms).
// One accumulator at TMEM columns [0, bn); the bias gradient (column sums of A) accumulates at [256, 272)
// through an extra N = 16 MMA against an all-ones K-major tile present in both CTAs' shared memory.
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_pair_mn_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                    const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                    const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t ones_base = base + p.ones_off;
  const uint32_t bar_base = base + p.num_stages * p.stage_bytes + 8192u;
  const uint32_t full0 = bar_base, empty0 = bar_base + 8 * TC_MAX_STAGES;
  const uint32_t tfull0 = bar_base + 16 * TC_MAX_STAGES, tempty0 = tfull0 + 16;
  const uint32_t tmem_slot = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, 1);
    }
    ptx::mbar_init(tfull0, 1);
    ptx::mbar_init(tempty0, 2 * TC_EPI_WARPS);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmAh);
    ptx::prefetch_tensormap(&tmAl);
    ptx::prefetch_tensormap(&tmBh);
    ptx::prefetch_tensormap(&tmBl);
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish_pair();
  }
  {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_raw + (ones_base - raw));
    for (int i = threadIdx.x; i < 8192 / 4; i += TC_THREADS) ones[i] = 0x3F803F80u;
    ptx::fence_proxy_async();
  }
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int num_pairs = (p.num_a + 1) / 2;
  const int tiles_ab = num_pairs * p.num_b;
  const int total_tiles = tiles_ab * p.num_z;
  const int first_tile = (int)blockIdx.x / 2, tile_step = (int)gridDim.x / 2;
  const int half_bn = p.bn / 2;
  const int half_atoms = half_bn / 64;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      uint32_t s = 0, ph = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int z = tile / tiles_ab, rem = tile - z * tiles_ab;
        const int ta = (rem / p.num_b) * 2 + (int)crank, tb = rem % p.num_b;
        const int64_t r_beg = (int64_t)z * p.red_chunk;
        const int64_t r_end = r_beg + p.red_chunk < p.red ? r_beg + p.red_chunk : p.red;
        const int a0 = ta * TC_BM, b0 = tb * p.bn + (int)crank * half_bn;
        for (int64_t r0 = r_beg; r0 < r_end; r0 += p.bk) {
          ptx::mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t fb_local = full0 + 8 * s;
          if (leader) ptx::mbar_expect_tx(fb_local, 2 * p.tx_bytes);
          const uint32_t fb = ptx::mapa(fb_local, 0);
          const uint32_t sa_hi = base + s * p.stage_bytes, sa_lo = sa_hi + p.a_plane;
          const uint32_t sb_hi = sa_lo + p.a_plane, sb_lo = sb_hi + p.b_plane_bytes;
          for (int j = 0; j < TC_BM / 64; ++j) {
            ptx::tma_load_2d_pair(sa_hi + j * p.atom_bytes, &tmAh, fb, a0 + 64 * j, (int32_t)r0);
            ptx::tma_load_2d_pair(sa_lo + j * p.atom_bytes, &tmAl, fb, a0 + 64 * j, (int32_t)r0);
          }
          for (int j = 0; j < half_atoms; ++j) {
            ptx::tma_load_2d_pair(sb_hi + j * p.atom_bytes, &tmBh, fb, b0 + 64 * j, (int32_t)r0);
            ptx::tma_load_2d_pair(sb_lo + j * p.atom_bytes, &tmBl, fb, b0 + 64 * j, (int32_t)r0);
          }
          if (++s == (uint32_t)p.num_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ------------------------------------------------------------ MMA issuer (leader CTA, one thread)
      const uint32_t idesc = ptx::make_idesc_bf16(2 * TC_BM, p.bn, 1, 1);
      const uint32_t idesc_db = ptx::make_idesc_bf16(2 * TC_BM, 16, 1, 0);   // ones tile read K-major
      const uint64_t d_ones = ptx::make_smem_desc(ones_base, 0u, 1024);
      const uint32_t lbo = p.atom_bytes;
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
        const int z = tile / tiles_ab, rem = tile - z * tiles_ab;
        const int64_t r_beg = (int64_t)z * p.red_chunk;
        const int64_t r_end = r_beg + p.red_chunk < p.red ? r_beg + p.red_chunk : p.red;
        const bool do_db = p.db != nullptr && (rem % p.num_b) == 0;
        ptx::mbar_wait(tempty0, (uint32_t)(it & 1) ^ 1);
        ptx::tc_fence_after();
        uint32_t first = 0;
        for (int64_t r0 = r_beg; r0 < r_end; r0 += p.bk) {
          ptx::mbar_wait(full0 + 8 * s, ph);
          ptx::tc_fence_after();
          const uint32_t sa_hi = base + s * p.stage_bytes, sa_lo = sa_hi + p.a_plane;
          const uint32_t sb_hi = sa_lo + p.a_plane, sb_lo = sb_hi + p.b_plane_bytes;
          for (int k = 0; k < p.bk / 16; ++k) {
            const uint64_t da_hi = ptx::make_smem_desc(sa_hi + k * 2048u, lbo, 1024);
            const uint64_t da_lo = ptx::make_smem_desc(sa_lo + k * 2048u, lbo, 1024);
            const uint64_t db_hi = ptx::make_smem_desc(sb_hi + k * 2048u, lbo, 1024);
            const uint64_t db_lo = ptx::make_smem_desc(sb_lo + k * 2048u, lbo, 1024);
            ptx::mma_bf16_ss_pair(tmem_base, da_hi, db_hi, idesc, first);
            if (do_db) {
              ptx::mma_bf16_ss_pair(tmem_base + 256, da_hi, d_ones, idesc_db, first);
              ptx::mma_bf16_ss_pair(tmem_base + 256, da_lo, d_ones, idesc_db, 1);
            }
            first = 1;
            ptx::mma_bf16_ss_pair(tmem_base, da_hi, db_lo, idesc, 1);
            ptx::mma_bf16_ss_pair(tmem_base, da_lo, db_hi, idesc, 1);
          }
          ptx::mma_commit_pair(empty0 + 8 * s, (uint16_t)0x3);
          if (++s == (uint32_t)p.num_stages) { s = 0; ph ^= 1; }
        }
        ptx::mma_commit_pair(tfull0, (uint16_t)0x3);
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue warps (both CTAs, own 128 rows)
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int cw = p.bn >> 2;
    const int cbeg = chalf * cw, cend = cbeg + cw;
    int it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
      const int z = tile / tiles_ab, rem = tile - z * tiles_ab;
      const int ta = (rem / p.num_b) * 2 + (int)crank, tb = rem % p.num_b;
      const int64_t row = (int64_t)ta * TC_BM + q * 32 + lane;
      const int col0 = tb * p.bn;
      const bool row_ok = row < p.rows_a;
      ptx::mbar_wait(tfull0, (uint32_t)(it & 1));
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c = cbeg + 32 * ci;
        if (c < cend) {
          uint32_t r0[16], r1[16];
          const bool two = c + 32 <= cend;
          ptx::tmem_ld16(taddr + c, r0);
          if (two) ptx::tmem_ld16(taddr + c + 16, r1);
          ptx::tmem_ld_wait();
          if (row_ok) {
            if (col0 + c < p.cols_b) epilogue_chunk16<EPI_F32>(p, r0, row, col0 + c, z, nullptr, 0u);
            if (two && col0 + c + 16 < p.cols_b) epilogue_chunk16<EPI_F32>(p, r1, row, col0 + c + 16, z, nullptr, 0u);
          }
        }
      }
      if (p.db != nullptr && tb == 0 && chalf == 0) {
        uint32_t r0[16];
        ptx::tmem_ld16(taddr + 256, r0);
        ptx::tmem_ld_wait();
        if (row_ok) p.db[(int64_t)z * p.rows_a + row] = __uint_as_float(r0[0]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(tempty0);
        else ptx::mbar_arrive_cluster(ptx::mapa(tempty0, 0));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------- operand planes
// fp32 [rows][cols] (row stride rs) -> bf16 hi/lo planes [rows][pitch]; transpose: out[c][r] = in[r][c].
__global__ void split_planes_kernel(const float* __restrict__ src, int64_t rs, int64_t rows, int cols,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    int64_t pitch, int transpose) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (!transpose) {
    // one warp per row, each lane converts PAIRS of columns -> 4-byte stores, 128 B per warp store
    for (int64_t r = warp; r < rows; r += nwarps) {
      const float* sr = src + r * rs;
      uint32_t* hr = reinterpret_cast<uint32_t*>(hi + r * pitch);
      uint32_t* lr = reinterpret_cast<uint32_t*>(lo + r * pitch);
#pragma unroll 2
      for (int c = 2 * lane; c < cols; c += 64) {
        const float a = sr[c], b = (c + 1 < cols) ? sr[c + 1] : 0.f;
        const uint32_t hp = pack_bf16x2(a, b);
        const float ah = __uint_as_float(hp << 16), bh = __uint_as_float(hp & 0xffff0000u);
        hr[c >> 1] = hp;
        lr[c >> 1] = pack_bf16x2(a - ah, b - bh);
      }
    }
  } else {
    for (int64_t r = warp; r < rows; r += nwarps)
      for (int c = lane; c < cols; c += 32) {
        const float v = src[r * rs + c];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[(int64_t)c * pitch + r] = h;
        lo[(int64_t)c * pitch + r] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
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

// 2D bf16 tensor [rows][cols] with row pitch `pitch` elements; box = {box_cols (64: SWIZZLE_128B, 32: SWIZZLE_64B),
// box_rows}.
static int make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t pitch, int box_rows,
                    int box_cols = 64) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("tc: cuTensorMapEncodeTiled entry point unavailable");
    return GANTTS_E_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld pitch=%lld box_rows=%d", (int)r,
              (long long)rows, (long long)cols, (long long)pitch, box_rows);
    return GANTTS_E_CUDA;
  }
  return GANTTS_OK;
}

static int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}

static int num_sms() {
  static int n[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return 148;
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

static inline int64_t pitch_for(int64_t cols) { return (cols + 15) / 16 * 16; }   // 32-byte rows
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
  GANTTS_PDL_LAUNCH((split_planes_kernel), nb, 256, 0, st, src, rs, rows, cols, pl.hi, pl.lo, pl.pitch, transpose);
  GANTTS_LAUNCH_CHECK("split_planes_kernel");
  return GANTTS_OK;
}

static int pick_bn(int n, int cap_override = 0) {
  static int cap_env = -1;
  if (cap_env < 0) {
    const char* e = getenv("GANTTS_B200_BN");
    cap_env = e ? atoi(e) : 256;
    if (cap_env < 64 || cap_env > 256 || cap_env % 64) cap_env = 256;
  }
  const int cap = cap_override ? cap_override : cap_env;
  int bn = (n + 63) / 64 * 64;
  if (bn <= cap) return bn;
  int tiles = (n + cap - 1) / cap;
  bn = ((n + tiles - 1) / tiles + 63) / 64 * 64;
  return bn;
}

// Epilogue description filled by the callers of launch_gemm_kk.
struct EpiArgs {
  int epi = EPI_F32;
  // EPI_F32
  float* C = nullptr;
  int64_t ldc = 0;
  int accumulate = 0;
  // EPI_PLANES_*
  __nv_bfloat16 *out_hi = nullptr, *out_lo = nullptr;
  int64_t out_pitch = 0;
  uint32_t* code = nullptr;           // derivative code plane (written by PLANES_FWD, read by PLANES_BWD)
  int64_t code_pitch = 0;
  // math
  const float* bias = nullptr;
  int act = GANTTS_ACT_NONE;
  float slope = 0.f, p = 0.f;
  uint64_t seed = 0;
  int64_t row0 = 0;                   // global index of A's first row (row windows of a larger matrix)
};

static void fill_epilogue(GemmParams& p, const EpiArgs& e) {
  p.C = e.C;
  p.ldc = e.ldc;
  p.accumulate = e.accumulate;
  p.vec_ok = (!e.accumulate && e.C && (e.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(e.C) & 15) == 0) ? 1 : 0;
  if (p.vec_ok && (e.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(e.C) & 31) == 0) p.vec_ok = 2;   // STG.256
  p.out_hi = e.out_hi;
  p.out_lo = e.out_lo;
  p.out_pitch = e.out_pitch;
  p.code = e.code;
  p.code_pitch = e.code_pitch;
  p.bias = e.bias;
  p.act = e.act;
  p.slope = e.slope;
  p.keep_scale = e.p > 0.f ? 1.f / (1.f - e.p) : 1.f;
  p.thresh = e.p > 0.f ? (uint32_t)(e.p * 65536.f + 0.5f) : 0u;
  p.seed = e.seed;
  p.row0 = e.row0;
  static int dbg = -1;
  if (dbg < 0) {
    const char* v = getenv("GANTTS_B200_DBG");
    dbg = v ? atoi(v) : 0;
    if (dbg)
      fprintf(stderr, "gantts_b200: GANTTS_B200_DBG=%d removes parts of the GEMM kernels (phase timing only): "
                      "RESULTS ARE WRONG, do not use for training or benchmarks\n", dbg);
  }
  p.dbg = (uint32_t)dbg;
}

// Reduction elements per smem stage.  Measured on cfg2 (profiles/r01_gemm_experiments.md): the K-major kernels
// are ~4% faster with 64 (two or three 128-byte-swizzled stages), the MN-major weight-gradient kernel ~7% faster
// with 32 (twice as many, half as large stages).  GANTTS_B200_BK=32|64 forces one value for both.
static int stage_bk(bool mn) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GANTTS_B200_BK");
    v = e ? atoi(e) : 0;
    if (v != 32 && v != 64) v = 0;
  }
  return v ? v : (mn ? 32 : 64);
}

// Stage fp32 output tiles with unaligned row strides through shared memory (GANTTS_B200_F32_STAGE=0 disables).
static int use_f32_stage() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GANTTS_B200_F32_STAGE");
    v = e ? atoi(e) : 1;
  }
  return v;
}

// B-resident pair kernel for narrow layers: OPT-IN (GANTTS_B200_BRES=1).  Measured on B200 at cfg2
// (profiles/r02_gemm_experiments.md): correct, but the discriminator's 256-wide launches did not get faster (34.6 us
// against 33.5 us per forward launch, step 1.298 against 1.279 ms) -- they are bound by the epilogue and its tape
// writes, not by the operand ingest the resident weights remove.  Read per call so tests can switch it.
static int use_bres() {
  const char* e = getenv("GANTTS_B200_BRES");
  return e ? atoi(e) : 0;
}

static int use_pdl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GANTTS_B200_PDL");
    v = e ? atoi(e) : 1;
  }
  return v;
}

template <bool MN, int EPI, int CL>
static int launch_kernel(const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh,
                         const CUtensorMap& mBl, const GemmParams& p, cudaStream_t st) {
  const size_t smem = (size_t)p.num_stages * p.stage_bytes + (MN ? 8192 : 0) + 1024 + 256 + TC_BIAS_SMEM +
                      (p.f32_stage_off ? TC_F32_STAGE_SMEM : 0);
  // the opt-in is per device (and context): remember it per device ordinal, not per process
  static bool attr[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    GANTTS_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel<MN, EPI, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  const int units = ((p.num_a + CL - 1) / CL) * p.num_b * p.num_z;     // tiles (CL=1) or pair tiles (CL=2)
  int grid = units * CL < num_sms() ? units * CL : num_sms() / CL * CL;
  prof_begin(MN ? PROF_GEMM_MN : PROF_GEMM_KK, 2.0 * (double)p.rows_a * p.cols_b * (double)p.red, st);
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (use_pdl()) {
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    if (CL > 1) {
      at[na].id = cudaLaunchAttributeClusterDimension;
      at[na].val.clusterDim.x = CL;
      at[na].val.clusterDim.y = 1;
      at[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16x3_kernel<MN, EPI, CL>, mAh, mAl, mBh, mBl, p);
    if (e != cudaSuccess) {
      prof_end(st);
      return cuda_fail(e, "cudaLaunchKernelEx(gemm)");
    }
  }
  prof_end(st);
  GANTTS_LAUNCH_CHECK("gemm_bf16x3_kernel");
  return GANTTS_OK;
}

template <int EPI, bool BRES = false>
static int launch_pair_kernel(const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh,
                              const CUtensorMap& mBl, const GemmParams& p, cudaStream_t st) {
  const size_t smem = (size_t)(BRES ? p.bres_bytes : 0u) + (size_t)p.num_stages * p.stage_bytes + 1024 + 256 + TC_BIAS_SMEM;
  static bool attr[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    GANTTS_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<EPI, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  const int units = ((p.num_a + 1) / 2) * p.num_b;
  const int grid = units * 2 < num_sms() ? units * 2 : num_sms() / 2 * 2;
  prof_begin(PROF_GEMM_KK, 2.0 * (double)p.rows_a * p.cols_b * (double)p.red, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  at[na].id = cudaLaunchAttributeClusterDimension;
  at[na].val.clusterDim.x = 2;
  at[na].val.clusterDim.y = 1;
  at[na].val.clusterDim.z = 1;
  ++na;
  if (use_pdl()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_pair_kernel<EPI, BRES>, mAh, mAl, mBh, mBl, p);
  prof_end(st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm pair)");
  GANTTS_LAUNCH_CHECK("gemm_pair_kernel");
  return GANTTS_OK;
}

static int launch_pair_mn_kernel(const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh,
                                 const CUtensorMap& mBl, const GemmParams& p, cudaStream_t st) {
  const size_t smem = (size_t)p.num_stages * p.stage_bytes + 8192 + 1024 + 256;
  static bool attr[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    GANTTS_CUDA(cudaFuncSetAttribute(gemm_pair_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  const int units = ((p.num_a + 1) / 2) * p.num_b * p.num_z;
  const int grid = units * 2 < num_sms() ? units * 2 : num_sms() / 2 * 2;
  prof_begin(PROF_GEMM_MN, 2.0 * (double)p.rows_a * p.cols_b * (double)p.red, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  at[na].id = cudaLaunchAttributeClusterDimension;
  at[na].val.clusterDim.x = 2;
  at[na].val.clusterDim.y = 1;
  at[na].val.clusterDim.z = 1;
  ++na;
  if (use_pdl()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_pair_mn_kernel, mAh, mAl, mBh, mBl, p);
  prof_end(st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm pair mn)");
  GANTTS_LAUNCH_CHECK("gemm_pair_mn_kernel");
  return GANTTS_OK;
}

static int use_cluster() {
  static int v = -1;
  if (v < 0) {
    // 0 (default) = auto: CTA-pair MMA (cta_group::2, mode 3) for wide outputs (>= 2 column tiles), single
    // CTA otherwise -- measured on B200 (profiles/r01_gemm_experiments.md): pairs are ~7 % faster on the
    // 512-wide generator layers and ~5 % slower on the 256-wide discriminator layers.  1 = always single
    // CTA, 2 = 2-CTA cluster with B-tile multicast (no gain: multicast does not reduce the per-SM ingest),
    // 3 = always CTA pairs.
    const char* e = getenv("GANTTS_B200_CLUSTER");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// Staged fp32 epilogue (single-CTA kernels): reserve the 32 KB transpose scratch behind the bias block and give the
// stages what is left.  No-op when the output row stride is aligned (vector stores) or GANTTS_B200_F32_STAGE=0.
static void maybe_stage_f32(GemmParams& p, const EpiArgs& e, bool mn, uint32_t budget_bytes) {
  p.f32_stage_off = 0;
  if (!use_f32_stage() || e.epi != EPI_F32 || p.vec_ok || !p.C) return;
  const int stages = (int)((budget_bytes - TC_F32_STAGE_SMEM) / p.stage_bytes);
  if (stages < 2) return;
  p.num_stages = stages < p.num_stages ? stages : p.num_stages;
  if (p.bias_off) p.bias_off = (uint32_t)p.num_stages * p.stage_bytes + 256u;
  if (mn) p.ones_off = (uint32_t)p.num_stages * p.stage_bytes;
  p.f32_stage_off = (uint32_t)p.num_stages * p.stage_bytes + (mn ? 8192u : 0u) + 256u + TC_BIAS_SMEM;
}

// out[rows_a][cols_b] = epi(A * B^T)   (K-major planes A [rows_a][red], B [cols_b][red]).
static int launch_gemm_kk_one(const Planes& A, const Planes& B, const EpiArgs& e, cudaStream_t st, int bn_cap = 0,
                              bool allow_pair = true) {
  GemmParams p{};
  p.rows_a = A.rows;
  p.cols_b = (int)B.rows;
  p.red = A.cols;
  if (B.cols != A.cols) {
    set_error("gemm_kk: reduction extents differ (%lld vs %lld)", (long long)A.cols, (long long)B.cols);
    return GANTTS_E_BADARG;
  }
  p.bk = stage_bk(false);
  const uint32_t rowb = (uint32_t)p.bk * 2;                 // bytes of one K-major smem row (128 or 64)
  p.a_plane = TC_BM * rowb;
  p.sbo = 8 * rowb;
  p.kstep = 32;
  p.desc_layout = p.bk == 64 ? 2u : 4u;
  p.red_chunk = (p.red + p.bk - 1) / p.bk * p.bk;
  p.bn = pick_bn(p.cols_b, bn_cap);
  p.num_a = (int)((p.rows_a + TC_BM - 1) / TC_BM);
  p.num_b = (p.cols_b + p.bn - 1) / p.bn;
  p.num_z = 1;
  p.b_plane_bytes = ((uint32_t)p.bn * rowb + 1023) / 1024 * 1024;
  p.stage_bytes = 2 * p.a_plane + 2 * p.b_plane_bytes;
  p.tx_bytes = 2 * p.a_plane + 2 * (uint32_t)p.bn * rowb;
  p.num_stages = (int)((216 * 1024) / p.stage_bytes);
  if (p.num_stages > TC_MAX_STAGES) p.num_stages = TC_MAX_STAGES;
  p.tmem_cols = 512;
  p.c_zstride = 0;
  fill_epilogue(p, e);
  // bias staged in smem after the barrier block when it fits (padded to whole column tiles)
  p.bias_off = (e.bias && (size_t)p.num_b * p.bn * sizeof(float) <= TC_BIAS_SMEM)
                   ? (uint32_t)p.num_stages * p.stage_bytes + 256u : 0u;
  // B-resident CTA pairs: one column tile (N <= 256) and K <= 256 -- the discriminator's layers
  if (allow_pair && use_bres() && p.num_b == 1 && p.red <= 256 && p.num_a >= 4 && p.bn >= 64 && (e.epi != EPI_F32 || p.vec_ok)) {
    const int nkb = (int)((p.red + 63) / 64);
    p.bk = 32;
    p.a_plane = TC_BM * 64u;                                       // 128 rows x 32 bf16, SWIZZLE_64B
    p.b_plane_bytes = (uint32_t)(p.bn / 2) * 128u;                 // (N/2) rows x 64 bf16, SWIZZLE_128B, per K block
    p.bres_bytes = (uint32_t)nkb * 2u * p.b_plane_bytes;
    p.stage_bytes = 2 * p.a_plane;
    p.tx_bytes = 2 * p.a_plane;
    p.num_stages = (int)((216 * 1024 - p.bres_bytes) / p.stage_bytes);
    if (p.num_stages > TC_MAX_STAGES) p.num_stages = TC_MAX_STAGES;
    p.bias_off = (e.bias && (size_t)p.num_b * p.bn * sizeof(float) <= TC_BIAS_SMEM)
                     ? p.bres_bytes + (uint32_t)p.num_stages * p.stage_bytes + 256u : 0u;
    CUtensorMap mAh, mAl, mBh, mBl;
    int rc;
    if ((rc = make_map(&mAh, A.hi, A.rows, A.cols, A.pitch, TC_BM, 32))) return rc;
    if ((rc = make_map(&mAl, A.lo, A.rows, A.cols, A.pitch, TC_BM, 32))) return rc;
    if ((rc = make_map(&mBh, B.hi, B.rows, B.cols, B.pitch, p.bn / 2, 64))) return rc;
    if ((rc = make_map(&mBl, B.lo, B.rows, B.cols, B.pitch, p.bn / 2, 64))) return rc;
    switch (e.epi) {
      case EPI_F32: return launch_pair_kernel<EPI_F32, true>(mAh, mAl, mBh, mBl, p, st);
      case EPI_PLANES_FWD: return launch_pair_kernel<EPI_PLANES_FWD, true>(mAh, mAl, mBh, mBl, p, st);
      case EPI_PLANES_BWD: return launch_pair_kernel<EPI_PLANES_BWD, true>(mAh, mAl, mBh, mBl, p, st);
    }
  }
  // GANTTS_B200_CLUSTER=3: CTA-pair MMA (cta_group::2), each CTA holds half of the B tile
  if (allow_pair && (use_cluster() == 3 || (use_cluster() == 0 && p.num_b >= 2)) && p.num_a >= 2) {
    p.b_plane_bytes = ((uint32_t)(p.bn / 2) * rowb + 1023) / 1024 * 1024;
    p.stage_bytes = 2 * p.a_plane + 2 * p.b_plane_bytes;
    p.tx_bytes = 2 * p.a_plane + 2 * (uint32_t)(p.bn / 2) * rowb;
    p.num_stages = (int)((216 * 1024) / p.stage_bytes);
    if (p.num_stages > TC_MAX_STAGES) p.num_stages = TC_MAX_STAGES;
    p.bias_off = (e.bias && (size_t)p.num_b * p.bn * sizeof(float) <= TC_BIAS_SMEM)
                     ? (uint32_t)p.num_stages * p.stage_bytes + 256u : 0u;
    CUtensorMap mAh, mAl, mBh, mBl;
    int rc;
    if ((rc = make_map(&mAh, A.hi, A.rows, A.cols, A.pitch, TC_BM, p.bk))) return rc;
    if ((rc = make_map(&mAl, A.lo, A.rows, A.cols, A.pitch, TC_BM, p.bk))) return rc;
    if ((rc = make_map(&mBh, B.hi, B.rows, B.cols, B.pitch, p.bn / 2, p.bk))) return rc;
    if ((rc = make_map(&mBl, B.lo, B.rows, B.cols, B.pitch, p.bn / 2, p.bk))) return rc;
    switch (e.epi) {
      case EPI_F32: return launch_pair_kernel<EPI_F32>(mAh, mAl, mBh, mBl, p, st);
      case EPI_PLANES_FWD: return launch_pair_kernel<EPI_PLANES_FWD>(mAh, mAl, mBh, mBl, p, st);
      case EPI_PLANES_BWD: return launch_pair_kernel<EPI_PLANES_BWD>(mAh, mAl, mBh, mBl, p, st);
    }
  }
  // 2-CTA clusters (B tile multicast) when there are at least two row tiles
  const bool cl2 = allow_pair && use_cluster() == 2 && p.num_a >= 2;
  if (!cl2) maybe_stage_f32(p, e, false, 216 * 1024);
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, A.hi, A.rows, A.cols, A.pitch, TC_BM, p.bk))) return rc;
  if ((rc = make_map(&mAl, A.lo, A.rows, A.cols, A.pitch, TC_BM, p.bk))) return rc;
  if ((rc = make_map(&mBh, B.hi, B.rows, B.cols, B.pitch, cl2 ? p.bn / 2 : p.bn, p.bk))) return rc;
  if ((rc = make_map(&mBl, B.lo, B.rows, B.cols, B.pitch, cl2 ? p.bn / 2 : p.bn, p.bk))) return rc;
  if (cl2) {
    switch (e.epi) {
      case EPI_F32: return launch_kernel<false, EPI_F32, 2>(mAh, mAl, mBh, mBl, p, st);
      case EPI_PLANES_FWD: return launch_kernel<false, EPI_PLANES_FWD, 2>(mAh, mAl, mBh, mBl, p, st);
      case EPI_PLANES_BWD: return launch_kernel<false, EPI_PLANES_BWD, 2>(mAh, mAl, mBh, mBl, p, st);
    }
  }
  switch (e.epi) {
    case EPI_F32: return launch_kernel<false, EPI_F32, 1>(mAh, mAl, mBh, mBl, p, st);
    case EPI_PLANES_FWD: return launch_kernel<false, EPI_PLANES_FWD, 1>(mAh, mAl, mBh, mBl, p, st);
    case EPI_PLANES_BWD: return launch_kernel<false, EPI_PLANES_BWD, 1>(mAh, mAl, mBh, mBl, p, st);
  }
  set_error("gemm_kk: bad epilogue %d", e.epi);
  return GANTTS_E_BADARG;
}

// Tail balancing, OPT-IN (GANTTS_B200_TAIL=1).  The persistent K-major kernels take whole rounds of tiles: 250 pair
// tiles on 74 CTA pairs are 3.38 rounds of work that cost 4 (the 512-wide generator layers at cfg2), 500 tiles on 148
// CTAs likewise.  The rows of the incomplete last round are cut off and given to a SECOND launch with narrower column
// tiles (BN/2 or BN/4, single CTAs) so that they spread over all SMs in half or three quarters of a round.
// Measured on B200 at cfg2 (profiles/r02_gemm_experiments.md): results bitwise identical, step 1.329 ms against
// 1.243 ms without -- the 16 extra launches (fill, drain, barrier/TMEM set-up) cost more than the quarter round they
// save, so it stays off.
static int use_tail() {
  const char* e = getenv("GANTTS_B200_TAIL");
  return e ? atoi(e) : 0;
}

static int launch_gemm_kk(const Planes& A, const Planes& B, const EpiArgs& e, cudaStream_t st) {
  if (use_tail() && use_cluster() == 0) {
    const int N = (int)B.rows;
    const int bn = pick_bn(N);
    const int num_b = (N + bn - 1) / bn;
    const int64_t num_a = (A.rows + TC_BM - 1) / TC_BM;
    const bool pair = num_b >= 2 && num_a >= 2;
    const int units = pair ? num_sms() / 2 : num_sms();
    const int unit_rows = pair ? 2 * TC_BM : TC_BM;
    const int64_t row_tiles = (A.rows + unit_rows - 1) / unit_rows;
    const int64_t tiles = row_tiles * num_b;
    const int64_t rounds = tiles / units, rem = tiles % units;
    if (rounds >= 1 && rem > 0 && bn >= 128 && bn % 128 == 0) {
      const double frac = (double)rem / units;
      int best_k = 1;
      double best = 1.0;
      for (int k = 2; k <= 4; k *= 2) {
        if (bn / k < 64) break;
        const double cost = ceil(k * frac - 1e-9) / k + 0.04;        // + the second launch's fill/drain
        if (cost < best) { best = cost; best_k = k; }
      }
      const int64_t main_row_tiles = rounds * units / num_b;
      const int64_t main_rows = main_row_tiles * unit_rows;
      if (best_k > 1 && main_rows > 0 && main_rows < A.rows) {
        Planes A1 = A, A2 = A;
        A1.rows = main_rows;
        A2.rows = A.rows - main_rows;
        A2.hi += main_rows * A.pitch;
        A2.lo += main_rows * A.pitch;
        EpiArgs e2 = e;
        e2.row0 = e.row0 + main_rows;
        if (e2.C) e2.C += main_rows * e.ldc;
        if (e2.out_hi) { e2.out_hi += main_rows * e.out_pitch; e2.out_lo += main_rows * e.out_pitch; }
        if (e2.code) e2.code += main_rows * e.code_pitch;
        int rc = launch_gemm_kk_one(A1, B, e, st);
        if (rc) return rc;
        return launch_gemm_kk_one(A2, B, e2, st, bn / best_k, false);
      }
    }
  }
  return launch_gemm_kk_one(A, B, e, st);
}

// C[n][k] (+)= sum_m A[m][n] * B[m][k]  (MN-major planes A [red][rows_a], B [red][cols_b]);
// split over the reduction, partials in `partial`, reduced deterministically into C (ld = cols_b).
// CTA pairs for the weight-gradient GEMM (gemm_pair_mn_kernel) need each half of the B tile to be whole
// 64-wide atoms.  Measured on cfg2 (profiles/r01_gemm_experiments.md): correct but 4 % SLOWER than the
// single-CTA kernel (200 vs 192 us for the four generator weight gradients) although it stages a third
// fewer operand bytes per SM -- so it is opt-in (GANTTS_B200_CLUSTER=3), not the default.
static bool mn_pair_ok(int cols_b) { return pick_bn(cols_b) % 128 == 0 && use_cluster() == 3; }

static size_t mn_partial_bytes(int64_t red, int rows_a, int cols_b, int* splits_out, int64_t* chunk_out) {
  int bn = pick_bn(cols_b);
  int tiles = ((rows_a + TC_BM - 1) / TC_BM) * ((cols_b + bn - 1) / bn);
  if (mn_pair_ok(cols_b)) tiles = 2 * ((((rows_a + TC_BM - 1) / TC_BM) + 1) / 2) * ((cols_b + bn - 1) / bn);
  const int bk = stage_bk(true);
  int64_t blocks = (red + bk - 1) / bk;
  int64_t splits = num_sms() / tiles;
  if (splits < 1) splits = 1;
  if (splits > blocks) splits = blocks;
  int64_t chunk = (blocks + splits - 1) / splits * bk;
  splits = (red + chunk - 1) / chunk;
  if (splits_out) *splits_out = (int)splits;
  if (chunk_out) *chunk_out = chunk;
  return ((size_t)splits * ((size_t)rows_a * cols_b + rows_a) * sizeof(float) + 255) / 256 * 256;
}

// Deferred deterministic split reductions: several (partial -> out) jobs summed by ONE launch.
constexpr int REDUCE_MAX_JOBS = 2 * GANTTS_MAX_LAYERS;
struct ReduceList {
  int n = 0;
  const float* partial[REDUCE_MAX_JOBS];
  float* out[REDUCE_MAX_JOBS];
  int splits[REDUCE_MAX_JOBS];
  int64_t len[REDUCE_MAX_JOBS];
  int64_t off[REDUCE_MAX_JOBS + 1];
};

// out[j][e] = sum_z partial[j][z][e] for a list of (partial, out) pairs: the split-K reduction of a model's weight
// gradients in one launch.  A block covers 32 float4 columns; its 8 warps take the splits z = w, w+8, ... (all loads of a
// warp independent and in flight together), the per-warp sums meet in shared memory and are added in warp order -- a fixed
// summation order, so the result is deterministic.  (The first version walked the splits sequentially per thread: 24 us for
// 42 MB of L2-resident partials at cfg2, 3.8 % issue utilisation -- profiles/r02_launches.md.)
constexpr int MR_WARPS = 8;
__global__ void __launch_bounds__(32 * MR_WARPS) multi_reduce_kernel(ReduceList rl, int accumulate) {
  pdl_entry();
  __shared__ float4 part_s[MR_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t total4 = rl.off[rl.n];                        // in units of 4 elements
  for (int64_t base = (int64_t)blockIdx.x * 32; base < total4; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + lane;
    int j = 0;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t e = 0, n = 0;
    bool vec = false;
    if (i < total4) {
      while (j + 1 < rl.n && i >= rl.off[j + 1]) ++j;
      e = (i - rl.off[j]) * 4;
      n = rl.len[j];
      vec = e + 3 < n && (n & 3) == 0;
      const float* part = rl.partial[j];
      const int splits = rl.splits[j];
      if (vec) {
#pragma unroll 4
        for (int z = w; z < splits; z += MR_WARPS) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(part + (int64_t)z * n + e));
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
      } else {
        for (int z = w; z < splits; z += MR_WARPS) {
          const float* pz = part + (int64_t)z * n + e;
          if (e < n) s.x += pz[0];
          if (e + 1 < n) s.y += pz[1];
          if (e + 2 < n) s.z += pz[2];
          if (e + 3 < n) s.w += pz[3];
        }
      }
    }
    part_s[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < total4) {
      float4 t = part_s[0][lane];
#pragma unroll
      for (int k = 1; k < MR_WARPS; ++k) {
        const float4 v = part_s[k][lane];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      float* out = rl.out[j];
      if (vec) {
        float4* o = reinterpret_cast<float4*>(out + e);
        if (accumulate) { const float4 c = *o; t.x += c.x; t.y += c.y; t.z += c.z; t.w += c.w; }
        *o = t;
      } else {
        const float tv[4] = {t.x, t.y, t.z, t.w};
        for (int q = 0; q < 4; ++q)
          if (e + q < n) out[e + q] = accumulate ? out[e + q] + tv[q] : tv[q];
      }
    }
    __syncthreads();
  }
}

static int flush_reduce(ReduceList& rl, int accumulate, cudaStream_t st) {
  if (rl.n == 0) return GANTTS_OK;
  rl.off[0] = 0;
  for (int j = 0; j < rl.n; ++j) rl.off[j + 1] = rl.off[j] + (rl.len[j] + 3) / 4;
  int64_t nb = (rl.off[rl.n] + 31) / 32;
  if (nb > (int64_t)num_sms() * 16) nb = (int64_t)num_sms() * 16;
  GANTTS_PDL_LAUNCH((multi_reduce_kernel), (unsigned)nb, 32 * MR_WARPS, 0, st, rl, accumulate);
  GANTTS_LAUNCH_CHECK("multi_reduce_kernel");
  rl.n = 0;
  return GANTTS_OK;
}

// gb (optional) receives the column sums of A (the bias gradient) from the same launch.  With `defer`
// the split reductions are queued (the partial buffer must then stay untouched until flush_reduce).
static int launch_gemm_mn(const Planes& A, const Planes& B, float* C, float* gb, int accumulate, float* partial,
                          cudaStream_t st, ReduceList* defer = nullptr) {
  GemmParams p{};
  p.rows_a = A.cols;
  p.cols_b = (int)B.cols;
  p.red = A.rows;
  if (B.rows != A.rows) {
    set_error("gemm_mn: reduction extents differ (%lld vs %lld)", (long long)A.rows, (long long)B.rows);
    return GANTTS_E_BADARG;
  }
  int splits;
  int64_t chunk;
  mn_partial_bytes(p.red, (int)p.rows_a, p.cols_b, &splits, &chunk);
  p.red_chunk = chunk;
  p.num_z = splits;
  p.bn = pick_bn(p.cols_b);
  p.num_a = (int)((p.rows_a + TC_BM - 1) / TC_BM);
  p.num_b = (p.cols_b + p.bn - 1) / p.bn;
  const int nb_atoms = (p.bn + 63) / 64;
  p.bk = stage_bk(true);
  p.atom_bytes = (uint32_t)p.bk * 128;                      // [bk reduction rows][128 B of 64 MN elements]
  p.a_plane = (TC_BM / 64) * p.atom_bytes;
  p.sbo = 1024;
  p.kstep = 2048;
  p.desc_layout = 2;
  const bool pair = mn_pair_ok(p.cols_b);
  p.b_plane_bytes = (uint32_t)(pair ? nb_atoms / 2 : nb_atoms) * p.atom_bytes;
  p.stage_bytes = 2 * p.a_plane + 2 * p.b_plane_bytes;
  p.tx_bytes = p.stage_bytes;
  p.num_stages = (int)((212 * 1024) / p.stage_bytes);
  if (p.num_stages > TC_MAX_STAGES) p.num_stages = TC_MAX_STAGES;
  p.ones_off = (uint32_t)p.num_stages * p.stage_bytes;
  p.tmem_cols = 512;
  const bool direct = (splits == 1 && !accumulate);
  const int64_t n = (int64_t)p.rows_a * p.cols_b;
  float* db_partial = partial + (int64_t)splits * n;
  EpiArgs e;
  e.epi = EPI_F32;
  e.C = direct ? C : partial;
  e.ldc = p.cols_b;
  fill_epilogue(p, e);
  p.c_zstride = n;
  p.db = gb ? (direct ? gb : db_partial) : nullptr;
  if (!pair) maybe_stage_f32(p, e, true, 212 * 1024);
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map(&mAh, A.hi, A.rows, A.cols, A.pitch, p.bk))) return rc;
  if ((rc = make_map(&mAl, A.lo, A.rows, A.cols, A.pitch, p.bk))) return rc;
  if ((rc = make_map(&mBh, B.hi, B.rows, B.cols, B.pitch, p.bk))) return rc;
  if ((rc = make_map(&mBl, B.lo, B.rows, B.cols, B.pitch, p.bk))) return rc;
  if (pair) {
    if ((rc = launch_pair_mn_kernel(mAh, mAl, mBh, mBl, p, st))) return rc;
  } else if ((rc = launch_kernel<true, EPI_F32, 1>(mAh, mAl, mBh, mBl, p, st))) {
    return rc;
  }
  if (!direct) {
    if (defer && defer->n + 2 <= REDUCE_MAX_JOBS) {
      int j = defer->n++;
      defer->partial[j] = partial; defer->out[j] = C; defer->splits[j] = splits; defer->len[j] = n;
      if (gb) {
        j = defer->n++;
        defer->partial[j] = db_partial; defer->out[j] = gb; defer->splits[j] = splits; defer->len[j] = p.rows_a;
      }
      return GANTTS_OK;
    }
    splitk_reduce_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, st>>>(partial, splits, n, C, accumulate);
    GANTTS_LAUNCH_CHECK("splitk_reduce_kernel(tc gW)");
    if (gb) {
      splitk_reduce_kernel<<<(unsigned)((p.rows_a + 1023) / 1024), 256, 0, st>>>(db_partial, splits, p.rows_a, gb,
                                                                               accumulate);
      GANTTS_LAUNCH_CHECK("splitk_reduce_kernel(tc gb)");
    }
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
  e.epi = EPI_F32;
  e.C = y;
  e.ldc = y_rs;
  e.bias = bias;
  e.act = act;
  e.slope = slope;
  e.p = act == GANTTS_ACT_LEAKY_DROPOUT ? p : 0.f;
  e.seed = seed;
  return launch_gemm_kk(X, Wp, e, st);
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
    e.epi = EPI_F32;
    e.C = gx;
    e.ldc = gx_rs;
    if ((rc = launch_gemm_kk(G, Wt, e, st))) return rc;             // gx[m][k] = sum_n gz[m][n] W[n][k]
  }
  if (gW) {
    if ((rc = launch_split(x, x_rs, M, K, X, 0, st))) return rc;
    if ((rc = launch_gemm_mn(G, X, gW, nullptr, accumulate, partial, st))) return rc;
  }
  return GANTTS_OK;
}

}  // namespace gantts
