// On-chip fused MLP stack on CTA pairs (tcgen05 cta_group::2): the whole chain of K-major GEMMs of a narrow MLP
// (the discriminator of reference gantts/models.py:121-141: 58 -> 256 -> 256 -> 256 -> 1) per 256-row tile WITHOUT the
// activation round trip through HBM between layers (SURVEY.md 2.2 K1 fusion target).
//
//   forward  (BWD = false):  H_{l+1} = Dropout(LeakyReLU(H_l W_l^T + b_l)),  y = sigmoid(H_L w + b)
//       A_0 (input planes) arrives by TMA; the epilogue of layer l converts the fp32 accumulator (TMEM) into bf16
//       hi/lo planes and writes them into shared memory in the UMMA K-major SWIZZLE_128B layout: that tile is (a) the
//       A operand of layer l+1 and (b) the source of a TMA bulk store to the tape in HBM (planes for the backward;
//       the 2-bit derivative codes go out with ordinary stores).  The last hidden layer's epilogue folds the
//       single-output Linear + sigmoid (a GEMV) in.
//   backward (BWD = true):   gZ_{l-1} = (gZ_l W_l) * act'(H_l),  gx = gZ_0 W_0
//       the head gZ_L = (gy * sigma'(y)) w^T * act'(H_L) is produced by the epilogue warps straight into shared
//       memory (it is elementwise in gy, w and the saved codes); gradient planes are stored to HBM only when the
//       weight-gradient GEMMs need them; the last layer's fp32 result may be accumulated into a column window of a
//       wider matrix (the scatter into g_static of the fused step).
//
// Only the weights stream (L2-resident, 0.6 MB): per pair tile the L2 -> SM operand traffic halves and the HBM
// traffic drops to the tape writes.  Pipeline per CTA pair (persistent, one pair tile = 2 x 128 rows):
//   warp 0      TMA producer: ring of 16 KB slots carrying, in consumption order, [A_0 hi, A_0 lo,] then for every layer
//               and every 32-wide reduction block the CTA's half of the weight tile (N/2 rows, hi+lo, SWIZZLE_64B)
//   warp 1      MMA issuer (leader CTA): 3 tcgen05.mma (hi*hi, hi*lo, lo*hi) per K=16 step, accumulators ping-pong
//               between two 256-column TMEM buffers, so the MMAs of layer l+1 start as soon as the epilogue of layer l
//               has delivered the FIRST 64-column chunk of its output (chunk c of the output = reduction chunk c of l+1)
//   warps 2-17  epilogue: all sixteen warps work on the same 64-column chunk (16 columns each), chunk after chunk
//   warp 18     plane store: one thread issues the TMA bulk stores of each finished chunk and tells the epilogue warps
//               when the shared-memory tile may be overwritten
#include "common.cuh"
#include "tc_ptx.cuh"

namespace gantts {

constexpr int CH_MAX_LAYERS = 4;               // MMA layers per chain
constexpr int CH_THREADS = TC_THREADS + 32;    // + the plane-store warp
constexpr int CH_STORE_WARP = TC_THREADS / 32;
constexpr uint32_t CH_SLOT = 16384;            // ring slot: 128 rows x 128 B (A_0 plane) or 2 x [128 rows x 64 B] (B hi|lo)
constexpr int CH_SLOTS = 5;
constexpr uint32_t CH_A_CHUNK = 32768;         // one 64-wide K chunk of A: hi 16 KB | lo 16 KB
constexpr uint32_t CH_A_BYTES = 4 * CH_A_CHUNK;
constexpr uint32_t CH_OFF_RING = CH_A_BYTES;
constexpr uint32_t CH_OFF_BIAS = CH_OFF_RING + CH_SLOTS * CH_SLOT;     // [CH_MAX_LAYERS][256] floats
constexpr uint32_t CH_OFF_VEC = CH_OFF_BIAS + CH_MAX_LAYERS * 1024;    // [256] floats: GEMV weights (fwd) / head weights (bwd)
constexpr uint32_t CH_OFF_PART = CH_OFF_VEC + 1024;                    // [4][128] floats: GEMV partial sums
constexpr uint32_t CH_OFF_BAR = CH_OFF_PART + 2048;
constexpr uint32_t CH_SMEM = CH_OFF_BAR + 256 + 1024;                  // + alignment slack

struct ChainLayer {
  int N;                 // output columns, padded to a multiple of 64 (<= 256)
  int n_valid;           // real output columns
  int K;                 // reduction extent, padded to a multiple of 64 (<= 256)
  const float* bias;     // fwd: [n_valid]
  int store_planes;      // the layer's output planes go to HBM (ChainMaps::s_hi/s_lo: fwd tape H_{l+1}; bwd gZ planes)
  uint32_t* code;        // fwd: derivative codes of this layer's output (written); bwd: codes of H_l (read), null on the last
  int64_t code_pitch;
  uint64_t seed;         // fwd: dropout seed of this layer
};

struct ChainParams {
  int64_t M;
  int num_layers;
  ChainLayer L[CH_MAX_LAYERS];
  float slope, keep_scale;
  uint32_t thresh;
  // forward tail: y = act(H_last w + b) (single output column)
  const float* w_last;
  const float* b_last;
  float* y;
  int64_t y_rs;
  int sigmoid;
  // backward head: gZ_head[r][c] = gy[r] (* y (1 - y)) * w_head[c] * act'(code_head[r][c]), K_0 = head_cols
  const float* gy;
  int64_t gy_rs;
  const float* yv;
  int64_t yv_rs;
  const float* w_head;
  int head_valid;                       // real number of head columns (w_head entries)
  const uint32_t* code_head;
  int64_t code_head_pitch;
  // backward tail: fp32 result of the last layer, rows >= c_row0 only
  float* C;
  int64_t ldc;
  int64_t c_row0;
  int c_accumulate;
  uint32_t dbg;   // timing experiments (GANTTS_B200_CHAIN_DBG): 1 no plane/code stores, 2 no dropout hash, 4 no proxy fence,
                  // 8 no shared-memory A stores, 16 no MMAs -- RESULTS ARE WRONG with any bit set
};

struct ChainMaps {
  CUtensorMap a_hi, a_lo;                                   // forward input planes, box [128 rows][64]
  CUtensorMap b_hi[CH_MAX_LAYERS], b_lo[CH_MAX_LAYERS];     // weight planes, box [N/2 rows][32]
  CUtensorMap s_hi[CH_MAX_LAYERS], s_lo[CH_MAX_LAYERS];     // output planes of layer l (store), box [128 rows][64]
};

namespace ptx {
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// cluster-scope acquire: the data published before the (remote) arrive lives in the PEER CTA's shared memory
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("gantts_b200: chain mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 2D tile store shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}  // namespace ptx

// 16 fp32 values of row `r` (0..127 within the CTA), columns [kcol, kcol+16) of a 64-wide chunk -> the chunk's hi and lo
// tiles in shared memory, UMMA/TMA K-major SWIZZLE_128B placement: 16-byte unit u of row r sits at unit u ^ (r & 7).
__device__ __forceinline__ void store_a_chunk16(uint32_t chunk_base, int r, int kcol, const float* v) {
  uint32_t h[8], l[8];
  split8(v, h, l);
  split8(v + 8, h + 4, l + 4);
  const uint32_t rowb = chunk_base + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
  const uint32_t u0 = (uint32_t)(kcol >> 3);
  const uint32_t a0 = rowb + ((u0 ^ (uint32_t)(r & 7)) << 4), a1 = rowb + (((u0 + 1) ^ (uint32_t)(r & 7)) << 4);
  ptx::st_shared_v4(a0, h[0], h[1], h[2], h[3]);
  ptx::st_shared_v4(a1, h[4], h[5], h[6], h[7]);
  ptx::st_shared_v4(a0 + 16384u, l[0], l[1], l[2], l[3]);
  ptx::st_shared_v4(a1 + 16384u, l[4], l[5], l[6], l[7]);
}

template <bool BWD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CH_THREADS, 1)
chain_pair_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (base - raw);                 // generic pointer to the aligned base
  const uint32_t ring0 = base + CH_OFF_RING;
  float* bias_s = reinterpret_cast<float*>(sgen + CH_OFF_BIAS);
  float* vec_s = reinterpret_cast<float*>(sgen + CH_OFF_VEC);
  float* part_s = reinterpret_cast<float*>(sgen + CH_OFF_PART);
  const uint32_t bar0 = base + CH_OFF_BAR;
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * CH_SLOTS;            // ring
  const uint32_t tfull0 = bar0 + 16 * CH_SLOTS, tempty0 = tfull0 + 16;  // accumulators (2 each)
  const uint32_t aready0 = tempty0 + 16;     // A chunk c written by the epilogue warps of BOTH CTAs (4, leader's copy)
  const uint32_t cready0 = aready0 + 32;     // chunk c written by THIS CTA's epilogue warps: ready for the plane store (4)
  const uint32_t sfree0 = cready0 + 32;      // the plane stores of a layer have finished reading shared memory (1)
  const uint32_t tmem_slot = sfree0 + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const bool leader = crank == 0;
  const int NL = p.num_layers;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < CH_SLOTS; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull0 + 8 * a, 1);
      ptx::mbar_init(tempty0 + 8 * a, 2 * TC_EPI_WARPS);     // both CTAs' epilogue warps (leader's copy is used)
    }
    for (int c = 0; c < 4; ++c) {
      ptx::mbar_init(aready0 + 8 * c, 2 * TC_EPI_WARPS);
      ptx::mbar_init(cready0 + 8 * c, TC_EPI_WARPS);
    }
    ptx::mbar_init(sfree0, 1);
    ptx::fence_barrier_init();
    if (!BWD) {
      ptx::prefetch_tensormap(&maps.a_hi);
      ptx::prefetch_tensormap(&maps.a_lo);
    }
    for (int l = 0; l < NL; ++l) {
      ptx::prefetch_tensormap(&maps.b_hi[l]);
      ptx::prefetch_tensormap(&maps.b_lo[l]);
      if (p.L[l].store_planes) {
        ptx::prefetch_tensormap(&maps.s_hi[l]);
        ptx::prefetch_tensormap(&maps.s_lo[l]);
      }
    }
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  if (!BWD) {
    for (int i = threadIdx.x; i < NL * 256; i += CH_THREADS) {
      const int l = i >> 8, c = i & 255;
      bias_s[i] = (c < p.L[l].n_valid && p.L[l].bias) ? p.L[l].bias[c] : 0.f;
    }
    if (p.w_last)
      for (int i = threadIdx.x; i < 256; i += CH_THREADS) vec_s[i] = i < p.L[NL - 1].n_valid ? p.w_last[i] : 0.f;
  } else {
    for (int i = threadIdx.x; i < 256; i += CH_THREADS) vec_s[i] = (i < p.head_valid && p.w_head) ? p.w_head[i] : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + (tmem_slot - base));

  const int64_t num_tiles = (p.M + 2 * TC_BM - 1) / (2 * TC_BM);
  const int64_t first_tile = blockIdx.x / 2, tile_step = gridDim.x / 2;
  // backward: the last layer (input gradient) is only needed for rows >= c_row0
  auto layers_of_tile = [&](int64_t tile) -> int {
    if (BWD && p.C == nullptr) return NL - 1;
    if (BWD && (tile + 1) * (2 * TC_BM) <= p.c_row0) return NL - 1;
    return NL;
  };
  const bool no_store = (p.dbg & 1u) != 0;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs)
      uint32_t s = 0, ph = 0;
      auto next_slot = [&]() { if (++s == (uint32_t)CH_SLOTS) { s = 0; ph ^= 1; } };
      for (int64_t tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int32_t row0 = (int32_t)(tile * 2 * TC_BM + (int64_t)crank * TC_BM);
        if (!BWD) {
          for (int pl = 0; pl < 2; ++pl) {
            ptx::mbar_wait(empty0 + 8 * s, ph ^ 1);
            if (leader) ptx::mbar_expect_tx(full0 + 8 * s, 2 * CH_SLOT);
            ptx::tma_load_2d_pair(ring0 + s * CH_SLOT, pl == 0 ? &maps.a_hi : &maps.a_lo, ptx::mapa(full0 + 8 * s, 0), 0,
                                  row0);
            next_slot();
          }
        }
        const int nl = layers_of_tile(tile);
        for (int l = 0; l < nl; ++l) {
          const int half_n = p.L[l].N / 2;
          const uint32_t plane = (uint32_t)half_n * 64u;                      // bytes of one B plane tile in a slot
          const int32_t brow0 = (int32_t)crank * half_n;
          for (int k0 = 0; k0 < p.L[l].K; k0 += 32) {
            ptx::mbar_wait(empty0 + 8 * s, ph ^ 1);
            if (leader) ptx::mbar_expect_tx(full0 + 8 * s, 4 * plane);         // hi + lo of both CTAs
            const uint32_t fb = ptx::mapa(full0 + 8 * s, 0);
            ptx::tma_load_2d_pair(ring0 + s * CH_SLOT, &maps.b_hi[l], fb, k0, brow0);
            ptx::tma_load_2d_pair(ring0 + s * CH_SLOT + 8192u, &maps.b_lo[l], fb, k0, brow0);
            next_slot();
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ------------------------------------------------------------ MMA issuer (leader CTA, one thread)
      uint32_t s = 0, ph = 0;
      auto next_slot = [&]() { if (++s == (uint32_t)CH_SLOTS) { s = 0; ph ^= 1; } };
      uint32_t acc_cnt = 0;              // accumulator uses so far (buffer = acc_cnt & 1)
      uint32_t ar_par = 0;               // bit c = parity of the next phase of aready[c]
      for (int64_t tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int nl = layers_of_tile(tile);
        for (int l = 0; l < nl; ++l) {
          const int acc = acc_cnt & 1;
          ptx::mbar_wait(tempty0 + 8 * acc, ((acc_cnt >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256;
          const uint32_t idesc = ptx::make_idesc_bf16(2 * TC_BM, p.L[l].N, 0, 0);
          const int nchunks = p.L[l].K / 64;
          uint32_t first = 0;
          uint32_t a_slot_hi = 0, a_slot_lo = 0;
          const bool a_from_ring = !BWD && l == 0;
          if (a_from_ring) {
            a_slot_hi = s;
            ptx::mbar_wait(full0 + 8 * s, ph);
            next_slot();
            a_slot_lo = s;
            ptx::mbar_wait(full0 + 8 * s, ph);
            next_slot();
          }
          for (int kc = 0; kc < nchunks; ++kc) {
            uint32_t a_hi, a_lo;
            if (a_from_ring) {
              a_hi = ring0 + a_slot_hi * CH_SLOT;
              a_lo = ring0 + a_slot_lo * CH_SLOT;
            } else {
              ptx::mbar_wait_cluster(aready0 + 8 * kc, (ar_par >> kc) & 1u);
              ar_par ^= 1u << kc;
              a_hi = base + kc * CH_A_CHUNK;
              a_lo = a_hi + 16384u;
            }
            for (int half = 0; half < 2; ++half) {
              ptx::mbar_wait(full0 + 8 * s, ph);
              ptx::tc_fence_after();
              const uint32_t sb_hi = ring0 + s * CH_SLOT, sb_lo = sb_hi + 8192u;
#pragma unroll
              for (int k = 0; k < ((p.dbg & 16u) ? 0 : 2); ++k) {
                const uint32_t ao = (uint32_t)(half * 2 + k) * 32u, bo = (uint32_t)k * 32u;
                const uint64_t da_hi = ptx::make_smem_desc(a_hi + ao, 0u, 1024u, 2u);
                const uint64_t da_lo = ptx::make_smem_desc(a_lo + ao, 0u, 1024u, 2u);
                const uint64_t db_hi = ptx::make_smem_desc(sb_hi + bo, 0u, 512u, 4u);
                const uint64_t db_lo = ptx::make_smem_desc(sb_lo + bo, 0u, 512u, 4u);
                ptx::mma_bf16_ss_pair(d_tmem, da_hi, db_hi, idesc, first);
                first = 1;
                ptx::mma_bf16_ss_pair(d_tmem, da_hi, db_lo, idesc, 1);
                ptx::mma_bf16_ss_pair(d_tmem, da_lo, db_hi, idesc, 1);
              }
              ptx::mma_commit_pair(empty0 + 8 * s, (uint16_t)0x3);
              next_slot();
            }
          }
          if (a_from_ring) {
            ptx::mma_commit_pair(empty0 + 8 * a_slot_hi, (uint16_t)0x3);
            ptx::mma_commit_pair(empty0 + 8 * a_slot_lo, (uint16_t)0x3);
          }
          ptx::mma_commit_pair(tfull0 + 8 * acc, (uint16_t)0x3);
          ++acc_cnt;
        }
      }
    }
  } else if (warp == CH_STORE_WARP) {
    if (lane == 0 && !no_store) {
      // ------------------------------------------------------------ plane store (both CTAs, own 128 rows)
      uint32_t cr_par = 0;               // bit c = parity of the next phase of cready[c]
      for (int64_t tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int32_t row0 = (int32_t)(tile * 2 * TC_BM + (int64_t)crank * TC_BM);
        const int nl = layers_of_tile(tile);
        for (int l = 0; l < nl; ++l) {
          if (!p.L[l].store_planes) continue;
          const int nch = p.L[l].N / 64;
          for (int c = 0; c < nch; ++c) {
            ptx::mbar_wait(cready0 + 8 * c, (cr_par >> c) & 1u);
            cr_par ^= 1u << c;
            ptx::tma_store_2d(&maps.s_hi[l], base + c * CH_A_CHUNK, c * 64, row0);
            ptx::tma_store_2d(&maps.s_lo[l], base + c * CH_A_CHUNK + 16384u, c * 64, row0);
            ptx::bulk_commit();
          }
          ptx::bulk_wait_read0();         // shared memory of this layer's tile has been read out
          ptx::mbar_arrive(sfree0);
        }
      }
      ptx::bulk_wait0();                  // every store has landed before the CTA retires
    }
  } else {
    // -------------------------------------------------------------- epilogue warps (both CTAs, own 128 rows)
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int j = (warp - 2) >> 2;           // 16-column piece of every 64-column chunk
    const int r = q * 32 + lane;             // row within the CTA
    const uint32_t aready_leader = ptx::mapa(aready0, 0);
    const uint32_t tempty_leader = ptx::mapa(tempty0, 0);
    uint32_t acc_cnt = 0;
    uint32_t st_issued = 0, st_waited = 0;   // layers handed to the store warp / whose sfree phase has been consumed
    // before the A buffer is overwritten: every plane store issued so far must have read it out
    auto wait_stores = [&]() {
      while (st_waited < st_issued) {
        if (lane == 0) ptx::mbar_wait(sfree0, st_waited & 1u);
        __syncwarp();
        ++st_waited;
      }
    };
    // chunk c of the A buffer is complete as far as this warp is concerned
    auto publish = [&](int c, bool to_mma, bool to_store) {
      if (!(p.dbg & 4u)) ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (to_mma) {
          if (leader) ptx::mbar_arrive(aready0 + 8 * c);
          else ptx::mbar_arrive_cluster(aready_leader + 8 * c);
        }
        if (to_store) ptx::mbar_arrive(cready0 + 8 * c);
      }
    };
    // backward head of one tile: gZ_head chunk by chunk into shared memory
    auto head = [&](int64_t tile) {
      const int64_t row = tile * 2 * TC_BM + (int64_t)crank * TC_BM + r;
      const bool row_ok = row < p.M;
      float g = 0.f;
      if (row_ok) {
        g = p.gy[row * p.gy_rs];
        if (p.yv) {
          const float yy = p.yv[row * p.yv_rs];
          g *= yy * (1.f - yy);
        }
      }
      const int nch = p.L[0].K / 64;
      uint32_t cw[4] = {0u, 0u, 0u, 0u};
      if (row_ok)
        for (int c = 0; c < nch; ++c) cw[c] = __ldg(p.code_head + row * p.code_head_pitch + c * 4 + j);
      const float dpos = p.keep_scale, dneg = p.slope * p.keep_scale, dzero = p.thresh ? 0.f : p.slope;
      wait_stores();
      for (int c = 0; c < nch; ++c) {
        const int col = c * 64 + j * 16;
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const uint32_t ce = (cw[c] >> (2 * e)) & 3u;
          v[e] = g * vec_s[col + e] * ((ce & 1u) ? dzero : ((ce & 2u) ? dneg : dpos));
        }
        store_a_chunk16(base + c * CH_A_CHUNK, r, j * 16, v);
        publish(c, true, false);
      }
    };
    if (BWD && first_tile < num_tiles) head(first_tile);
    for (int64_t tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int64_t row = tile * 2 * TC_BM + (int64_t)crank * TC_BM + r;
      const bool row_ok = row < p.M;
      const int nl = layers_of_tile(tile);
      for (int l = 0; l < nl; ++l) {
        const ChainLayer& L = p.L[l];
        const int acc = acc_cnt & 1;
        const bool last = l == nl - 1;
        const bool feeds_next = !last;                       // output is the A operand of layer l+1
        const bool is_out = BWD && l == NL - 1 && p.C != nullptr;          // fp32 input gradient, no planes
        const bool to_store = !is_out && L.store_planes && !no_store;
        const bool to_smem = (feeds_next || to_store) && !(p.dbg & 8u);
        const int nch = L.N / 64;
        // backward: derivative codes of this layer's output columns (act' of H_l), prefetched before the wait
        uint32_t cw[4] = {0u, 0u, 0u, 0u};
        if (BWD && L.code && row_ok)
          for (int c = 0; c < nch; ++c) cw[c] = __ldg(L.code + row * L.code_pitch + c * 4 + j);
        ptx::mbar_wait(tfull0 + 8 * acc, (acc_cnt >> 1) & 1);
        ptx::tc_fence_after();
        // backward: the A buffer is free once the LAST layer's MMAs are done -> produce the next tile's head first,
        // so that its first MMAs run under this tile's last epilogue
        // (only when this layer does not stage planes through the A buffer itself: otherwise the head follows it)
        const bool head_early = BWD && last && !to_store && tile + tile_step < num_tiles;
        const bool head_late = BWD && last && to_store && tile + tile_step < num_tiles;
        if (head_early) head(tile + tile_step);
        if (to_smem) wait_stores();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
        float gemv = 0.f;
        // one 16-column piece of chunk c: accumulator registers -> activation / gradient -> shared memory (+ codes)
        auto process = [&](int c, const uint32_t (&src)[16]) {
          const int col = c * 64 + j * 16;
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(src[e]);
          if (!BWD) {
            // reference gantts/models.py:137-139: Dropout(LeakyReLU(Linear(x)))
            const float* bs = bias_s + l * 256 + col;
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bs + e);
              v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], v[e] * p.slope);
            if (p.thresh && !(p.dbg & 2u)) {
              const uint32_t quarter_n = (uint32_t)(L.n_valid + 3) >> 2, limit = drop_limit(p.thresh);
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const DropBits d = dropout_quad_bits(L.seed, (uint32_t)row, quarter_n, (uint32_t)(col + e) >> 2);
                v[e] = (d.a << 16) > limit ? v[e] * p.keep_scale : 0.f;
                v[e + 1] = d.a > limit ? v[e + 1] * p.keep_scale : 0.f;
                v[e + 2] = (d.b << 16) > limit ? v[e + 2] * p.keep_scale : 0.f;
                v[e + 3] = d.b > limit ? v[e + 3] * p.keep_scale : 0.f;
              }
            }
            if (L.n_valid < L.N) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (col + e >= L.n_valid) v[e] = 0.f;
            }
            uint32_t code = 0;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if (v[e] == 0.f) code |= 1u << (2 * e);
              if (v[e] < 0.f) code |= 2u << (2 * e);
            }
            if (to_smem) store_a_chunk16(base + c * CH_A_CHUNK, r, j * 16, v);
            if (L.code && row_ok && col < L.n_valid && !no_store) L.code[row * L.code_pitch + c * 4 + j] = code;
            if (last && p.w_last) {
#pragma unroll
              for (int e = 0; e < 16; ++e) gemv = fmaf(v[e], vec_s[col + e], gemv);
            }
          } else if (!is_out) {
            // gZ_{l-1} = acc * act'(H_l): derivative class from the saved 2-bit codes
            const float dpos = p.keep_scale, dneg = p.slope * p.keep_scale, dzero = p.thresh ? 0.f : p.slope;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const uint32_t ce = (cw[c] >> (2 * e)) & 3u;
              v[e] *= (ce & 1u) ? dzero : ((ce & 2u) ? dneg : dpos);
            }
            if (to_smem) store_a_chunk16(base + c * CH_A_CHUNK, r, j * 16, v);
          } else {
            // input gradient: fp32, rows >= c_row0, optionally accumulated into a column window of a wider matrix
            if (row_ok && row >= p.c_row0) {
              float* crow = p.C + (row - p.c_row0) * p.ldc + col;
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (col + e < L.n_valid) crow[e] = p.c_accumulate ? crow[e] + v[e] : v[e];
            }
          }
          if (feeds_next || to_store) publish(c, feeds_next, to_store);
        };
        // two register sets: the next chunk's accumulator columns travel (tcgen05.ld) while this one is processed
        uint32_t ra[16], rb[16];
        ptx::tmem_ld16(taddr + j * 16, ra);
        for (int c = 0; c < nch; c += 2) {
          ptx::tmem_ld_wait();
          if (c + 1 < nch) ptx::tmem_ld16(taddr + (c + 1) * 64 + j * 16, rb);
          process(c, ra);
          if (c + 1 < nch) {
            ptx::tmem_ld_wait();
            if (c + 2 < nch) ptx::tmem_ld16(taddr + (c + 2) * 64 + j * 16, ra);
            process(c + 1, rb);
          }
        }
        if (to_store) ++st_issued;
        if (head_late) head(tile + tile_step);
        // accumulator drained
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) ptx::mbar_arrive(tempty0 + 8 * acc);
          else ptx::mbar_arrive_cluster(tempty_leader + 8 * acc);
        }
        ++acc_cnt;
        if (!BWD && last && p.w_last) {
          // y = act(H w + b): four 64-column partial sums per row through shared memory
          part_s[j * 128 + r] = gemv;
          ptx::named_bar_sync(1, 32 * TC_EPI_WARPS);
          if (j == 0 && row_ok) {
            const float z = part_s[r] + part_s[128 + r] + part_s[256 + r] + part_s[384 + r] + p.b_last[0];
            p.y[row * p.y_rs] = p.sigmoid ? 1.f / (1.f + expf(-z)) : z;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------- host side
// Where the chain kernel replaces the per-layer launches (GANTTS_B200_CHAIN, bit mask): 1 = forward, 2 = backward
// without weight gradients (the adversarial pass: nothing but the input gradient leaves the chip), 4 = backward with
// weight gradients (the gradient planes still go to HBM for the weight-gradient GEMMs).
// DEFAULT OFF.  Measured on B200 at cfg2 (profiles/r02_chain.md): the kernel is correct (the whole -m gpu suite passes
// with modes 1|2) but SLOWER in the step than the per-layer launches it replaces (1.345 vs 1.273 ms/step with 1|2,
// 1.445 with 7; tensor pipe 17 % active): inside one tile the layers are strictly serial -- epilogue of layer l, then
// MMAs of layer l+1 -- and a 128 x 256 activation tile (128 KB as hi/lo planes) leaves no shared memory for a second
// tile in flight, whereas the per-layer kernels overlap the epilogue of tile i with the MMAs of tile i+1.
constexpr int CHAIN_FWD = 1, CHAIN_BWD_NOGRAD = 2, CHAIN_BWD_GRAD = 4;
static int use_chain() {          // read on every call (cheap) so that tests can switch it per case
  const char* e = getenv("GANTTS_B200_CHAIN");
  return e ? atoi(e) : 0;
}

static inline int pad64(int v) { return (v + 63) / 64 * 64; }

// Shapes the chain kernel covers: input width <= 64, hidden widths <= 256, a single output column, 2..4 hidden layers.
static bool chain_shape_ok(const gantts_mlp_t* m, int mode) {
  const int L = m->num_layers;
  if (!(use_chain() & mode) || L < 3 || L - 1 > CH_MAX_LAYERS) return false;   // >= 2 hidden layers (see mlp_bwd_impl)
  if (m->dims[L] != 1 || m->dims[0] > 64) return false;
  for (int l = 1; l < L; ++l)
    if (m->dims[l] > 256 || m->dims[l] < 16 || (m->dims[l] & 15)) return false;
  return true;
}

static uint32_t chain_dbg() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GANTTS_B200_CHAIN_DBG");
    v = e ? atoi(e) : 0;
    if (v)
      fprintf(stderr, "gantts_b200: GANTTS_B200_CHAIN_DBG=%d removes parts of the chain kernel (phase timing only): "
                      "RESULTS ARE WRONG, do not use for training or benchmarks\n", v);
  }
  return (uint32_t)v;
}

template <bool BWD>
static int launch_chain(const ChainMaps& maps, const ChainParams& p_in, cudaStream_t st) {
  ChainParams p = p_in;
  p.dbg = chain_dbg();
  static bool attr[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    GANTTS_CUDA(cudaFuncSetAttribute(chain_pair_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  const int64_t tiles = (p.M + 2 * TC_BM - 1) / (2 * TC_BM);
  int grid = (int)(tiles * 2 < num_sms() ? tiles * 2 : num_sms() / 2 * 2);
  double flops = 0.0;
  for (int l = 0; l < p.num_layers; ++l) {
    const bool tail = BWD && l == p.num_layers - 1;
    if (tail && !p.C) continue;
    flops += 2.0 * (double)(tail ? p.M - p.c_row0 : p.M) * p.L[l].n_valid * (double)p.L[l].K;
  }
  prof_begin(PROF_CHAIN, flops, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  int na = 0;
  if (use_pdl()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, chain_pair_kernel<BWD>, maps, p);
  prof_end(st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(chain)");
  GANTTS_LAUNCH_CHECK("chain_pair_kernel");
  return GANTTS_OK;
}

}  // namespace gantts
