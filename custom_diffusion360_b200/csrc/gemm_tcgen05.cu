// gemm_tcgen05.cu — persistent, warp-specialised bf16 GEMM / implicit-GEMM 3x3 convolution for
// sm_100a.  One CTA per SM (or one CTA PAIR per two SMs, cta_group::2); roles per CTA:
//   warp 0 (1 thread)  TMA producer: A and W tiles -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1 (1 thread)  tcgen05.mma issuer (leader CTA only in pair mode)
//   warp 2             TMEM allocator / deallocator
//   warps 4-11         epilogue: two warps per TMEM lane quarter, each taking half of the tile's
//                      columns, 64 columns ("slab") at a time:
//                        tcgen05.ld -> +bias / +per-image emb bias / SiLU / GEGLU (fp32 registers)
//                        + residual slab (TMA-loaded into 128B-swizzled smem, prefetched a slab
//                          ahead, i.e. during the next tile's main loop)
//                        -> bf16 slab in swizzled smem -> ONE TMA store per slab.
//                      Row-per-thread global loads/stores (32 sectors per warp instruction) made
//                      the epilogue slower than a K<=1280 main loop; the TMA path is fully
//                      coalesced and clips M/N tails by itself.  fp32 outputs and odd leading
//                      dimensions take the direct path.
// The accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.
//
// Pair mode (CG == 2): a cluster of two CTAs computes a 256 x BN tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256).  Each CTA loads its own 128 rows of A and HALF of the
// W tile (BN/2 rows), so per-SM operand traffic from L2 drops by a third.  Both CTAs' TMA loads
// complete on the leader's `full` barrier; the leader's MMA thread releases smem slots /
// publishes accumulators to both CTAs with multicast tcgen05.commit; the peer's epilogue warps
// hand TMEM back with remote mbarrier arrives.
//
// Convolution mode feeds the SAME main loop from a 4-D NHWC tensor map: the A tile of K-block
// (tap, c-chunk) is the box [64 ch, tw, th, tb] at pixel offset (dx-1, dy-1); out-of-image
// coordinates are zero-filled by TMA, which implements the pad-1 border without a halo copy.
//
// Reference arithmetic replaced: see include/cd360.h (cd360_gemm_bf16).
#include <stdlib.h>

#include "cd360_common.cuh"

// Phase tracing for tools/gemm_trace.py (never compiled into libcd360.so): per-CTA clock stamps.
#ifdef CD360_GEMM_TRACE
__device__ unsigned long long* g_cd360_trace = nullptr;
__device__ __forceinline__ unsigned long long cd360_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define CD360_TRACE(slot)                                                              \
  do {                                                                                 \
    if (g_cd360_trace != nullptr) g_cd360_trace[blockIdx.x * 32 + (slot)] = cd360_gtimer(); \
  } while (0)
// SM-clock stamps of epilogue warp 4 inside the LAST tile (slots 16..31)
#define CD360_TRACE_CLK(cond, slot)                                                    \
  do {                                                                                 \
    if ((cond) && g_cd360_trace != nullptr)                                            \
      g_cd360_trace[blockIdx.x * 32 + (slot)] = static_cast<unsigned long long>(clock64()); \
  } while (0)
extern "C" int cd360_gemm_set_trace(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_cd360_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#else
#define CD360_TRACE(slot) do {} while (0)
#define CD360_TRACE_CLK(cond, slot) do {} while (0)
#endif

namespace cd360 {

constexpr int BM = 128;  // rows per CTA
constexpr int BK = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int WSLAB_BYTES = 32 * 64 * 2;    // 32 rows x 64 bf16 = 4 KiB: one epilogue warp's staging slab

struct GemmKParams {
  int M, N, N_out;
  int num_m_blocks, num_n_blocks;
  int kb0, kb1;  // 64-wide K blocks of the two A segments (linear mode)
  int conv;      // 0 / 1
  int cblocks;   // C / 64 (conv)
  int H, W;      // conv geometry
  int tw, th, tb;  // conv A box (pixels x rows x images), tw*th*tb == 128
  const float* bias;
  const float* row_bias;
  int rows_per_group;
  long long ld_row_bias;
  const __nv_bfloat16* residual;
  long long ldr;
  void* out;
  long long ldo;
  int out_fp32;
  int act;
  int geglu;
  int epi_tma;  // 1: slab epilogue through smem + TMA (bf16 out, 16-byte aligned rows)
  // LayerNorm folded into the GEMM (gamma in W, beta in bias): out = rstd*(acc - mu*colsum) + bias
  const float* ln_stats;   // [M, ln_slabs, 2] partial (sum, sumsq) of every A row, or NULL
  int ln_slabs;
  float ln_inv_c, ln_eps;
  const float* ln_colsum;  // [N]
  float* stats_out;        // [M, N_out/64, 2] partial (sum, sumsq) of the bf16 output rows, or NULL
  int stats_slabs;
  // split-K (small-M GEMMs whose few tiles cannot pull the weights at HBM rate): tile index =
  // ks * (m_blocks * n_blocks) + mn; split ks covers k-blocks [ks*kb_per, min(nkb, (ks+1)*kb_per));
  // split ks stores its fp32 partial tile into slice ks of `out` ([ksplit][M][ldo] fp32, plain
  // stores: deterministic), summed in fixed order by cd360_splitk_finish
  int ksplit, kb_per;
  long long split_stride;  // elements between the partial slices
  int dyn_n;               // 1: the last n-block's MMAs are issued with its real width (CD360_GEMM_DYNN=0: full BN)
  // 1: both operands are stored contraction-major-OUTER ("TN": A^T as [K, M], W^T as [K, N], row-major) and
  // are consumed in place as MN-major UMMA operands: a stage holds, per operand, 64-column chunks
  // [64 K-rows][128 B] (one TMA box {64 MN, 64 K} each).  Weight gradients dW = dY^T X contract over the
  // token rows of two activation matrices; this replaces two transposing copies per GEMM.
  int tn;
};

template <int BN, int STAGES, int CG>
struct GemmSmem {
  static constexpr int BNC = BN / CG;  // W rows this CTA stages
  static constexpr int B_TILE_BYTES = BNC * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;       // [8 epilogue warps][2] 4 KiB slabs
  static constexpr int BIAS_OFFSET = EPI_OFFSET + kEpiWarps * 2 * WSLAB_BYTES;  // float[2][BN] tile bias
  static constexpr int BAR_OFFSET = BIAS_OFFSET + 2 * BN * 4;
  // full[STAGES] empty[STAGES] tmem_full[2] tmem_empty[2] res_full[8 warps][2] + tmem ptr
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4 + 2 * kEpiWarps) * 8 + 16;
  // no static __shared__ in the kernel: the dynamic window starts at offset 0 of the CTA's shared
  // memory and is 1024 B aligned (checked at kernel entry), so no alignment slack is reserved
  static constexpr int DYN_BYTES = TOTAL;
  static_assert(DYN_BYTES <= 227 * 1024, "smem budget exceeded");
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep 1024 B alignment");
};

// ---- epilogue helpers --------------------------------------------------------------------------
__device__ __forceinline__ void load_bias32(float (&v)[32], const float* __restrict__ p, int col0,
                                            int nvalid) {
  if (nvalid == 32) {
    const float4* p4 = reinterpret_cast<const float4*>(p + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = __ldg(p4 + j);
      v[4 * j + 0] += b.x;
      v[4 * j + 1] += b.y;
      v[4 * j + 2] += b.z;
      v[4 * j + 3] += b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid) v[j] += __ldg(p + col0 + j);
  }
}

// tile bias staged in shared memory by the epilogue warps while the main loop runs (a global
// load per 32-column chunk put an L2 round trip on the critical path of every tile)
__device__ __forceinline__ void add_bias32_smem(float (&v)[32], const float* sb) {
  const float4* p4 = reinterpret_cast<const float4*>(sb);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b = p4[j];
    v[4 * j + 0] += b.x;
    v[4 * j + 1] += b.y;
    v[4 * j + 2] += b.z;
    v[4 * j + 3] += b.w;
  }
}

// direct (row-per-thread) store path: fp32 outputs, odd leading dimensions, N tails < 8
__device__ __forceinline__ void store_chunk32(float (&v)[32], const GemmKParams& p, long long row,
                                              int ocol0, int nvalid, int ks = 0) {
  if (p.residual != nullptr) {
    const __nv_bfloat16* r = p.residual + row * p.ldr + ocol0;
    if (nvalid == 32) {
      const uint4* r4 = reinterpret_cast<const uint4*>(r);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u = r4[j];
        float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
               f3 = unpack_bf16x2(u.w);
        v[8 * j + 0] += f0.x; v[8 * j + 1] += f0.y; v[8 * j + 2] += f1.x; v[8 * j + 3] += f1.y;
        v[8 * j + 4] += f2.x; v[8 * j + 5] += f2.y; v[8 * j + 6] += f3.x; v[8 * j + 7] += f3.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += __bfloat162float(r[j]);
    }
  }
  if (p.out_fp32) {
    float* o = reinterpret_cast<float*>(p.out) + ks * p.split_stride + row * p.ldo + ocol0;
    if (nvalid == 32 && (p.ldo & 3) == 0) {
      float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) o[j] = v[j];
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + ocol0;
    if (nvalid == 32 && (p.ldo & 7) == 0) {
      uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
        u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        o4[j] = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) o[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// ---- kernel ------------------------------------------------------------------------------------
template <int BN, int STAGES, int CG, int MC>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0,
                         const __grid_constant__ CUtensorMap tmA1,
                         const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmOut,
                         const __grid_constant__ CUtensorMap tmRes, const GemmKParams p) {
  using L = GemmSmem<BN, STAGES, CG>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // swizzled TMA / UMMA tiles need 1024 B alignment
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + 2 * kEpiWarps);

  // Warp index broadcast from lane 0: provably warp-uniform, so the role branches are uniform branches
  // and the single-lane roles run as WHOLE warps with one elected lane issuing the asynchronous
  // instruction.  Under `lane == 0` (a divergent region) every cp.async.bulk.tensor / tcgen05.mma /
  // tcgen05.commit had its operands in vector registers and was wrapped by ptxas in an
  // ELECT + R2UR.BROADCAST x5 + BRA.U.ANY loop (~100 clk per instruction); with uniform control flow
  // the addresses and descriptors live in uniform registers and the UTCHMMAs issue back to back.
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // cluster = MC CTA pairs (or single CTAs); pair `pr` works on M block m_blk*MC + pr of the SAME
  // N block, so with MC == 2 the two pairs share every W tile (TMA multicast below)
  const uint32_t crank = (CG * MC > 1) ? cluster_ctarank() : 0u;
  const uint32_t rank = crank % CG;          // rank inside the pair
  const uint32_t pr = crank / CG;            // pair inside the cluster
  const uint32_t leader_rank = pr * CG;      // cluster rank of this pair's MMA-issuing CTA
  const bool leader = rank == 0;
  const int unit = blockIdx.x / (CG * MC);   // cluster (or lone CTA) index
  const int num_units = gridDim.x / (CG * MC);
  const int mn_tiles = p.num_m_blocks * p.num_n_blocks;
  const int num_tiles = mn_tiles * p.ksplit;
  const int nkb = p.conv ? 9 * p.cblocks : (p.kb0 + p.kb1);
  constexpr uint32_t TMEM_COLS = 2 * BN;

  CD360_TL(0);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  if (threadIdx.x == 0) CD360_TRACE(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    if (p.epi_tma) {
      tma_prefetch_desc(&tmOut);
      if (p.residual != nullptr) tma_prefetch_desc(&tmRes);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], CG);  // pair: leader's expect_tx arrive + peer's remote arrive
      mbar_init(&empty_bar[s], MC);  // one multicast commit per pair sharing the slot
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], kEpiWarps * CG);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int b = 0; b < 2 * kEpiWarps; ++b) mbar_init(&res_full[b], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) {
      tmem_alloc_2sm(tmem_ptr, TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG * MC > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) CD360_TRACE(1);
  pdl_wait();  // nothing above touches global memory written by earlier kernels
  if (threadIdx.x == 0) CD360_TRACE(2);

  if (warp == 0) {
    // ================================ TMA producer (whole warp; one elected lane issues) ========
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units) {
      const int ks = tile / mn_tiles;
      const int mn = tile - ks * mn_tiles;
      const int m_blk = mn % p.num_m_blocks;
      const int n_blk = mn / p.num_m_blocks;
      const int kb_begin = ks * p.kb_per;
      const int kb_end = min(nkb, kb_begin + p.kb_per);
      const int m0 = ((m_blk * MC + static_cast<int>(pr)) * CG + static_cast<int>(rank)) * BM;
      // N extent this tile really has (multiple of 16): the last n-block of N = 320 / 640 / 1920 is 64 / 128
      // columns wide, and the MMA is issued with THAT N (no tensor-core work on zero padding: 6.6 % of the
      // step's MMA work, and the step runs at the power cap).  In pair mode each CTA supplies N/2 rows of W
      // from the start of its smem tile, so CTA `rank` loads from row n_blk*BN + rank * ntile/2.
      const int ntile = (MC == 1 && p.dyn_n && !p.tn) ? min(BN, (p.N - n_blk * BN + 15) & ~15) : BN;
      const int n0 = n_blk * BN + static_cast<int>(rank) * (ntile / CG);
      int cb = 0, cy = 0, cx = 0;
      if (p.conv) {
        const int hw = p.H * p.W;
        cb = m0 / hw;
        const int rem = m0 - cb * hw;
        cy = rem / p.W;
        cx = rem - cy * p.W;
      }
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        uint8_t* sb = sa + A_TILE_BYTES;
        uint64_t* fb = &full_bar[stage];
        if (elect_one_sync()) {
        if (CG == 1) {
          mbar_arrive_expect_tx(fb, L::STAGE_BYTES);
        } else if (leader) {
          mbar_arrive_expect_tx(fb, 2 * L::STAGE_BYTES);
        } else {
          mbar_arrive_remote(fb, leader_rank);
        }
        if (p.tn) {
          // MN-major operands: chunk j = columns [64 j, 64 j + 64) of this CTA's M rows / W rows, 64 K-rows each
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) {
            if (CG == 2) tma_load_2d_2sm(sa + j * 8192, &tmA0, fb, m0 + j * 64, kb * BK);
            else tma_load_2d(sa + j * 8192, &tmA0, fb, m0 + j * 64, kb * BK);
          }
#pragma unroll
          for (int j = 0; j < L::BNC / 64; ++j) {
            if (CG == 2) tma_load_2d_2sm(sb + j * 8192, &tmB, fb, n0 + j * 64, kb * BK);
            else tma_load_2d(sb + j * 8192, &tmB, fb, n0 + j * 64, kb * BK);
          }
        } else {
        if (p.conv) {
          const int tap = kb / p.cblocks;
          const int kc = kb - tap * p.cblocks;
          const int dy = tap / 3, dx = tap - dy * 3;
          if (CG == 2) tma_load_4d_2sm(sa, &tmA0, fb, kc * BK, cx + dx - 1, cy + dy - 1, cb);
          else tma_load_4d(sa, &tmA0, fb, kc * BK, cx + dx - 1, cy + dy - 1, cb);
        } else {
          const CUtensorMap* tm = kb < p.kb0 ? &tmA0 : &tmA1;
          const int kk = (kb < p.kb0 ? kb : kb - p.kb0) * BK;
          if (CG == 2) tma_load_2d_2sm(sa, tm, fb, kk, m0);
          else tma_load_2d(sa, tm, fb, kk, m0);
        }
        if (MC == 2) {
          // this CTA fetches one half of its W rows and multicasts it to the CTA of the same pair
          // rank in the other pair (which fetches the other half): W crosses L2 -> SM once per
          // cluster instead of once per pair
          constexpr int QR = L::BNC / 2;
          tma_load_2d_2sm_mc(sb + pr * QR * 128, &tmB, fb, kb * BK, n0 + static_cast<int>(pr) * QR,
                             static_cast<uint16_t>((1u << rank) | (1u << (CG + rank))));
        } else if (CG == 2) {
          tma_load_2d_2sm(sb, &tmB, fb, kb * BK, n0);
        } else {
          tma_load_2d(sb, &tmB, fb, kb * BK, n0);
        }
        }
        if (tile == unit && kb == kb_begin) CD360_TRACE(3);
        if (tile + num_units >= num_tiles && kb == kb_end - 1) CD360_TRACE(4);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ================================ MMA issuer (whole warp; one elected lane issues) ==========
    int stage = 0;
    uint32_t phase = 0;
    int t = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units, ++t) {
      const int n_blk_t = (tile % mn_tiles) / p.num_m_blocks;
      const int ntile = (MC == 1 && p.dyn_n && !p.tn) ? min(BN, (p.N - n_blk_t * BN + 15) & ~15) : BN;   // see the producer
      const uint32_t idesc = make_idesc_bf16(BM * CG, ntile, p.tn != 0, p.tn != 0);
      const int buf = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * BN);
      const int kb_begin = (tile / mn_tiles) * p.kb_per;
      const int kb_end = min(nkb, kb_begin + p.kb_per);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (t == 0 && kb == kb_begin) CD360_TRACE(5);
        if (t == 0 && kb == kb_begin + 1) CD360_TRACE(6);
        const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_TILE_BYTES;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 K-elements = 32 B inside the 128-byte rows; MN-major: 16 K-rows = 2048 B
            const uint64_t adesc = p.tn ? make_smem_desc_sw128_mn(a_addr + k * 2048, 8192)
                                        : make_smem_desc_sw128(a_addr + k * 32);
            const uint64_t bdesc = p.tn ? make_smem_desc_sw128_mn(b_addr + k * 2048, 8192)
                                        : make_smem_desc_sw128(b_addr + k * 32);
            const uint32_t accumulate = (kb != kb_begin || k != 0) ? 1u : 0u;
            if (CG == 2) umma_bf16_2sm(tmem_d, adesc, bdesc, idesc, accumulate);
            else umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
          }
          // free the smem slot (in both CTAs) once these MMAs retire
          if (CG == 2) umma_commit_2sm(&empty_bar[stage], static_cast<uint16_t>((1u << (CG * MC)) - 1u));
          else umma_commit(&empty_bar[stage]);
          if (kb == kb_end - 1) {
            if (CG == 2) umma_commit_2sm(&tmem_full[buf], static_cast<uint16_t>(0x3u << (pr * 2)));
            else umma_commit(&tmem_full[buf]);
            if (tile + num_units >= num_tiles) CD360_TRACE(7);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int q = warp & 3;               // TMEM lane quarter this warp may access (warp % 4)
    const int half = (warp - 4) >> 2;     // which half of the tile's columns
    const int row_in_tile = q * 32 + lane;
    // Each warp stages ITS 32 rows x 64 columns in its own two 4 KiB buffers and issues its own TMA
    // stores / residual loads (box 64 x 32): no cross-warp barrier anywhere in the slab loop.  The
    // residual slab is TMA-loaded into the buffer, added in registers, and the result is written
    // back IN PLACE and stored from there; the two buffers alternate slab by slab.
    uint8_t* wbuf = smem + L::EPI_OFFSET + (warp - 4) * 2 * WSLAB_BYTES;
    uint64_t* wres = res_full + (warp - 4) * 2;
    const int sw = lane & 7;
    const bool use_res = p.epi_tma && p.residual != nullptr;
    // out slabs this half produces per tile, and the slab's first OUTPUT column inside the tile
    constexpr int SLABS = BN / 128;                 // plain: 64-col slabs per half
    const int n_slabs = p.geglu ? (BN / 256 > 0 ? BN / 256 : 1) : SLABS;
    const int out_tile_cols = p.geglu ? BN / 2 : BN;
    auto slab_col = [&](int n_blk, int s) {         // first output column of slab s of this half
      return n_blk * out_tile_cols + (half * n_slabs + s) * 64;
    };
    uint32_t slab_cnt = 0;  // slabs this warp has processed: buffer = cnt & 1, parity = (cnt >> 1) & 1
    if (use_res && unit < num_tiles) {
      const int m_blk = unit % p.num_m_blocks, n_blk = unit / p.num_m_blocks;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&wres[0], WSLAB_BYTES);
        tma_load_2d(wbuf, &tmRes, &wres[0], slab_col(n_blk, 0),
                    ((m_blk * MC + static_cast<int>(pr)) * CG + static_cast<int>(rank)) * BM + q * 32);
      }
      __syncwarp();
    }
    int t = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units, ++t) {
      const int ks = tile / mn_tiles;   // split-K: partial tile ks of output tile mn
      const int mn = tile - ks * mn_tiles;
      const int m_blk = mn % p.num_m_blocks;
      const int n_blk = mn / p.num_m_blocks;
      const int buf = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      const int row0 = ((m_blk * MC + static_cast<int>(pr)) * CG + static_cast<int>(rank)) * BM;
      const long long row = static_cast<long long>(row0) + row_in_tile;
      const bool row_ok = row < p.M;
      const int n0 = n_blk * BN;
      float ln_mu = 0.f, ln_rstd = 1.f;
      if (p.ln_stats != nullptr && row_ok) {  // finalise this row's LayerNorm statistics
        const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + row * p.ln_slabs;
        float s1 = 0.f, s2 = 0.f;
        if ((p.ln_slabs & 1) == 0 && p.ln_slabs <= 32) {
          // all loads in flight at once (a serial L2 round trip per slab would cost microseconds)
          const float4* st4 = reinterpret_cast<const float4*>(st);
          const int n4 = p.ln_slabs >> 1;
          float4 t4[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            t4[i] = (i < n4) ? __ldg(st4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            s1 += t4[i].x + t4[i].z;
            s2 += t4[i].y + t4[i].w;
          }
        } else {
          for (int i = 0; i < p.ln_slabs; ++i) {
            const float2 t2 = __ldg(st + i);
            s1 += t2.x;
            s2 += t2.y;
          }
        }
        ln_mu = s1 * p.ln_inv_c;
        ln_rstd = rsqrtf(fmaxf(s2 * p.ln_inv_c - ln_mu * ln_mu, 0.f) + p.ln_eps);
      }
      // tile bias (and, with a folded LayerNorm, the column sums of W') staged in smem; columns
      // beyond N read as 0.  Plain: bias double buffered by tile parity.  LN folded: [bias | colsum]
      // single buffered, released by the barrier at the end of the tile.
      const bool ln_in = p.ln_stats != nullptr;
      float* s_stage = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);
      const float* s_bias = s_stage + (ln_in ? 0 : buf * BN);
      const float* s_cs = s_stage + BN;
      if (p.bias != nullptr || ln_in) {
        for (int e = (warp - 4) * 32 + lane; e < BN; e += kEpiWarps * 32) {
          const bool ok = n0 + e < p.N;
          s_stage[(ln_in ? 0 : buf * BN) + e] = (ok && p.bias != nullptr) ? __ldg(p.bias + n0 + e) : 0.f;
          if (ln_in) s_stage[BN + e] = ok ? __ldg(p.ln_colsum + n0 + e) : 0.f;
        }
        named_bar_sync(5, kEpiWarps * 32);
      }
      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();
      if (warp == 4 && lane == 0) {
        if (t == 0) CD360_TRACE(8);
        if (tile + num_units >= num_tiles) CD360_TRACE(9);
      }
      const bool trace_me = warp == 4 && lane == 0 && tile + num_units >= num_tiles;
      (void)trace_me;
      CD360_TRACE_CLK(trace_me, 16);
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
      const float* rb = nullptr;
      if (p.row_bias != nullptr && row_ok) rb = p.row_bias + (row / p.rows_per_group) * p.ld_row_bias;

      if (p.epi_tma) {
        // ---------------- slab path: registers -> swizzled smem -> TMA store ----------------
        for (int s = 0; s < n_slabs; ++s) {
          float v[64];
          if (!p.geglu) {
            const int c0 = (half * SLABS + s) * 2;  // first 32-col chunk of this slab
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(taddr + c0 * 32, r0);
            tmem_ld_32x32b_x32(taddr + (c0 + 1) * 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float vv[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) vv[j] = __uint_as_float(h == 0 ? r0[j] : r1[j]);
              const int col0 = n0 + (c0 + h) * 32;
              const int nvalid = max(0, min(32, p.N - col0));
              if (nvalid > 0) {
                if (ln_in) {
                  const float4* cs4 = reinterpret_cast<const float4*>(s_cs + (c0 + h) * 32);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 c = cs4[j];
                    vv[4 * j + 0] = fmaf(-ln_mu, c.x, vv[4 * j + 0]) * ln_rstd;
                    vv[4 * j + 1] = fmaf(-ln_mu, c.y, vv[4 * j + 1]) * ln_rstd;
                    vv[4 * j + 2] = fmaf(-ln_mu, c.z, vv[4 * j + 2]) * ln_rstd;
                    vv[4 * j + 3] = fmaf(-ln_mu, c.w, vv[4 * j + 3]) * ln_rstd;
                  }
                }
                if (p.bias != nullptr) add_bias32_smem(vv, s_bias + (c0 + h) * 32);
                if (rb != nullptr) load_bias32(vv, rb, col0, nvalid);
              }
              if (p.act != CD360_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 32; ++j) vv[j] = apply_act(vv[j], p.act);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) v[h * 32 + j] = vv[j];
            }
          } else {
            // x columns [c*32..] and their gates [BN/2 + c*32..] of the pre-interleaved W tile
            const int c0 = (half * n_slabs + s) * 2;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t rx[32], rg[32];
              tmem_ld_32x32b_x32(taddr + (c0 + h) * 32, rx);
              tmem_ld_32x32b_x32(taddr + BN / 2 + (c0 + h) * 32, rg);
              tmem_ld_wait();
              float xv[32], gv[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                xv[j] = __uint_as_float(rx[j]);
                gv[j] = __uint_as_float(rg[j]);
              }
              if (ln_in) {
                const float4* cx4 = reinterpret_cast<const float4*>(s_cs + (c0 + h) * 32);
                const float4* cg4 = reinterpret_cast<const float4*>(s_cs + BN / 2 + (c0 + h) * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 a = cx4[j], g = cg4[j];
                  xv[4 * j + 0] = fmaf(-ln_mu, a.x, xv[4 * j + 0]) * ln_rstd;
                  xv[4 * j + 1] = fmaf(-ln_mu, a.y, xv[4 * j + 1]) * ln_rstd;
                  xv[4 * j + 2] = fmaf(-ln_mu, a.z, xv[4 * j + 2]) * ln_rstd;
                  xv[4 * j + 3] = fmaf(-ln_mu, a.w, xv[4 * j + 3]) * ln_rstd;
                  gv[4 * j + 0] = fmaf(-ln_mu, g.x, gv[4 * j + 0]) * ln_rstd;
                  gv[4 * j + 1] = fmaf(-ln_mu, g.y, gv[4 * j + 1]) * ln_rstd;
                  gv[4 * j + 2] = fmaf(-ln_mu, g.z, gv[4 * j + 2]) * ln_rstd;
                  gv[4 * j + 3] = fmaf(-ln_mu, g.w, gv[4 * j + 3]) * ln_rstd;
                }
              }
              if (p.bias != nullptr) {
                add_bias32_smem(xv, s_bias + (c0 + h) * 32);
                add_bias32_smem(gv, s_bias + BN / 2 + (c0 + h) * 32);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) v[h * 32 + j] = xv[j] * gelu_erf_f(gv[j]);
            }
          }
          CD360_TRACE_CLK(trace_me && s < 2, 17 + 5 * s);
          if (s == n_slabs - 1) {  // all TMEM reads of this tile done: hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2 && !leader) mbar_arrive_remote(&tmem_empty[buf], leader_rank);
              else mbar_arrive(&tmem_empty[buf]);
            }
          }
          uint8_t* sbuf = wbuf + (slab_cnt & 1u) * WSLAB_BYTES;
          uint8_t* brow = sbuf + lane * 128;
          if (use_res) {
            mbar_wait(&wres[slab_cnt & 1u], (slab_cnt >> 1) & 1u);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 u = *reinterpret_cast<const uint4*>(brow + ((c ^ sw) << 4));
              float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
                     f3 = unpack_bf16x2(u.w);
              v[8 * c + 0] += f0.x; v[8 * c + 1] += f0.y; v[8 * c + 2] += f1.x;
              v[8 * c + 3] += f1.y; v[8 * c + 4] += f2.x; v[8 * c + 5] += f2.y;
              v[8 * c + 6] += f3.x; v[8 * c + 7] += f3.y;
            }
          } else {
            // the store issued from this buffer two slabs ago must have finished reading it
            // (bulk groups are per thread: elect.sync picks the same lane for the same full mask)
            if (elect_one_sync()) tma_store_wait_read_1();
            __syncwarp();
          }
          CD360_TRACE_CLK(trace_me && s < 2, 18 + 5 * s);
          float so1 = 0.f, so2 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]);
            u.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
            u.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
            u.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
            *reinterpret_cast<uint4*>(brow + ((c ^ sw) << 4)) = u;
            if (p.stats_out != nullptr) {  // moments of the ROUNDED values a LayerNorm would read
              const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
                           f3 = unpack_bf16x2(u.w);
              so1 += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
              so2 = fmaf(f0.x, f0.x, so2); so2 = fmaf(f0.y, f0.y, so2);
              so2 = fmaf(f1.x, f1.x, so2); so2 = fmaf(f1.y, f1.y, so2);
              so2 = fmaf(f2.x, f2.x, so2); so2 = fmaf(f2.y, f2.y, so2);
              so2 = fmaf(f3.x, f3.x, so2); so2 = fmaf(f3.y, f3.y, so2);
            }
          }
          if (p.stats_out != nullptr && row_ok && slab_col(n_blk, s) < p.N_out)
            reinterpret_cast<float2*>(p.stats_out)[row * p.stats_slabs + slab_col(n_blk, s) / 64] =
                make_float2(so1, so2);
          fence_proxy_async_smem();
          __syncwarp();  // the warp's 32 rows are complete in smem
          CD360_TRACE_CLK(trace_me && s < 2, 20 + 5 * s);
          // coordinates of the next residual slab (this or the next tile), computed in uniform code
          int nxt_t = tile, nxt_s = s + 1;
          if (nxt_s == n_slabs) { nxt_t = tile + num_units; nxt_s = 0; }
          const bool nxt_ok = use_res && nxt_t < num_tiles;
          const int nxt_m = nxt_t % p.num_m_blocks, nxt_n = nxt_t / p.num_m_blocks;
          const int nxt_col = slab_col(nxt_n, nxt_s);
          const int nxt_row = ((nxt_m * MC + static_cast<int>(pr)) * CG + static_cast<int>(rank)) * BM + q * 32;
          const int out_col = slab_col(n_blk, s), out_row = row0 + q * 32;
          uint64_t* nb = &wres[(slab_cnt + 1u) & 1u];
          uint8_t* nbuf = wbuf + ((slab_cnt + 1u) & 1u) * WSLAB_BYTES;
          if (elect_one_sync()) {
            tma_store_2d(&tmOut, sbuf, out_col, out_row);
            tma_store_commit();
            if (nxt_ok) {  // prefetch the next residual slab into the other buffer
              tma_store_wait_read_1();  // the previous slab's store has released that buffer
              mbar_arrive_expect_tx(nb, WSLAB_BYTES);
              tma_load_2d(nbuf, &tmRes, nb, nxt_col, nxt_row);
            }
          }
          __syncwarp();
          ++slab_cnt;
          CD360_TRACE_CLK(trace_me && s < 2, 21 + 5 * s);
        }
      } else {
        // ---------------- direct path ----------------
        if (!p.geglu) {
          constexpr int CH = BN / 64;  // 32-column chunks per half
#pragma unroll 1
          for (int ci = 0; ci < CH; ++ci) {
            const int c = half * CH + ci;
            const int col0 = n0 + c * 32;
            if (col0 >= p.N) break;  // warp-uniform
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + c * 32, r);
            tmem_ld_wait();
            if (row_ok) {
              const int nvalid = min(32, p.N - col0);
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
              if (p.bias != nullptr) add_bias32_smem(v, s_bias + c * 32);
              if (rb != nullptr) load_bias32(v, rb, col0, nvalid);
              if (p.act != CD360_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
              }
              store_chunk32(v, p, row, col0, nvalid, ks);
            }
          }
        } else {
          constexpr int CH = BN / 128;  // (x, gate) chunk pairs per half
#pragma unroll 1
          for (int ci = 0; ci < (CH > 0 ? CH : 1); ++ci) {
            const int c = half * CH + ci;
            if (CH == 0 && half == 1) break;
            uint32_t rx[32], rg[32];
            tmem_ld_32x32b_x32(taddr + c * 32, rx);
            tmem_ld_32x32b_x32(taddr + BN / 2 + c * 32, rg);
            tmem_ld_wait();
            if (row_ok) {
              float v[32], g[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[j] = __uint_as_float(rx[j]);
                g[j] = __uint_as_float(rg[j]);
              }
              if (p.bias != nullptr) {
                add_bias32_smem(v, s_bias + c * 32);
                add_bias32_smem(g, s_bias + BN / 2 + c * 32);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = v[j] * gelu_erf_f(g[j]);
              store_chunk32(v, p, row, n_blk * (BN / 2) + c * 32, 32);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2 && !leader) mbar_arrive_remote(&tmem_empty[buf], leader_rank);
          else mbar_arrive(&tmem_empty[buf]);
        }
      }
      if (ln_in) named_bar_sync(6, kEpiWarps * 32);  // [bias | colsum] may be overwritten
    }
    if (warp == 4 && lane == 0) CD360_TRACE(10);
    // smem must outlive the bulk stores' READS; the writes are complete at grid completion
    __syncwarp();
    if (p.epi_tma && elect_one_sync()) tma_store_wait_read();
    if (warp == 4 && lane == 0) CD360_TRACE(11);
    CD360_TRACE_CLK(warp == 4 && lane == 0, 27);
  }

  // ---- teardown ----
  tc_fence_before();
  if (CG * MC > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) CD360_TRACE(12);
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// bf16 tensor map of rank `rank`; dims/strides innermost first; strides in BYTES for dims 1..rank-1
int encode_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool l2_256) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return CD360_ERR_LAUNCH;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B,
                  l2_256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CD360_OK : CD360_ERR_LAUNCH;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsB200;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = kNumSMsB200;
  }
  return n;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

template <int BN, int STAGES, int CG, int MC>
static int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                       const CUtensorMap& o, const CUtensorMap& r, const GemmKParams& p,
                       int max_ctas, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, CG>;
  static bool attr_set = false;
  auto kern = gemm_bf16_tcgen05_kernel<BN, STAGES, CG, MC>;
  constexpr int CL = CG * MC;  // CTAs per cluster
  static int max_units = 0;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) !=
        cudaSuccess)
      return CD360_ERR_LAUNCH;
    max_units = num_sms() / CL;
    if (CL > 2) {  // clusters of 4 may not tile every GPC: ask the driver how many fit at once
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(num_sms() / CL * CL);
      cfg.blockDim = dim3(kGemmThreads);
      cfg.dynamicSmemBytes = L::DYN_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CL;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters > 0 &&
          nclusters < max_units)
        max_units = nclusters;
    }
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks * p.ksplit;
  int units = max_units;
  if (max_ctas > 0 && max_ctas / CL >= 1 && max_ctas / CL < units) units = max_ctas / CL;
  if (tiles < units) units = tiles;
  if (launch_ex(kern, dim3(units * CL), dim3(kGemmThreads), L::DYN_BYTES, stream, CL, a0, a1, b, o,
                r, p) != cudaSuccess)
    return CD360_ERR_LAUNCH;
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

// tile configuration: block_n = 128 selects the single-CTA 128x128 kernel, 256 / 512 the
// 256x256 CTA-pair (cta_group::2) kernel, 1024 the pair kernel in clusters of two pairs that share
// W through TMA multicast; 0 = heuristic (pair whenever M spans two CTAs, multicast whenever the
// number of 256-row blocks is even).  Returns 128, 512 or 1024.
static int pick_config(int M, int N, int K, int geglu, int requested) {
  static int pair_ok = -1;
  if (pair_ok < 0) {
    const char* e = getenv("CD360_GEMM_PAIR");
    pair_ok = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  static int mc_ok = -1;
  if (mc_ok < 0) {
    const char* e = getenv("CD360_GEMM_MULTICAST");
    // measured on B200 (profiles/README_r01.md): no faster for one-wave shapes, slower for
    // multi-wave ones (fewer co-resident clusters of four; L2 already merges the two pairs' reads)
    mc_ok = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (requested == 128) return 128;
  if (geglu && (N % 256) != 0) return 128;
  if (requested == 256 || requested == 512) return 512;
  if (requested == 1024) return 1024;
  if (N <= 128 || M <= 128 || !pair_ok) return (geglu && (N % 256) == 0) ? 512 : 128;
  {
    // small problems (training step at batch 1: M = 256 ... 1024 rows): when 128x128 tiles fill at most
    // one wave of single CTAs, twice as many CTAs pull the operands and the epilogue tail of each is half
    // as long as a CTA pair's 128x256 share (CD360_GEMM_SMALL=0 keeps the pair tiles, for A/B runs)
    static int small_ok = -1;
    if (small_ok < 0) {
      const char* e = getenv("CD360_GEMM_SMALL");
      small_ok = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const long long single_tiles =
        static_cast<long long>((M + BM - 1) / BM) * static_cast<long long>((N + 127) / 128);
    // up to ONE wave of single CTAs.  Training step, A/B in one session: threshold num_sms / 2: 45.19 ms,
    // num_sms: 44.32 ms, num_sms only for K < 2560: 44.88 ms, 2 x num_sms: slower again.
    // (CD360_GEMM_SMALL_MAX=<tiles>, CD360_GEMM_SMALL_MAXK=<K>: A/B runs of the thresholds)
    static int small_max = -1, small_maxk = -1;
    if (small_max < 0) {
      const char* e = getenv("CD360_GEMM_SMALL_MAX");
      small_max = (e != nullptr && atoi(e) > 0) ? atoi(e) : num_sms();
      const char* k = getenv("CD360_GEMM_SMALL_MAXK");
      small_maxk = (k != nullptr && atoi(k) > 0) ? atoi(k) : (1 << 30);
    }
    if (small_ok && !geglu &&
        (single_tiles <= num_sms() / 2 || (single_tiles <= small_max && K < small_maxk)))
      return 128;
  }
  const int pair_blocks = (M + 2 * BM - 1) / (2 * BM);
  return (mc_ok && pair_blocks >= 2 && (pair_blocks & 1) == 0) ? 1024 : 512;
}

}  // namespace cd360

using namespace cd360;

extern "C" int cd360_geglu_pack_block(int32_t n_total) {
  if (n_total % 256 == 0) return 128;
  if (n_total % 128 == 0) return 64;
  return CD360_ERR_SHAPE;
}

extern "C" int cd360_gemm_bf16(const cd360_gemm_args* a, cd360_stream_t stream_) {
  if (a == nullptr || a->a0 == nullptr || a->w == nullptr || a->out == nullptr)
    return CD360_ERR_NULL;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (a->M <= 0 || a->N <= 0) return CD360_ERR_SHAPE;
  const int cfgsel = pick_config(a->M, a->N, a->conv ? 9 * a->C : a->k0 + a->k1, a->geglu, a->block_n);
  const int CGsel = cfgsel >= 512 ? 2 : 1;
  const int MCsel = cfgsel == 1024 ? 2 : 1;
  const int BN = cfgsel >= 512 ? 256 : 128;
  if (a->geglu && (a->N % BN != 0 || (a->N & 1))) return CD360_ERR_SHAPE;
  if (a->geglu && a->act != CD360_ACT_NONE) return CD360_ERR_UNSUPPORTED;
  if (a->tn) {  // contraction over the ROWS of a0 [K, M] and w [K, N]: plain / split-K epilogues only
    if (a->conv || a->k1 != 0 || a->geglu || a->ln_stats || a->stats_out || a->row_bias || cfgsel == 1024)
      return CD360_ERR_UNSUPPORTED;
    if ((a->M & 7) || (a->N & 7) || (a->lda0 & 7) || a->lda0 < a->M || (a->ldw & 7) || a->ldw < a->N)
      return CD360_ERR_ALIGN;
  }

  GemmKParams p{};
  p.M = a->M;
  p.N = a->N;
  p.N_out = a->geglu ? a->N / 2 : a->N;
  p.num_m_blocks = (a->M + BM * CGsel * MCsel - 1) / (BM * CGsel * MCsel);  // per cluster step
  p.num_n_blocks = (a->N + BN - 1) / BN;
  p.bias = a->bias;
  p.row_bias = a->row_bias;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.ld_row_bias = a->ld_row_bias > 0 ? a->ld_row_bias : a->N;
  if (a->row_bias && (p.ld_row_bias & 3)) return CD360_ERR_ALIGN;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.ldr = a->ldr;
  p.out = a->out;
  p.ldo = a->ldo;
  p.out_fp32 = a->out_fp32;
  p.act = a->act;
  p.geglu = a->geglu;
  if (a->row_bias != nullptr && a->geglu) return CD360_ERR_UNSUPPORTED;
  {
    static int dyn = -1;
    if (dyn < 0) {
      const char* e = getenv("CD360_GEMM_DYNN");
      dyn = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    p.dyn_n = dyn;
  }
  p.tn = a->tn ? 1 : 0;
  p.ksplit = 1;
  if (a->k_splits > 1) {
    // partial tiles go to fp32 slices of `out`; everything an epilogue would apply (bias, residual,
    // activation) belongs to cd360_splitk_finish
    if (!a->out_fp32 || a->bias || a->row_bias || a->residual || a->geglu || a->ln_stats ||
        a->stats_out || a->act != CD360_ACT_NONE || a->conv)
      return CD360_ERR_UNSUPPORTED;
    if (a->split_stride < static_cast<int64_t>(a->M) * a->ldo || (a->split_stride & 3))
      return CD360_ERR_SHAPE;
    p.split_stride = a->split_stride;
  }
  if ((reinterpret_cast<uintptr_t>(a->a0) & 15) || (reinterpret_cast<uintptr_t>(a->w) & 15) ||
      (reinterpret_cast<uintptr_t>(a->out) & 15) ||
      (a->residual && (reinterpret_cast<uintptr_t>(a->residual) & 15)) ||
      (a->bias && (reinterpret_cast<uintptr_t>(a->bias) & 15)) ||
      (a->row_bias && (reinterpret_cast<uintptr_t>(a->row_bias) & 15)))
    return CD360_ERR_ALIGN;
  if (a->residual && (a->ldr & 7)) return CD360_ERR_ALIGN;
  if (a->ldo < p.N_out) return CD360_ERR_SHAPE;
  if ((a->bias || a->row_bias) && (a->N & 3)) return CD360_ERR_ALIGN;
  // slab/TMA epilogue: bf16 output with 16-byte aligned rows (TMA clips the M / N tails)
  p.epi_tma = (!a->out_fp32 && (a->ldo & 7) == 0 && (p.N_out & 7) == 0) ? 1 : 0;
  if (a->geglu && BN == 128) p.epi_tma = 0;  // 64 output columns per tile: one slab, direct path
  {
    const char* e = getenv("CD360_GEMM_EPI_TMA");
    if (e != nullptr && e[0] == '0') p.epi_tma = 0;
  }
  if (a->ln_stats != nullptr || a->stats_out != nullptr) {
    if (!p.epi_tma || a->conv) return CD360_ERR_UNSUPPORTED;
    if (a->ln_stats != nullptr) {
      if (a->k1 != 0) return CD360_ERR_UNSUPPORTED;  // the folded LayerNorm spans one K segment
      if (a->ln_colsum == nullptr) return CD360_ERR_NULL;
      if (a->ln_slabs <= 0 || (reinterpret_cast<uintptr_t>(a->ln_stats) & 7) ||
          (reinterpret_cast<uintptr_t>(a->ln_colsum) & 15))
        return CD360_ERR_ALIGN;
      p.ln_stats = a->ln_stats;
      p.ln_slabs = a->ln_slabs;
      p.ln_inv_c = 1.0f / static_cast<float>(a->k0);
      p.ln_eps = a->ln_eps;
      p.ln_colsum = a->ln_colsum;
    }
    if (a->stats_out != nullptr) {
      if ((p.N_out % 64) != 0 || (reinterpret_cast<uintptr_t>(a->stats_out) & 7))
        return CD360_ERR_SHAPE;
      p.stats_out = a->stats_out;
      p.stats_slabs = p.N_out / 64;
    }
  }

  CUtensorMap tmA0, tmA1, tmB, tmOut, tmRes;
  int ktot;
  int rc;
  if (a->conv) {
    const int C = a->C, H = a->H, W = a->W, B = a->B;
    if (C <= 0 || (C % BK) != 0 || !is_pow2(H) || !is_pow2(W) || B <= 0) return CD360_ERR_SHAPE;
    if (static_cast<long long>(B) * H * W != a->M) return CD360_ERR_SHAPE;
    // W <= 128: a tile covers whole image rows; wider images (the VAE decoder, up to 1024):
    // 128-pixel segments of one row — powers of two, so a tile never straddles rows or images
    const int tw = W < BM ? W : BM;
    int th = BM / tw;
    if (th > H) th = H;
    const int tb = BM / (tw * th);
    p.conv = 1;
    p.cblocks = C / BK;
    p.H = H;
    p.W = W;
    p.tw = tw;
    p.th = th;
    p.tb = tb;
    ktot = 9 * C;
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W),
                        static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2,
                           static_cast<uint64_t>(H) * W * C * 2};
    uint32_t box[4] = {BK, static_cast<uint32_t>(tw), static_cast<uint32_t>(th),
                       static_cast<uint32_t>(tb)};
    rc = encode_tmap_bf16(&tmA0, a->a0, 4, dims, strides, box, false);
    if (rc != CD360_OK) return rc;
    tmA1 = tmA0;
  } else if (a->tn) {
    if (a->k0 <= 0) return CD360_ERR_SHAPE;
    p.kb0 = (a->k0 + BK - 1) / BK;   // K tail: TMA zero-fills the rows past k0
    p.kb1 = 0;
    ktot = a->k0;
    uint64_t dims[2] = {static_cast<uint64_t>(a->M), static_cast<uint64_t>(a->k0)};
    uint64_t strides[1] = {static_cast<uint64_t>(a->lda0) * 2};
    uint32_t box[2] = {64, BK};
    rc = encode_tmap_bf16(&tmA0, a->a0, 2, dims, strides, box, false);
    if (rc != CD360_OK) return rc;
    tmA1 = tmA0;
  } else {
    if (a->k0 <= 0 || (a->k0 & 7) || (a->lda0 & 7) || a->lda0 < a->k0) return CD360_ERR_ALIGN;
    if (a->k1 < 0) return CD360_ERR_SHAPE;
    if (a->k1 > 0) {
      if (a->a1 == nullptr) return CD360_ERR_NULL;
      if ((a->k0 % BK) != 0 || (a->k1 & 7) || (a->lda1 & 7) || a->lda1 < a->k1 ||
          (reinterpret_cast<uintptr_t>(a->a1) & 15))
        return CD360_ERR_ALIGN;
    }
    p.kb0 = (a->k0 + BK - 1) / BK;
    p.kb1 = (a->k1 + BK - 1) / BK;
    ktot = a->k0 + a->k1;
    {
      uint64_t dims[2] = {static_cast<uint64_t>(a->k0), static_cast<uint64_t>(a->M)};
      uint64_t strides[1] = {static_cast<uint64_t>(a->lda0) * 2};
      uint32_t box[2] = {BK, BM};
      rc = encode_tmap_bf16(&tmA0, a->a0, 2, dims, strides, box, false);
      if (rc != CD360_OK) return rc;
    }
    if (a->k1 > 0) {
      uint64_t dims[2] = {static_cast<uint64_t>(a->k1), static_cast<uint64_t>(a->M)};
      uint64_t strides[1] = {static_cast<uint64_t>(a->lda1) * 2};
      uint32_t box[2] = {BK, BM};
      rc = encode_tmap_bf16(&tmA1, a->a1, 2, dims, strides, box, false);
      if (rc != CD360_OK) return rc;
    } else {
      tmA1 = tmA0;
    }
  }
  {
    const int nkb_host = p.conv ? 9 * p.cblocks : (p.kb0 + p.kb1);
    p.kb_per = nkb_host;
    if (a->k_splits > 1 && nkb_host > 1) {
      const int want = a->k_splits < nkb_host ? a->k_splits : nkb_host;
      p.kb_per = (nkb_host + want - 1) / want;
      p.ksplit = (nkb_host + p.kb_per - 1) / p.kb_per;  // no empty split
    }
  }
  if (a->tn) {
    uint64_t dims[2] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->k0)};
    uint64_t strides[1] = {static_cast<uint64_t>(a->ldw) * 2};
    uint32_t box[2] = {64, BK};
    rc = encode_tmap_bf16(&tmB, a->w, 2, dims, strides, box, false);
    if (rc != CD360_OK) return rc;
  } else {
    uint64_t dims[2] = {static_cast<uint64_t>(ktot), static_cast<uint64_t>(a->N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ktot) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(BN / CGsel / MCsel)};
    rc = encode_tmap_bf16(&tmB, a->w, 2, dims, strides, box, true);
    if (rc != CD360_OK) return rc;
  }
  tmOut = tmB;
  tmRes = tmB;
  if (p.epi_tma) {
    uint64_t dims[2] = {static_cast<uint64_t>(p.N_out), static_cast<uint64_t>(a->M)};
    uint32_t box[2] = {64, 32};  // one epilogue warp's rows
    uint64_t so[1] = {static_cast<uint64_t>(a->ldo) * 2};
    rc = encode_tmap_bf16(&tmOut, a->out, 2, dims, so, box, false);
    if (rc != CD360_OK) return rc;
    if (a->residual != nullptr) {
      uint64_t sr[1] = {static_cast<uint64_t>(a->ldr) * 2};
      rc = encode_tmap_bf16(&tmRes, a->residual, 2, dims, sr, box, false);
      if (rc != CD360_OK) return rc;
    }
  }
  if (cfgsel == 1024)
    return launch_gemm<256, 5, 2, 2>(tmA0, tmA1, tmB, tmOut, tmRes, p, a->max_ctas, stream);
  if (cfgsel == 512)
    return launch_gemm<256, 5, 2, 1>(tmA0, tmA1, tmB, tmOut, tmRes, p, a->max_ctas, stream);
  return launch_gemm<128, 5, 1, 1>(tmA0, tmA1, tmB, tmOut, tmRes, p, a->max_ctas, stream);
}

CD360_TL_SETTER(gemm)
