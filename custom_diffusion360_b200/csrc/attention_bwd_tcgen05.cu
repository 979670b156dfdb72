// attention_bwd_tcgen05.cu — backward of softmax(Q K^T / 8) V (head dim 64) on tcgen05 / TMEM / TMA.
//
// The reference obtains these gradients from autograd through xformers.ops.memory_efficient_attention
// (sgm/modules/attention.py:406).  Flash-style recomputation, two kernels (no atomics between them):
//
//   attn_bwd_dq_tcgen05_kernel     one CTA per 128 queries of one (batch, head), two passes over the key
//                                  tiles (128 keys each):
//       pass 1  S = Q K_j^T -> running row maximum / sum -> lse2_i = log2 sum_j 2^(c s_ij); D_i = dO_i . O_i
//       pass 2  S = Q K_j^T, dP = dO V_j^T (two SS MMAs into TMEM) -> the row threads form
//               P = 2^(c S - lse2), dS = P o (dP - D) / 8 -> bf16 dS into TMEM -> dQ += dS K_j (TS MMA,
//               K_j consumed in place as an MN-major B operand)
//   attn_bwd_dkdv_tcgen05_kernel   one CTA per 128 keys, walks the query tiles:
//       S^T = K Q_i^T, dP^T = V dO_i^T -> P^T, dS^T (row = key; lse2 / D of the query tile broadcast from
//       shared memory) -> bf16 into TMEM -> dV += P^T dO_i, dK += dS^T Q_i (TS MMAs, dO_i / Q_i in place
//       as MN-major B operands).  With nsplit > 1 the query tiles are divided among nsplit CTAs per key
//       tile whose partial sums meet in fp32 accumulators (few keys x very many queries: reference_attn).
//
// Same building blocks as attention_tcgen05.cu: 4-D TMA maps over the [B, n, heads*64] projection
// buffers (no head permutes), whole-warp roles with one elected lane issuing (operands in uniform
// registers), P / dS handed to the second MMA through TMEM.  c = log2(e) / 8.
#include <cstdio>

#include "cd360_common.cuh"

namespace cd360 {

int encode_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool l2_256);

constexpr int BW_D = 64;
constexpr int BW_T = 128;                 // tile edge (queries / keys)
constexpr int BW_TILE_BYTES = BW_T * 128; // 128 rows x 64 bf16
constexpr int BW_STAGES = 3;
constexpr int BW_THREADS = 256;
constexpr int BW_S_FIX = 0;                                   // two resident tiles (Q, dO | K, V)
constexpr int BW_S_RING = 2 * BW_TILE_BYTES;                  // ring of (tile a, tile b)
constexpr int BW_S_VEC = BW_S_RING + BW_STAGES * 2 * BW_TILE_BYTES;   // float [2][2][128]: lse2 / D of a query tile
constexpr int BW_S_BAR = BW_S_VEC + 2 * 2 * 128 * 4;
constexpr int BW_SMEM_BYTES = BW_S_BAR + 128;
static_assert(BW_SMEM_BYTES <= 227 * 1024, "smem budget");
constexpr uint32_t BW_TMEM_COLS = 512;

struct BwdParams {
  const __nv_bfloat16 *o, *dout;   // dq kernel: rows for D_i
  long long ldo, lddo;
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  float *lse2, *dsum;              // [B, H, nq]
  float* kv_acc;                   // nsplit > 1: fp32 [2][B*nkv][H*64]
  int nq, nkv, heads, nsplit;
  float scale_log2, scale;
};

__device__ __forceinline__ float bw_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bw_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// =================================================================================================
// dQ (and the per-row statistics lse2 / D)
// =================================================================================================
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_dq_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                           const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                           const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW_S_BAR);
  uint64_t* fix_full = bars + 0;              // Q and dO tiles landed
  uint64_t* kv_full = bars + 1;               // [3]
  uint64_t* kv_empty = kv_full + BW_STAGES;   // [3]
  uint64_t* s_full = kv_empty + BW_STAGES;    // S (pass 2: and dP) complete in TMEM
  uint64_t* s_free = s_full + 1;              // the row threads hold them in registers
  uint64_t* ds_full = s_free + 1;             // dS written to TMEM
  uint64_t* dq_done = ds_full + 1;            // dQ += dS K retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dq_done + 1);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
  const int nt = (p.nkv + BW_T - 1) / BW_T;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(fix_full, 1);
    for (int s = 0; s < BW_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(ds_full, 4);
    mbar_init(dq_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, BW_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DS = 256, COL_DQ = 320;
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(fix_full, 2 * BW_TILE_BYTES);
      tma_load_4d(smem + BW_S_FIX, &tmQ, fix_full, 0, head, q_tile * BW_T, batch);
      tma_load_4d(smem + BW_S_FIX + BW_TILE_BYTES, &tmDO, fix_full, 0, head, q_tile * BW_T, batch);
    }
    __syncwarp();
    for (int n = 0; n < 2 * nt; ++n) {          // pass 1: K only; pass 2: K and V
      const int stage = n % BW_STAGES;
      const int j = n < nt ? n : n - nt;
      mbar_wait(&kv_empty[stage], ((n / BW_STAGES) & 1) ^ 1);
      uint8_t* sk = smem + BW_S_RING + stage * 2 * BW_TILE_BYTES;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&kv_full[stage], n < nt ? BW_TILE_BYTES : 2 * BW_TILE_BYTES);
        tma_load_4d(sk, &tmK, &kv_full[stage], 0, head, j * BW_T, batch);
        if (n >= nt) tma_load_4d(sk + BW_TILE_BYTES, &tmV, &kv_full[stage], 0, head, j * BW_T, batch);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, BW_T, false);
    constexpr uint32_t idesc_q = make_idesc_bf16(128, BW_D, true);    // B (K tile) MN-major
    const uint32_t q_addr = smem_u32(smem + BW_S_FIX);
    const uint32_t do_addr = q_addr + BW_TILE_BYTES;
    mbar_wait(fix_full, 0);
    for (int n = 0; n < 2 * nt; ++n) {
      const int stage = n % BW_STAGES;
      const bool pass2 = n >= nt;
      mbar_wait(&kv_full[stage], (n / BW_STAGES) & 1);
      if (n > 0) mbar_wait(s_free, (n - 1) & 1);   // S / dP of the previous tile are in registers
      tc_fence_after();
      const uint32_t k_addr = smem_u32(smem + BW_S_RING + stage * 2 * BW_TILE_BYTES);
      const uint32_t v_addr = k_addr + BW_TILE_BYTES;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < BW_D / 16; ++k)
          umma_bf16(tmem_base + COL_S, make_smem_desc_sw128(q_addr + k * 32), make_smem_desc_sw128(k_addr + k * 32),
                    idesc_s, k != 0 ? 1u : 0u);
        if (pass2) {
#pragma unroll
          for (int k = 0; k < BW_D / 16; ++k)
            umma_bf16(tmem_base + COL_DP, make_smem_desc_sw128(do_addr + k * 32),
                      make_smem_desc_sw128(v_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        if (!pass2) umma_commit(&kv_empty[stage]);
      }
      __syncwarp();
      if (pass2) {
        const int j = n - nt;
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < BW_T / 16; ++k)   // 16 keys = 8 TMEM columns of packed bf16 pairs
            umma_bf16_ts(tmem_base + COL_DQ, tmem_base + COL_DS + k * 8, make_smem_desc_sw128(k_addr + k * 16 * 128),
                         idesc_q, (j | k) != 0 ? 1u : 0u);
          umma_commit(dq_done);
          umma_commit(&kv_empty[stage]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ================================ row threads ================================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int qrow = q_tile * BW_T + row;
    const bool row_ok = qrow < p.nq;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // D_i = dO_i . O_i (fp32 over the bf16 rows)
    float dsum = 0.f;
    if (row_ok) {
      const uint4* po = reinterpret_cast<const uint4*>(p.o + (static_cast<long long>(batch) * p.nq + qrow) * p.ldo + head * BW_D);
      const uint4* pd = reinterpret_cast<const uint4*>(p.dout + (static_cast<long long>(batch) * p.nq + qrow) * p.lddo + head * BW_D);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 a = __ldg(po + c), b = __ldg(pd + c);
        const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
        const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
        dsum += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
      }
    }
    // ---- pass 1: lse2
    float m = -INFINITY, l = 0.f;   // m in log2 units (c * max s)
    for (int n = 0; n < nt; ++n) {
      const int valid = min(BW_T, p.nkv - n * BW_T);
      mbar_wait(s_full, n & 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[64];
        tmem_ld_32x32b_x32(t_lane + COL_S + hh * 64, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld_32x32b_x32(t_lane + COL_S + hh * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld_wait();
        if (hh == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (hh * 64 + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        const float m_new = fmaxf(m, mx * p.scale_log2);
        if (m_new > -INFINITY) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (hh * 64 + i < valid) acc += bw_ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_new));
          l = l * bw_ex2(m - m_new) + acc;
          m = m_new;
        }
      }
    }
    const float lse2 = m + bw_lg2(l);
    if (row_ok) {
      const long long idx = (static_cast<long long>(batch) * p.heads + head) * p.nq + qrow;
      p.lse2[idx] = lse2;
      p.dsum[idx] = dsum;
    }
    // ---- pass 2: dS -> TMEM
    for (int j = 0; j < nt; ++j) {
      const int n = nt + j;
      const int valid = min(BW_T, p.nkv - j * BW_T);
      mbar_wait(s_full, n & 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t rs[64], rp[64];
        tmem_ld_32x32b_x32(t_lane + COL_S + hh * 64, *reinterpret_cast<uint32_t(*)[32]>(&rs[0]));
        tmem_ld_32x32b_x32(t_lane + COL_S + hh * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&rs[32]));
        tmem_ld_32x32b_x32(t_lane + COL_DP + hh * 64, *reinterpret_cast<uint32_t(*)[32]>(&rp[0]));
        tmem_ld_32x32b_x32(t_lane + COL_DP + hh * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&rp[32]));
        tmem_ld_wait();
        if (hh == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          float d0 = 0.f, d1 = 0.f;
          if (hh * 64 + i < valid) {
            const float pr = bw_ex2(fmaf(__uint_as_float(rs[i]), p.scale_log2, -lse2));
            d0 = pr * (__uint_as_float(rp[i]) - dsum) * p.scale;
          }
          if (hh * 64 + i + 1 < valid) {
            const float pr = bw_ex2(fmaf(__uint_as_float(rs[i + 1]), p.scale_log2, -lse2));
            d1 = pr * (__uint_as_float(rp[i + 1]) - dsum) * p.scale;
          }
          pk[i >> 1] = pack_bf16x2(d0, d1);
        }
        if (hh == 0 && j > 0) {   // dQ += dS(j-1) K must have retired before dS is overwritten
          mbar_wait(dq_done, (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st_32x32b_x32(t_lane + COL_DS + hh * 32, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    // ---- epilogue: dQ -> bf16 -> global
    mbar_wait(dq_done, (nt - 1) & 1);
    tc_fence_after();
    __nv_bfloat16* dst = p.dq + (static_cast<long long>(batch) * p.nq + qrow) * p.lddq + head * BW_D;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t t[32];
      tmem_ld_32x32b_x32(t_lane + COL_DQ + c * 32, t);
      tmem_ld_wait();
      if (row_ok) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(t[8 * g + 0]), __uint_as_float(t[8 * g + 1]));
          u.y = pack_bf16x2(__uint_as_float(t[8 * g + 2]), __uint_as_float(t[8 * g + 3]));
          u.z = pack_bf16x2(__uint_as_float(t[8 * g + 4]), __uint_as_float(t[8 * g + 5]));
          u.w = pack_bf16x2(__uint_as_float(t[8 * g + 6]), __uint_as_float(t[8 * g + 7]));
          d4[g] = u;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BW_TMEM_COLS);
  }
}

// =================================================================================================
// dK / dV
// =================================================================================================
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_dkdv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                             const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                             const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW_S_BAR);
  uint64_t* fix_full = bars + 0;              // K and V tiles landed
  uint64_t* qd_full = bars + 1;               // [3] Q_i / dO_i
  uint64_t* qd_empty = qd_full + BW_STAGES;   // [3]
  uint64_t* s_full = qd_empty + BW_STAGES;    // S^T and dP^T complete
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;              // P^T and dS^T written to TMEM
  uint64_t* kv_done = p_full + 1;             // dV / dK MMAs of the tile retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(kv_done + 1);
  float* vec = reinterpret_cast<float*>(smem + BW_S_VEC);   // [parity][lse2 | D][128]

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int key_tile = static_cast<int>(blockIdx.x) / p.nsplit;
  const int split = static_cast<int>(blockIdx.x) - key_tile * p.nsplit;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q_tiles = (p.nq + BW_T - 1) / BW_T;
  const int per = (q_tiles + p.nsplit - 1) / p.nsplit;
  const int i0 = split * per;
  const int i1 = min(q_tiles, i0 + per);
  const int ni = max(i1 - i0, 0);

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(fix_full, 1);
    for (int s = 0; s < BW_STAGES; ++s) {
      mbar_init(&qd_full[s], 1);
      mbar_init(&qd_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(kv_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, BW_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_P = 256, COL_DS = 320, COL_DV = 384, COL_DK = 448;
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(fix_full, 2 * BW_TILE_BYTES);
      tma_load_4d(smem + BW_S_FIX, &tmK, fix_full, 0, head, key_tile * BW_T, batch);
      tma_load_4d(smem + BW_S_FIX + BW_TILE_BYTES, &tmV, fix_full, 0, head, key_tile * BW_T, batch);
    }
    __syncwarp();
    for (int n = 0; n < ni; ++n) {
      const int stage = n % BW_STAGES;
      mbar_wait(&qd_empty[stage], ((n / BW_STAGES) & 1) ^ 1);
      uint8_t* sq = smem + BW_S_RING + stage * 2 * BW_TILE_BYTES;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&qd_full[stage], 2 * BW_TILE_BYTES);
        tma_load_4d(sq, &tmQ, &qd_full[stage], 0, head, (i0 + n) * BW_T, batch);
        tma_load_4d(sq + BW_TILE_BYTES, &tmDO, &qd_full[stage], 0, head, (i0 + n) * BW_T, batch);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, BW_T, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, BW_D, true);    // B (dO_i / Q_i tile) MN-major
    const uint32_t k_addr = smem_u32(smem + BW_S_FIX);
    const uint32_t v_addr = k_addr + BW_TILE_BYTES;
    mbar_wait(fix_full, 0);
    for (int n = 0; n < ni; ++n) {
      const int stage = n % BW_STAGES;
      mbar_wait(&qd_full[stage], (n / BW_STAGES) & 1);
      if (n > 0) mbar_wait(s_free, (n - 1) & 1);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(smem + BW_S_RING + stage * 2 * BW_TILE_BYTES);
      const uint32_t do_addr = q_addr + BW_TILE_BYTES;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < BW_D / 16; ++k)     // S^T = K Q_i^T
          umma_bf16(tmem_base + COL_S, make_smem_desc_sw128(k_addr + k * 32), make_smem_desc_sw128(q_addr + k * 32),
                    idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BW_D / 16; ++k)     // dP^T = V dO_i^T
          umma_bf16(tmem_base + COL_DP, make_smem_desc_sw128(v_addr + k * 32), make_smem_desc_sw128(do_addr + k * 32),
                    idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, n & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < BW_T / 16; ++k) {   // 16 queries = 8 TMEM columns of packed bf16 pairs
          umma_bf16_ts(tmem_base + COL_DV, tmem_base + COL_P + k * 8, make_smem_desc_sw128(do_addr + k * 16 * 128),
                       idesc_o, (n | k) != 0 ? 1u : 0u);
          umma_bf16_ts(tmem_base + COL_DK, tmem_base + COL_DS + k * 8, make_smem_desc_sw128(q_addr + k * 16 * 128),
                       idesc_o, (n | k) != 0 ? 1u : 0u);
        }
        umma_commit(kv_done);
        umma_commit(&qd_empty[stage]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================ row threads (row = key) ================================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int key = key_tile * BW_T + row;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int tid = threadIdx.x - 128;   // 0..127
    for (int n = 0; n < ni; ++n) {
      // statistics of the query tile -> shared memory (query >= nq: lse2 = +inf, i.e. P = 0)
      float* vl = vec + (n & 1) * 256;
      {
        const int qi = (i0 + n) * BW_T + tid;
        const long long idx = (static_cast<long long>(batch) * p.heads + head) * p.nq + qi;
        vl[tid] = qi < p.nq ? p.lse2[idx] : INFINITY;
        vl[128 + tid] = qi < p.nq ? p.dsum[idx] : 0.f;
      }
      named_bar_sync(1, 128);
      mbar_wait(s_full, n & 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t rs[64], rp[64];
        tmem_ld_32x32b_x32(t_lane + COL_S + hh * 64, *reinterpret_cast<uint32_t(*)[32]>(&rs[0]));
        tmem_ld_32x32b_x32(t_lane + COL_S + hh * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&rs[32]));
        tmem_ld_32x32b_x32(t_lane + COL_DP + hh * 64, *reinterpret_cast<uint32_t(*)[32]>(&rp[0]));
        tmem_ld_32x32b_x32(t_lane + COL_DP + hh * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&rp[32]));
        tmem_ld_wait();
        if (hh == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
        uint32_t pp[32], pd[32];
        const float4* l4 = reinterpret_cast<const float4*>(vl + hh * 64);
        const float4* d4 = reinterpret_cast<const float4*>(vl + 128 + hh * 64);
#pragma unroll
        for (int i4 = 0; i4 < 16; ++i4) {
          const float4 ls = l4[i4], ds = d4[i4];
          const float lsv[4] = {ls.x, ls.y, ls.z, ls.w};
          const float dsv[4] = {ds.x, ds.y, ds.z, ds.w};
          float pr[4], dd[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i4 * 4 + e;
            pr[e] = bw_ex2(fmaf(__uint_as_float(rs[i]), p.scale_log2, -lsv[e]));
            dd[e] = pr[e] * (__uint_as_float(rp[i]) - dsv[e]) * p.scale;
          }
          pp[i4 * 2] = pack_bf16x2(pr[0], pr[1]);
          pp[i4 * 2 + 1] = pack_bf16x2(pr[2], pr[3]);
          pd[i4 * 2] = pack_bf16x2(dd[0], dd[1]);
          pd[i4 * 2 + 1] = pack_bf16x2(dd[2], dd[3]);
        }
        if (hh == 0 && n > 0) {   // the dV / dK MMAs of tile n-1 must have retired before P^T / dS^T change
          mbar_wait(kv_done, (n - 1) & 1);
          tc_fence_after();
        }
        tmem_st_32x32b_x32(t_lane + COL_P + hh * 32, pp);
        tmem_st_32x32b_x32(t_lane + COL_DS + hh * 32, pd);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue
    if (ni > 0) {
      mbar_wait(kv_done, (ni - 1) & 1);
      tc_fence_after();
    }
    const bool key_ok = key < p.nkv;
#pragma unroll
    for (int which = 0; which < 2; ++which) {   // 0: dV, 1: dK
      const uint32_t col = which == 0 ? COL_DV : COL_DK;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t t[32];
        if (ni > 0) {
          tmem_ld_32x32b_x32(t_lane + col + c * 32, t);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) t[i] = 0u;
        }
        if (!key_ok) continue;
        if (p.nsplit > 1) {
          float* acc = p.kv_acc + (which == 0 ? static_cast<long long>(gridDim.z) * p.nkv * p.heads * BW_D : 0LL) +
                       (static_cast<long long>(batch) * p.nkv + key) * (p.heads * BW_D) + head * BW_D + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(acc + i, __uint_as_float(t[i]));
        } else {
          __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
          const long long ld = which == 0 ? p.lddv : p.lddk;
          uint4* d4o = reinterpret_cast<uint4*>(base + (static_cast<long long>(batch) * p.nkv + key) * ld + head * BW_D + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(t[8 * g + 0]), __uint_as_float(t[8 * g + 1]));
            u.y = pack_bf16x2(__uint_as_float(t[8 * g + 2]), __uint_as_float(t[8 * g + 3]));
            u.z = pack_bf16x2(__uint_as_float(t[8 * g + 4]), __uint_as_float(t[8 * g + 5]));
            u.w = pack_bf16x2(__uint_as_float(t[8 * g + 6]), __uint_as_float(t[8 * g + 7]));
            d4o[g] = u;
          }
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BW_TMEM_COLS);
  }
}

static int bw_map(CUtensorMap* tm, const void* base, long long ld, int heads, int n, int batch) {
  // element (b, n, h, d) at ((b*n_total + n)*ld + h*64 + d); box = 128 rows of one head
  uint64_t dims[4] = {BW_D, static_cast<uint64_t>(heads), static_cast<uint64_t>(n), static_cast<uint64_t>(batch)};
  uint64_t strides[3] = {BW_D * 2, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(n) * static_cast<uint64_t>(ld) * 2};
  uint32_t box[4] = {BW_D, 1, BW_T, 1};
  return encode_tmap_bf16(tm, base, 4, dims, strides, box, true);
}

// Launch helpers used by cd360_attention_bwd_bf16 / cd360_attention_bwd_kv_split_bf16 (attention_bwd.cu).
// which: bit 0 = dQ (+ statistics), bit 1 = dK / dV.
int attention_bwd_tcgen05(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                          const void* o, long long ldo, const void* dout, long long lddo, void* dq, long long lddq,
                          void* dk, long long lddk, void* dv, long long lddv, float* lse, float* dsum, float* kv_acc,
                          int batch, int heads, int nq, int nkv, int nsplit, int which, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attn_bwd_dq_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES) !=
            cudaSuccess ||
        cudaFuncSetAttribute(attn_bwd_dkdv_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             BW_SMEM_BYTES) != cudaSuccess)
      return CD360_ERR_LAUNCH;
    attr_done = true;
  }
  CUtensorMap tq, tdo, tk, tv;
  int rc = bw_map(&tq, q, ldq, heads, nq, batch);
  if (rc != CD360_OK) return rc;
  rc = bw_map(&tdo, dout, lddo, heads, nq, batch);
  if (rc != CD360_OK) return rc;
  rc = bw_map(&tk, k, ldk, heads, nkv, batch);
  if (rc != CD360_OK) return rc;
  rc = bw_map(&tv, v, ldv, heads, nkv, batch);
  if (rc != CD360_OK) return rc;
  BwdParams p{};
  p.o = reinterpret_cast<const __nv_bfloat16*>(o);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.ldo = ldo;
  p.lddo = lddo;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq;
  p.lddk = lddk;
  p.lddv = lddv;
  p.lse2 = lse;
  p.dsum = dsum;
  p.kv_acc = kv_acc;
  p.nq = nq;
  p.nkv = nkv;
  p.heads = heads;
  p.nsplit = nsplit > 0 ? nsplit : 1;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  if (which & 1) {
    const dim3 gq((nq + BW_T - 1) / BW_T, heads, batch);
    if (launch_ex(attn_bwd_dq_tcgen05_kernel, gq, dim3(BW_THREADS), BW_SMEM_BYTES, stream, 1, tq, tdo, tk, tv, p) !=
        cudaSuccess)
      return CD360_ERR_LAUNCH;
  }
  if (which & 2) {
    const int key_tiles = (nkv + BW_T - 1) / BW_T;
    const int q_tiles = (nq + BW_T - 1) / BW_T;
    if (p.nsplit > q_tiles) p.nsplit = q_tiles;
    const dim3 gk(static_cast<unsigned>(key_tiles * p.nsplit), heads, batch);
    if (launch_ex(attn_bwd_dkdv_tcgen05_kernel, gk, dim3(BW_THREADS), BW_SMEM_BYTES, stream, 1, tq, tdo, tk, tv, p) !=
        cudaSuccess)
      return CD360_ERR_LAUNCH;
  }
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

}  // namespace cd360
