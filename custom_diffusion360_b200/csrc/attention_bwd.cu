// attention_bwd.cu — backward of softmax(Q K^T / 8) V, head dim 64 (training step).
//
// The reference gets these gradients from autograd through xformers.ops.memory_efficient_attention
// (sgm/modules/attention.py:406); the training step needs dQ for every attention on the path (the
// gradient flows through the frozen UNet towards the pose weights) and dK/dV for self-attention.
// Cross-attention K/V come from frozen projections of the text context, which is a constant of
// this path (the conditioner is outside SURVEY §8), so no dK/dV is produced there.
//
// Flash-style recomputation, three kernels over 64 x 64 tiles, all reading Q/K/V/O/dO in place from
// the [B, n, heads*64] projection buffers (row strides in elements, like the forward kernel):
//   attention_bwd_stats : per query row  lse = log sum_j exp(s_ij / 8),  D = sum_d dO_id O_id
//   attention_bwd_dq    : dQ_i  = sum_j dS_ij K_j           (one CTA per 64 queries)
//   attention_bwd_dkdv  : dK_j  = sum_i dS_ij^T Q_i,  dV_j = sum_i P_ij^T dO_i  (one CTA per 64 keys)
// with P = exp(s/8 - lse), dS = P o (dO V^T - D) / 8.  The tile products run on the warp-level
// tensor-core path (mma.sync via nvcuda::wmma, bf16 in / fp32 accumulate): correct and compact, and
// the first version of this path — the backward is ~2.5x the forward FLOPs at training sizes
// (64x64 latents, <= 1024 keys); a tcgen05 version in the style of attention_tcgen05.cu is the
// follow-up once the training step is measured.
#include <mma.h>

#include "cd360_common.cuh"

namespace cd360 {
namespace wm = nvcuda::wmma;

// attention_bwd_tcgen05.cu: the tcgen05 / TMEM kernels (default path).  which: bit 0 = dQ + statistics,
// bit 1 = dK / dV (nsplit > 1: partial sums into kv_acc).
int attention_bwd_tcgen05(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                          const void* o, long long ldo, const void* dout, long long lddo, void* dq, long long lddq,
                          void* dk, long long lddk, void* dv, long long lddv, float* lse, float* dsum, float* kv_acc,
                          int batch, int heads, int nq, int nkv, int nsplit, int which, cudaStream_t stream);
// 0: tcgen05 (default), 1: mma.sync kernels of this file (CD360_ATTBWD=mma), 2: first-generation wmma
// kernels (CD360_ATTBWD=wmma); read per call so tests can switch
static int attn_bwd_impl() {
  const char* e = getenv("CD360_ATTBWD");
  if (e == nullptr) return 0;
  if (e[0] == 'm') return 1;
  if (e[0] == 'w') return 2;
  return 0;
}

constexpr int AB_T = 64;     // tile edge: queries / keys / head dim
constexpr int AB_LD = 72;    // bf16 smem row stride (elements): 144 B rows, fragment loads stay 32 B aligned
constexpr int AB_LDF = 68;   // fp32 smem row stride
constexpr int AB_THREADS = 128;
constexpr int AB_TILE_B = AB_T * AB_LD * 2;   // bytes of one bf16 tile
constexpr int AB_TILE_F = AB_T * AB_LDF * 4;  // bytes of one fp32 tile

typedef wm::fragment<wm::matrix_a, 16, 16, 16, __nv_bfloat16, wm::row_major> FragA;
typedef wm::fragment<wm::matrix_b, 16, 16, 16, __nv_bfloat16, wm::col_major> FragBc;
typedef wm::fragment<wm::matrix_b, 16, 16, 16, __nv_bfloat16, wm::row_major> FragBr;
typedef wm::fragment<wm::accumulator, 16, 16, 16, float> FragC;

struct AttnBwdParams {
  const __nv_bfloat16 *q, *k, *v, *o, *dout;
  long long ldq, ldk, ldv, ldo, lddo;
  float *lse, *dsum;  // [B, H, nq]
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  int nq, nkv, heads;
  float scale;  // 1 / sqrt(64)
};

// 64 rows x 64 bf16 of one (batch, head) -> smem [64][AB_LD]; rows >= n are zero filled
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld,
                                          int batch, int n, int head, int row0) {
  for (int i = threadIdx.x; i < AB_T * 8; i += AB_THREADS) {
    const int r = i >> 3, ch = i & 7;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < n)
      u = __ldg(reinterpret_cast<const uint4*>(src + (static_cast<long long>(batch) * n + row0 + r) * ld +
                                               head * AB_T + ch * 8));
    *reinterpret_cast<uint4*>(dst + r * AB_LD + ch * 8) = u;
  }
}

// C[16 x 64] (rows w16 of A) = A[16 x 64] * B^T, B stored [64 rows][64 k] row-major (so B^T is
// col-major with ld = AB_LD); result to fp32 smem rows w16
__device__ __forceinline__ void mma_abt_to_smem(const __nv_bfloat16* a, const __nv_bfloat16* b,
                                                float* c, int w16) {
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    FragC acc;
    wm::fill_fragment(acc, 0.f);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      FragA fa;
      FragBc fb;
      wm::load_matrix_sync(fa, a + w16 * AB_LD + kk * 16, AB_LD);
      wm::load_matrix_sync(fb, b + nb * 16 * AB_LD + kk * 16, AB_LD);
      wm::mma_sync(acc, fa, fb, acc);
    }
    wm::store_matrix_sync(c + w16 * AB_LDF + nb * 16, acc, AB_LDF, wm::mem_row_major);
  }
}
// acc[4] (16 x 64) += A[16 x 64] * B, A bf16 smem rows w16 (row-major), B stored [64 k][64 n]
__device__ __forceinline__ void mma_ab_acc(FragC (&acc)[4], const __nv_bfloat16* a,
                                           const __nv_bfloat16* b, int w16) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    FragA fa;
    wm::load_matrix_sync(fa, a + w16 * AB_LD + kk * 16, AB_LD);
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      FragBr fb;
      wm::load_matrix_sync(fb, b + kk * 16 * AB_LD + nb * 16, AB_LD);
      wm::mma_sync(acc[nb], fa, fb, acc[nb]);
    }
  }
}
// acc[4] (16 x 64 fp32, via smem staging rows w16) -> bf16 global rows [row0 + w16, +16), cols of `head`
__device__ __forceinline__ void store_acc_bf16(FragC (&acc)[4], float* stage, __nv_bfloat16* dst,
                                               long long ld, int batch, int n, int head, int row0,
                                               int w16) {
#pragma unroll
  for (int nb = 0; nb < 4; ++nb)
    wm::store_matrix_sync(stage + w16 * AB_LDF + nb * 16, acc[nb], AB_LDF, wm::mem_row_major);
  __syncwarp();
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < 16 * 8; i += 32) {
    const int r = i >> 3, ch = i & 7;
    const int row = row0 + w16 + r;
    if (row < n) {
      const float* s = stage + (w16 + r) * AB_LDF + ch * 8;
      uint4 u;
      u.x = pack_bf16x2(s[0], s[1]);
      u.y = pack_bf16x2(s[2], s[3]);
      u.z = pack_bf16x2(s[4], s[5]);
      u.w = pack_bf16x2(s[6], s[7]);
      *reinterpret_cast<uint4*>(dst + (static_cast<long long>(batch) * n + row) * ld + head * AB_T + ch * 8) = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// lse / D per query row.  grid (ceil(nq/64), heads, batch), 128 threads: warp w owns rows w*16..
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_stats_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_wait();  // inputs come from the preceding kernels of the stream
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem + AB_TILE_B);
  float* sS = reinterpret_cast<float*>(smem + 2 * AB_TILE_B);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * AB_T, head = blockIdx.y, batch = blockIdx.z;
  const int w16 = warp * 16;
  load_tile(sQ, p.q, p.ldq, batch, p.nq, head, row0);
  const int r = lane >> 1, half = lane & 1;  // two lanes per query row, 32 keys each
  float m = -INFINITY, l = 0.f;
  const int nt = (p.nkv + AB_T - 1) / AB_T;
  for (int j = 0; j < nt; ++j) {
    __syncthreads();  // previous tile consumed (and sQ visible on the first pass)
    load_tile(sK, p.k, p.ldk, batch, p.nkv, head, j * AB_T);
    __syncthreads();
    mma_abt_to_smem(sQ, sK, sS, w16);
    __syncwarp();
    const float* srow = sS + (w16 + r) * AB_LDF + half * 32;
    const int valid = p.nkv - j * AB_T - half * 32;  // keys of this half that exist
    float mx = -INFINITY;
#pragma unroll 8
    for (int c = 0; c < 32; ++c)
      if (c < valid) mx = fmaxf(mx, srow[c] * p.scale);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    const float m_new = fmaxf(m, mx);  // finite: every tile holds >= 1 valid key in half 0
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < 32; ++c)
      if (c < valid) s += __expf(srow[c] * p.scale - m_new);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    l = l * __expf(m - m_new) + s;
    m = m_new;
  }
  // D = sum_d dO O over this row's 64 channels (32 per lane)
  const int row = row0 + w16 + r;
  float dsum = 0.f;
  if (row < p.nq) {
    const __nv_bfloat16* po = p.o + (static_cast<long long>(batch) * p.nq + row) * p.ldo + head * AB_T + half * 32;
    const __nv_bfloat16* pd = p.dout + (static_cast<long long>(batch) * p.nq + row) * p.lddo + head * AB_T + half * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 uo = __ldg(reinterpret_cast<const uint4*>(po + c * 8));
      const uint4 ud = __ldg(reinterpret_cast<const uint4*>(pd + c * 8));
      const uint32_t ao[4] = {uo.x, uo.y, uo.z, uo.w}, ad[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fo = unpack_bf16x2(ao[e]), fd = unpack_bf16x2(ad[e]);
        dsum = fmaf(fo.x, fd.x, dsum);
        dsum = fmaf(fo.y, fd.y, dsum);
      }
    }
  }
  dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
  if (row < p.nq && half == 0) {
    const long long idx = (static_cast<long long>(batch) * p.heads + head) * p.nq + row;
    p.lse[idx] = m + __logf(l);
    p.dsum[idx] = dsum;
  }
}

// dS (bf16, rows w16) from the fp32 S / dP tiles of this warp: rows = queries, columns = keys
__device__ __forceinline__ void ds_rows_q(const float* sS, const float* sP, __nv_bfloat16* sDS,
                                          int w16, float lse, float dsum, bool row_ok, int kv_valid,
                                          float scale) {
  const int lane = threadIdx.x & 31;
  const int r = lane >> 1, half = lane & 1;
  const float* s = sS + (w16 + r) * AB_LDF + half * 32;
  const float* dp = sP + (w16 + r) * AB_LDF + half * 32;
  __nv_bfloat16* o = sDS + (w16 + r) * AB_LD + half * 32;
  const int valid = kv_valid - half * 32;
#pragma unroll 8
  for (int c = 0; c < 32; c += 2) {
    float d0 = 0.f, d1 = 0.f;
    if (row_ok && c < valid) d0 = __expf(s[c] * scale - lse) * (dp[c] - dsum) * scale;
    if (row_ok && c + 1 < valid) d1 = __expf(s[c + 1] * scale - lse) * (dp[c + 1] - dsum) * scale;
    *reinterpret_cast<uint32_t*>(o + c) = pack_bf16x2(d0, d1);
  }
}

// ---------------------------------------------------------------------------------------------
// dQ.  grid (ceil(nq/64), heads, batch)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_dq_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_wait();  // inputs come from the preceding kernels of the stream
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sDO = reinterpret_cast<__nv_bfloat16*>(smem + AB_TILE_B);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem + 2 * AB_TILE_B);
  __nv_bfloat16* sV = reinterpret_cast<__nv_bfloat16*>(smem + 3 * AB_TILE_B);
  __nv_bfloat16* sDS = reinterpret_cast<__nv_bfloat16*>(smem + 4 * AB_TILE_B);
  float* sS = reinterpret_cast<float*>(smem + 5 * AB_TILE_B);
  float* sP = reinterpret_cast<float*>(smem + 5 * AB_TILE_B + AB_TILE_F);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * AB_T, head = blockIdx.y, batch = blockIdx.z;
  const int w16 = warp * 16;
  load_tile(sQ, p.q, p.ldq, batch, p.nq, head, row0);
  load_tile(sDO, p.dout, p.lddo, batch, p.nq, head, row0);
  const int row = row0 + w16 + (lane >> 1);
  const bool row_ok = row < p.nq;
  float lse = 0.f, dsum = 0.f;
  if (row_ok) {
    const long long idx = (static_cast<long long>(batch) * p.heads + head) * p.nq + row;
    lse = p.lse[idx];
    dsum = p.dsum[idx];
  }
  FragC acc[4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) wm::fill_fragment(acc[nb], 0.f);
  const int nt = (p.nkv + AB_T - 1) / AB_T;
  for (int j = 0; j < nt; ++j) {
    __syncthreads();
    load_tile(sK, p.k, p.ldk, batch, p.nkv, head, j * AB_T);
    load_tile(sV, p.v, p.ldv, batch, p.nkv, head, j * AB_T);
    __syncthreads();
    mma_abt_to_smem(sQ, sK, sS, w16);    // S  = Q K^T
    mma_abt_to_smem(sDO, sV, sP, w16);   // dP = dO V^T
    __syncwarp();
    ds_rows_q(sS, sP, sDS, w16, lse, dsum, row_ok, p.nkv - j * AB_T, p.scale);
    __syncwarp();
    mma_ab_acc(acc, sDS, sK, w16);       // dQ += dS K
  }
  store_acc_bf16(acc, sS, p.dq, p.lddq, batch, p.nq, head, row0, w16);
}

// ---------------------------------------------------------------------------------------------
// dK / dV.  grid (ceil(nkv/64), heads, batch): warp w owns keys w*16.. of the tile
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_dkdv_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_wait();  // inputs come from the preceding kernels of the stream
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sV = reinterpret_cast<__nv_bfloat16*>(smem + AB_TILE_B);
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem + 2 * AB_TILE_B);
  __nv_bfloat16* sDO = reinterpret_cast<__nv_bfloat16*>(smem + 3 * AB_TILE_B);
  __nv_bfloat16* sPb = reinterpret_cast<__nv_bfloat16*>(smem + 4 * AB_TILE_B);  // P^T  [keys][queries]
  __nv_bfloat16* sDS = reinterpret_cast<__nv_bfloat16*>(smem + 5 * AB_TILE_B);  // dS^T [keys][queries]
  float* sS = reinterpret_cast<float*>(smem + 6 * AB_TILE_B);
  float* sP = reinterpret_cast<float*>(smem + 6 * AB_TILE_B + AB_TILE_F);
  float* sLse = reinterpret_cast<float*>(smem + 6 * AB_TILE_B + 2 * AB_TILE_F);
  float* sD = sLse + AB_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int key0 = blockIdx.x * AB_T, head = blockIdx.y, batch = blockIdx.z;
  const int w16 = warp * 16;
  load_tile(sK, p.k, p.ldk, batch, p.nkv, head, key0);
  load_tile(sV, p.v, p.ldv, batch, p.nkv, head, key0);
  FragC accK[4], accV[4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    wm::fill_fragment(accK[nb], 0.f);
    wm::fill_fragment(accV[nb], 0.f);
  }
  const int r = lane >> 1, half = lane & 1;
  const bool key_ok = key0 + w16 + r < p.nkv;
  const int nt = (p.nq + AB_T - 1) / AB_T;
  for (int i = 0; i < nt; ++i) {
    __syncthreads();
    load_tile(sQ, p.q, p.ldq, batch, p.nq, head, i * AB_T);
    load_tile(sDO, p.dout, p.lddo, batch, p.nq, head, i * AB_T);
    if (threadIdx.x < AB_T) {
      const int qrow = i * AB_T + threadIdx.x;
      const long long idx = (static_cast<long long>(batch) * p.heads + head) * p.nq + qrow;
      sLse[threadIdx.x] = qrow < p.nq ? p.lse[idx] : 0.f;
      sD[threadIdx.x] = qrow < p.nq ? p.dsum[idx] : 0.f;
    }
    __syncthreads();
    mma_abt_to_smem(sK, sQ, sS, w16);    // S^T  = K Q^T   [keys][queries]
    mma_abt_to_smem(sV, sDO, sP, w16);   // dP^T = V dO^T
    __syncwarp();
    {
      const float* s = sS + (w16 + r) * AB_LDF + half * 32;
      const float* dp = sP + (w16 + r) * AB_LDF + half * 32;
      __nv_bfloat16* op = sPb + (w16 + r) * AB_LD + half * 32;
      __nv_bfloat16* od = sDS + (w16 + r) * AB_LD + half * 32;
      const int q_valid = p.nq - i * AB_T - half * 32;
#pragma unroll 8
      for (int c = 0; c < 32; c += 2) {
        float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
        if (key_ok && c < q_valid) {
          p0 = __expf(s[c] * p.scale - sLse[half * 32 + c]);
          d0 = p0 * (dp[c] - sD[half * 32 + c]) * p.scale;
        }
        if (key_ok && c + 1 < q_valid) {
          p1 = __expf(s[c + 1] * p.scale - sLse[half * 32 + c + 1]);
          d1 = p1 * (dp[c + 1] - sD[half * 32 + c + 1]) * p.scale;
        }
        *reinterpret_cast<uint32_t*>(op + c) = pack_bf16x2(p0, p1);
        *reinterpret_cast<uint32_t*>(od + c) = pack_bf16x2(d0, d1);
      }
    }
    __syncwarp();
    mma_ab_acc(accV, sPb, sDO, w16);   // dV += P^T dO
    mma_ab_acc(accK, sDS, sQ, w16);    // dK += dS^T Q
  }
  store_acc_bf16(accK, sS, p.dk, p.lddk, batch, p.nkv, head, key0, w16);
  __syncwarp();
  store_acc_bf16(accV, sS, p.dv, p.lddv, batch, p.nkv, head, key0, w16);
}


// =================================================================================================
// Second generation (default): the same two-pass algorithm on raw mma.sync.m16n8k16 with the score
// tiles kept in REGISTERS (FlashAttention-2 style).  Per warp: 16 rows x 64 columns of S / dP as 8
// accumulator tiles; the accumulator layout of two adjacent n8 tiles IS the A-operand layout of the
// next k16 step, so P / dS feed the second product without touching shared memory.  K-major B
// operands (K^T, V^T, Q^T, dO^T) are plain 32-bit shared loads (rows padded to 144 B: conflict
// free); row-major B operands (K, Q, dO as [k][n]) come through ldmatrix.trans.
//   attention_bwd_dq_mma   : pass 1 lse (online), D = rowsum(dO o O), pass 2 dQ; writes lse / D
//   attention_bwd_dkdv_mma : dK, dV for 64 keys per CTA, looping over the query tiles
// Measured on the training step's shapes (profiles/README_r01.md): 52 ms -> see there.
// =================================================================================================
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four transposed 8x8 b16 matrices: lane l supplies the address of row (l & 7) of matrix (l >> 3)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
// A fragments (16 rows w16.. x 64 k) of a row-major smem tile: frag[ks] covers k = ks*16 .. +15
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const __nv_bfloat16* tile, int w16) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const __nv_bfloat16* p0 = tile + (w16 + g) * AB_LD + ks * 16 + 2 * t;
    a[ks][0] = *reinterpret_cast<const uint32_t*>(p0);
    a[ks][1] = *reinterpret_cast<const uint32_t*>(p0 + 8 * AB_LD);
    a[ks][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
    a[ks][3] = *reinterpret_cast<const uint32_t*>(p0 + 8 * AB_LD + 8);
  }
}
// acc[nt] (16 x 64, 8 tiles of n8) = A (frags) * B^T with B stored row-major [n rows][64 k]
__device__ __forceinline__ void mma_a_bt(float (&acc)[8][4], const uint32_t (&a)[4][4],
                                         const __nv_bfloat16* b) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const __nv_bfloat16* pb = b + (nt * 8 + g) * AB_LD + ks * 16 + 2 * t;
      mma16816(acc[nt], a[ks], *reinterpret_cast<const uint32_t*>(pb),
               *reinterpret_cast<const uint32_t*>(pb + 8));
    }
  }
}
// out[nt] (16 x 64) += X (16 x 64 in accumulator layout, converted to bf16 A fragments) * B with B
// stored row-major [64 k rows][64 n]
__device__ __forceinline__ void mma_x_b(float (&out)[8][4], const float (&x)[8][4],
                                        const __nv_bfloat16* b) {
  const int lane = threadIdx.x & 31;
  const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t a[4];
    a[0] = pack_bf16x2(x[2 * kt][0], x[2 * kt][1]);
    a[1] = pack_bf16x2(x[2 * kt][2], x[2 * kt][3]);
    a[2] = pack_bf16x2(x[2 * kt + 1][0], x[2 * kt + 1][1]);
    a[3] = pack_bf16x2(x[2 * kt + 1][2], x[2 * kt + 1][3]);
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      uint32_t r[4];
      ldmatrix_x4_trans(r, b + (kt * 16 + (mi & 1) * 8 + rr) * AB_LD + dt * 16 + (mi >> 1) * 8);
      mma16816(out[2 * dt], a, r[0], r[1]);
      mma16816(out[2 * dt + 1], a, r[2], r[3]);
    }
  }
}
// accumulator tiles (rows g / g+8 of the warp's 16) -> bf16 global, rows < n only
__device__ __forceinline__ void store_acc_rows(const float (&acc)[8][4], __nv_bfloat16* dst, long long ld,
                                               int batch, int n, int head, int row_lo) {
  const int lane = threadIdx.x & 31, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = row_lo + 8 * h;
    if (row < n) {
      __nv_bfloat16* p = dst + (static_cast<long long>(batch) * n + row) * ld + head * AB_T + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<uint32_t*>(p + nt * 8) = pack_bf16x2(acc[nt][2 * h], acc[nt][2 * h + 1]);
    }
  }
}

__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_dq_mma_kernel(const AttnBwdParams p) {
  __shared__ __align__(128) __nv_bfloat16 sK[AB_T * AB_LD];
  __shared__ __align__(128) __nv_bfloat16 sV[AB_T * AB_LD];
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * AB_T, head = blockIdx.y, batch = blockIdx.z;
  const int w16 = warp * 16;
  const int row_lo = row0 + w16 + g;
  // stage Q / dO through the K / V buffers into A fragments
  load_tile(sK, p.q, p.ldq, batch, p.nq, head, row0);
  load_tile(sV, p.dout, p.lddo, batch, p.nq, head, row0);
  __syncthreads();
  uint32_t qa[4][4], da[4][4];
  load_a_frags(qa, sK, w16);
  load_a_frags(da, sV, w16);
  // D = sum_d dO O for rows g / g+8 (this thread's 16 columns of each, then the quad)
  float dsum[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = row_lo + 8 * h;
    if (row < p.nq) {
      const __nv_bfloat16* po = p.o + (static_cast<long long>(batch) * p.nq + row) * p.ldo + head * AB_T + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float2 o = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(po + ks * 16 + 8 * c)));
          const float2 d = unpack_bf16x2(da[ks][h + 2 * c]);
          dsum[h] = fmaf(o.x, d.x, fmaf(o.y, d.y, dsum[h]));
        }
    }
    dsum[h] += __shfl_xor_sync(0xffffffffu, dsum[h], 1);
    dsum[h] += __shfl_xor_sync(0xffffffffu, dsum[h], 2);
  }
  const int nt_kv = (p.nkv + AB_T - 1) / AB_T;
  // ---- pass 1: lse ----
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  for (int j = 0; j < nt_kv; ++j) {
    __syncthreads();
    load_tile(sK, p.k, p.ldk, batch, p.nkv, head, j * AB_T);
    __syncthreads();
    float s[8][4];
    mma_a_bt(s, qa, sK);
    const int valid = p.nkv - j * AB_T;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * t + (e & 1);
        s[nt][e] = col < valid ? s[nt][e] * p.scale : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float m_new = fmaxf(m[h], mx[h]);   // finite: column 0 of every tile is a valid key
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) sum += __expf(s[nt][2 * h] - m_new) + __expf(s[nt][2 * h + 1] - m_new);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      l[h] = l[h] * __expf(m[h] - m_new) + sum;
      m[h] = m_new;
    }
  }
  const float lse[2] = {m[0] + __logf(l[0]), m[1] + __logf(l[1])};
  // ---- pass 2: dQ ----
  float dq[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[nt][e] = 0.f;
  for (int j = 0; j < nt_kv; ++j) {
    __syncthreads();
    load_tile(sK, p.k, p.ldk, batch, p.nkv, head, j * AB_T);
    load_tile(sV, p.v, p.ldv, batch, p.nkv, head, j * AB_T);
    __syncthreads();
    float s[8][4], dp[8][4];
    mma_a_bt(s, qa, sK);
    mma_a_bt(dp, da, sV);
    const int valid = p.nkv - j * AB_T;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * t + (e & 1);
        const int h = e >> 1;
        s[nt][e] = col < valid ? __expf(s[nt][e] * p.scale - lse[h]) * (dp[nt][e] - dsum[h]) * p.scale : 0.f;
      }
    mma_x_b(dq, s, sK);
  }
  store_acc_rows(dq, p.dq, p.lddq, batch, p.nq, head, row_lo);
  if (t == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = row_lo + 8 * h;
      if (row < p.nq) {
        const long long idx = (static_cast<long long>(batch) * p.heads + head) * p.nq + row;
        p.lse[idx] = lse[h];
        p.dsum[idx] = dsum[h];
      }
    }
  }
}

// accumulator tiles -> fp32 atomics into acc [B * n, heads * 64] (query-split partial sums)
__device__ __forceinline__ void atomic_acc_rows(const float (&acc)[8][4], float* dst, int heads, int batch, int n,
                                                int head, int row_lo) {
  const int lane = threadIdx.x & 31, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = row_lo + 8 * h;
    if (row < n) {
      float* p = dst + (static_cast<long long>(batch) * n + row) * (heads * AB_T) + head * AB_T + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        atomicAdd(p + nt * 8, acc[nt][2 * h]);
        atomicAdd(p + nt * 8 + 1, acc[nt][2 * h + 1]);
      }
    }
  }
}

// NSPLIT == 0: one CTA per (key tile, head, batch) walks every query tile and stores bf16 dK / dV.
// NSPLIT == 1 (query-split, few keys x very many queries: reference_attn's 24 samples per ray against
// the 77 text tokens): blockIdx.x = key_tile * nsplit + split, the CTA walks its share of the query
// tiles and adds its partial sums into the zeroed fp32 accumulators kv_acc = [dK | dV].
template <int NSPLIT>
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_dkdv_mma_kernel(const AttnBwdParams p, float* kv_acc, int nsplit) {
  __shared__ __align__(128) __nv_bfloat16 sQ[AB_T * AB_LD];
  __shared__ __align__(128) __nv_bfloat16 sDO[AB_T * AB_LD];
  __shared__ float sLse[AB_T], sD[AB_T];
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int key_tile = NSPLIT ? static_cast<int>(blockIdx.x) / nsplit : static_cast<int>(blockIdx.x);
  const int split = NSPLIT ? static_cast<int>(blockIdx.x) - key_tile * nsplit : 0;
  const int key0 = key_tile * AB_T, head = blockIdx.y, batch = blockIdx.z;
  const int w16 = warp * 16;
  const int key_lo = key0 + w16 + g;
  load_tile(sQ, p.k, p.ldk, batch, p.nkv, head, key0);
  load_tile(sDO, p.v, p.ldv, batch, p.nkv, head, key0);
  __syncthreads();
  uint32_t ka[4][4], va[4][4];
  load_a_frags(ka, sQ, w16);
  load_a_frags(va, sDO, w16);
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      dk[nt][e] = 0.f;
      dv[nt][e] = 0.f;
    }
  const bool key_ok[2] = {key_lo < p.nkv, key_lo + 8 < p.nkv};
  const int nt_q = (p.nq + AB_T - 1) / AB_T;
  const int per = NSPLIT ? (nt_q + nsplit - 1) / nsplit : nt_q;
  const int i_begin = split * per, i_end = min(nt_q, i_begin + per);
  for (int i = i_begin; i < i_end; ++i) {
    __syncthreads();
    load_tile(sQ, p.q, p.ldq, batch, p.nq, head, i * AB_T);
    load_tile(sDO, p.dout, p.lddo, batch, p.nq, head, i * AB_T);
    if (threadIdx.x < AB_T) {
      const int qrow = i * AB_T + threadIdx.x;
      const long long idx = (static_cast<long long>(batch) * p.heads + head) * p.nq + qrow;
      sLse[threadIdx.x] = qrow < p.nq ? p.lse[idx] : INFINITY;   // exp(-inf) = 0 for rows past the end
      sD[threadIdx.x] = qrow < p.nq ? p.dsum[idx] : 0.f;
    }
    __syncthreads();
    float st[8][4], dpt[8][4];
    mma_a_bt(st, ka, sQ);      // S^T  [keys][queries]
    mma_a_bt(dpt, va, sDO);    // dP^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * t + (e & 1);
        const float pr = key_ok[e >> 1] ? __expf(st[nt][e] * p.scale - sLse[col]) : 0.f;
        st[nt][e] = pr;                                          // P^T
        dpt[nt][e] = pr * (dpt[nt][e] - sD[col]) * p.scale;      // dS^T
      }
    mma_x_b(dv, st, sDO);      // dV += P^T dO
    mma_x_b(dk, dpt, sQ);      // dK += dS^T Q
  }
  if (NSPLIT) {
    atomic_acc_rows(dk, kv_acc, p.heads, batch, p.nkv, head, key_lo);
    atomic_acc_rows(dv, kv_acc + static_cast<long long>(gridDim.z) * p.nkv * p.heads * AB_T, p.heads, batch, p.nkv,
                    head, key_lo);
  } else {
    store_acc_rows(dk, p.dk, p.lddk, batch, p.nkv, head, key_lo);
    store_acc_rows(dv, p.dv, p.lddv, batch, p.nkv, head, key_lo);
  }
}

// fp32 accumulators [dK | dV] (each [rows, inner]) -> bf16 dk / dv (row strides lddk / lddv)
__global__ void __launch_bounds__(256)
attention_bwd_kv_finish_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dk, long long lddk,
                               __nv_bfloat16* __restrict__ dv, long long lddv, long long rows, int inner) {
  pdl_wait();
  const long long half = rows * inner;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < 2 * half;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 2) {
    const bool is_v = i >= half;
    const long long j = is_v ? i - half : i;
    const long long r = j / inner;
    const int c = static_cast<int>(j - r * inner);
    __nv_bfloat16* dst = (is_v ? dv + r * lddv : dk + r * lddk) + c;
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(acc[i], acc[i + 1]);
  }
}

constexpr int AB_SMEM_STATS = 2 * AB_TILE_B + AB_TILE_F;
constexpr int AB_SMEM_DQ = 5 * AB_TILE_B + 2 * AB_TILE_F;
constexpr int AB_SMEM_DKDV = 6 * AB_TILE_B + 2 * AB_TILE_F + 2 * AB_T * 4;

}  // namespace cd360

using namespace cd360;

extern "C" int cd360_attention_bwd_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk,
                                        const void* v, int64_t ldv, const void* o, int64_t ldo,
                                        const void* dout, int64_t lddo, void* dq, int64_t lddq,
                                        void* dk, int64_t lddk, void* dv, int64_t lddv,
                                        float* lse, float* dsum, int32_t batch, int32_t heads,
                                        int32_t nq, int32_t nkv, cd360_stream_t stream_) {
  if (!q || !k || !v || !o || !dout || !dq || !lse || !dsum) return CD360_ERR_NULL;
  if ((dk == nullptr) != (dv == nullptr)) return CD360_ERR_NULL;
  if (batch <= 0 || heads <= 0 || nq <= 0 || nkv <= 0 || batch > 65535 || heads > 65535)
    return CD360_ERR_SHAPE;
  const int64_t lds[8] = {ldq, ldk, ldv, ldo, lddo, lddq, dk ? lddk : 8, dv ? lddv : 8};
  for (int i = 0; i < 8; ++i)
    if ((lds[i] & 7) || lds[i] < (i >= 6 && !dk ? 8 : heads * AB_T)) return CD360_ERR_ALIGN;
  const void* ptrs[8] = {q, k, v, o, dout, dq, dk, dv};
  for (int i = 0; i < 8; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return CD360_ERR_ALIGN;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attention_bwd_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             AB_SMEM_STATS) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             AB_SMEM_DQ) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             AB_SMEM_DKDV) != cudaSuccess)
      return CD360_ERR_LAUNCH;
    attr_done = true;
  }
  AttnBwdParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.o = reinterpret_cast<const __nv_bfloat16*>(o);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddo = lddo;
  p.lse = lse; p.dsum = dsum;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.nq = nq; p.nkv = nkv; p.heads = heads;
  p.scale = 0.125f;
  const int impl = attn_bwd_impl();
  if (impl == 0)
    return attention_bwd_tcgen05(q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, dq, lddq, dk, lddk, dv, lddv, lse, dsum,
                                 nullptr, batch, heads, nq, nkv, 1, dk != nullptr ? 3 : 1, stream);
  const dim3 gq((nq + AB_T - 1) / AB_T, heads, batch);
  const int use_wmma = impl == 2;
  if (!use_wmma) {
    if (launch_ex(attention_bwd_dq_mma_kernel, gq, dim3(AB_THREADS), 0, stream, 1, p) != cudaSuccess)
      return CD360_ERR_LAUNCH;
    if (dk != nullptr) {
      const dim3 gk((nkv + AB_T - 1) / AB_T, heads, batch);
      if (launch_ex(attention_bwd_dkdv_mma_kernel<0>, gk, dim3(AB_THREADS), 0, stream, 1, p,
                    static_cast<float*>(nullptr), 1) != cudaSuccess)
        return CD360_ERR_LAUNCH;
    }
    CD360_CHECK_LAUNCH();
    return CD360_OK;
  }
  if (launch_ex(attention_bwd_stats_kernel, gq, dim3(AB_THREADS), AB_SMEM_STATS, stream, 1, p) != cudaSuccess)
    return CD360_ERR_LAUNCH;
  if (launch_ex(attention_bwd_dq_kernel, gq, dim3(AB_THREADS), AB_SMEM_DQ, stream, 1, p) != cudaSuccess)
    return CD360_ERR_LAUNCH;
  if (dk != nullptr) {
    const dim3 gk((nkv + AB_T - 1) / AB_T, heads, batch);
    if (launch_ex(attention_bwd_dkdv_kernel, gk, dim3(AB_THREADS), AB_SMEM_DKDV, stream, 1, p) != cudaSuccess)
      return CD360_ERR_LAUNCH;
  }
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_attention_bwd_kv_split_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk,
                                                 const void* v, int64_t ldv, const void* dout, int64_t lddo,
                                                 const float* lse, const float* dsum, float* kv_acc,
                                                 void* dk, int64_t lddk, void* dv, int64_t lddv,
                                                 int32_t batch, int32_t heads, int32_t nq, int32_t nkv,
                                                 int32_t nsplit, cd360_stream_t stream_) {
  if (!q || !k || !v || !dout || !lse || !dsum || !kv_acc || !dk || !dv) return CD360_ERR_NULL;
  if (batch <= 0 || heads <= 0 || nq <= 0 || nkv <= 0 || nsplit <= 0 || batch > 65535 || heads > 65535)
    return CD360_ERR_SHAPE;
  const int64_t lds[6] = {ldq, ldk, ldv, lddo, lddk, lddv};
  for (int i = 0; i < 6; ++i)
    if ((lds[i] & 7) || lds[i] < heads * AB_T) return CD360_ERR_ALIGN;
  const void* ptrs[7] = {q, k, v, dout, dk, dv, kv_acc};
  for (int i = 0; i < 7; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return CD360_ERR_ALIGN;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  AttnBwdParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.lddo = lddo;
  p.lse = const_cast<float*>(lse);
  p.dsum = const_cast<float*>(dsum);
  p.nq = nq; p.nkv = nkv; p.heads = heads;
  p.scale = 0.125f;
  if (attn_bwd_impl() == 0) {
    const int rc = attention_bwd_tcgen05(q, ldq, k, ldk, v, ldv, nullptr, 0, dout, lddo, nullptr, 0, nullptr, 0, nullptr,
                                         0, const_cast<float*>(lse), const_cast<float*>(dsum), kv_acc, batch, heads, nq,
                                         nkv, nsplit, 2, stream);
    if (rc != CD360_OK) return rc;
  } else {
    const int key_tiles = (nkv + AB_T - 1) / AB_T;
    const int q_tiles = (nq + AB_T - 1) / AB_T;
    if (nsplit > q_tiles) nsplit = q_tiles;
    const dim3 gk(static_cast<unsigned>(key_tiles * nsplit), heads, batch);
    if (launch_ex(attention_bwd_dkdv_mma_kernel<1>, gk, dim3(AB_THREADS), 0, stream, 1, p, kv_acc, nsplit) !=
        cudaSuccess)
      return CD360_ERR_LAUNCH;
  }
  const long long rows = static_cast<long long>(batch) * nkv;
  const long long pairs = rows * heads * AB_T;   // 2 * rows * inner / 2
  unsigned blocks = static_cast<unsigned>((pairs + 255) / 256);
  if (blocks > 1184u) blocks = 1184u;
  if (launch_ex(attention_bwd_kv_finish_kernel, dim3(blocks), dim3(256), 0, stream, 1,
                static_cast<const float*>(kv_acc), reinterpret_cast<__nv_bfloat16*>(dk),
                static_cast<long long>(lddk), reinterpret_cast<__nv_bfloat16*>(dv), static_cast<long long>(lddv),
                rows, heads * AB_T) != cudaSuccess)
    return CD360_ERR_LAUNCH;
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}
