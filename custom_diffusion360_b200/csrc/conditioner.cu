// conditioner.cu — the kernels of the text conditioner that are not dense contractions
// (reference sgm/modules/encoders/modules.py:377-517 FrozenCLIPEmbedder and
// :622-772 FrozenOpenCLIPEmbedder):
//   * token + positional embedding gather (HF CLIPTextEmbeddings / open_clip token_embedding +
//     positional_embedding, modules.py:498-501, 716-728),
//   * causal self-attention over the 77-token context, head dim 64 (CLIP-L: 12 heads, ViT-bigG text
//     tower: 20 heads; modules.py:447-453 `_build_causal_attention_mask`, open_clip `attn_mask`),
//   * row gather for the end-of-text pooling (modules.py:737-743).
// The projections / MLPs run on the tcgen05 GEMM (quick-GELU / GELU epilogues), LayerNorm on the
// shared row kernel.  This runs once per prompt: 77 x 77 scores per head are latency-bound work, so
// the attention is a shared-memory / warp-shuffle kernel (one CTA per (batch, head), K and V staged
// once in smem, one warp per query row), not a tensor-core one.
#include "cd360_common.cuh"

namespace cd360 {

// out[b*ctx + t, :] = bf16(tok_emb[ids[b, t], :] + pos_emb[t, :])
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tok_emb,
                    const float* __restrict__ pos_emb, __nv_bfloat16* __restrict__ out, int rows, int ctx,
                    int w, int vocab) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = w >> 2;
  const long long total = static_cast<long long>(rows) * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / nvec);
    const int c4 = static_cast<int>(i - static_cast<long long>(row) * nvec);
    int id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // nn.Embedding would raise; ids are validated on the host
    const float4 a = __ldg(reinterpret_cast<const float4*>(tok_emb + static_cast<long long>(id) * w) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos_emb + static_cast<long long>(row % ctx) * w) + c4);
    uint2 o;
    o.x = pack_bf16x2(a.x + b.x, a.y + b.y);
    o.y = pack_bf16x2(a.z + b.z, a.w + b.w);
    *reinterpret_cast<uint2*>(out + static_cast<long long>(row) * w + c4 * 4) = o;
  }
}

constexpr int CA_MAXN = 128;    // keys per sequence (77 for both text towers)
constexpr int CA_WARPS = 8;
constexpr int CA_KSTRIDE = 33;  // 32-bit words per K row (64 bf16 + 1 pad word): conflict-free column reads

// One CTA per (head, batch).  K [n][64] (padded rows) and V [n][64] staged in shared memory; warp w
// takes query rows w, w + 8, ...; lane l scores keys l, l + 32, ... (masked j > i), softmax by warp
// shuffles, then every lane accumulates its two output dims over the keys.
__global__ void __launch_bounds__(CA_WARPS * 32)
attention_causal_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k,
                        long long ldk, const __nv_bfloat16* __restrict__ v, long long ldv,
                        __nv_bfloat16* __restrict__ out, long long ldo, int n, float scale_log2e) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ uint32_t s_k[CA_MAXN * CA_KSTRIDE];
  __shared__ uint32_t s_v[CA_MAXN * 32];
  __shared__ float s_q[CA_WARPS][64];
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = static_cast<long long>(b) * n;
  for (int i = threadIdx.x; i < n * 32; i += blockDim.x) {
    const int j = i >> 5, c = i & 31;
    s_k[j * CA_KSTRIDE + c] = *reinterpret_cast<const uint32_t*>(k + (row0 + j) * ldk + head * 64 + c * 2);
    s_v[j * 32 + c] = *reinterpret_cast<const uint32_t*>(v + (row0 + j) * ldv + head * 64 + c * 2);
  }
  __syncthreads();
  for (int i = warp; i < n; i += CA_WARPS) {
    {
      const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(q + (row0 + i) * ldq + head * 64 + lane * 2));
      s_q[warp][lane * 2] = f.x;
      s_q[warp][lane * 2 + 1] = f.y;
    }
    __syncwarp();
    float sc[CA_MAXN / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < CA_MAXN / 32; ++t) {
      const int j = lane + 32 * t;
      float acc = -INFINITY;
      if (j <= i) {
        acc = 0.f;
        const uint32_t* kr = s_k + j * CA_KSTRIDE;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float2 kk = unpack_bf16x2(kr[c]);
          acc = fmaf(s_q[warp][2 * c], kk.x, acc);
          acc = fmaf(s_q[warp][2 * c + 1], kk.y, acc);
        }
        acc *= scale_log2e;
      }
      sc[t] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < CA_MAXN / 32; ++t) {
      sc[t] = (lane + 32 * t <= i) ? exp2f(sc[t] - mx) : 0.f;
      sum += sc[t];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int t = 0; t < CA_MAXN / 32; ++t) {
      if (32 * t > i) break;  // warp-uniform
      const int jn = min(32, i + 1 - 32 * t);
      for (int jj = 0; jj < jn; ++jj) {
        const float pj = __shfl_sync(0xffffffffu, sc[t], jj);
        const float2 vv = unpack_bf16x2(s_v[(32 * t + jj) * 32 + lane]);
        o0 = fmaf(pj, vv.x, o0);
        o1 = fmaf(pj, vv.y, o1);
      }
    }
    *reinterpret_cast<uint32_t*>(out + (row0 + i) * ldo + head * 64 + lane * 2) = pack_bf16x2(o0 * inv, o1 * inv);
    __syncwarp();  // s_q[warp] is rewritten by the next row
  }
}

// out[b, :] = fp32(x[idx[b], :])
__global__ void __launch_bounds__(256)
gather_rows_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const int32_t* __restrict__ idx,
                   float* __restrict__ out, int nrows, int c, long long src_rows) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(nrows) * c;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / c);
    const int col = static_cast<int>(i - static_cast<long long>(r) * c);
    long long s = idx[r];
    s = s < 0 ? 0 : (s >= src_rows ? src_rows - 1 : s);
    out[i] = __bfloat162float(x[s * ldx + col]);
  }
}

}  // namespace cd360

using namespace cd360;

extern "C" int cd360_embed_tokens(const int32_t* ids, const float* tok_emb, const float* pos_emb, void* out,
                                  int32_t rows, int32_t ctx, int32_t w, int32_t vocab, cd360_stream_t stream_) {
  if (!ids || !tok_emb || !pos_emb || !out) return CD360_ERR_NULL;
  if (rows <= 0 || ctx <= 0 || w <= 0 || (w & 3) || vocab <= 0 || rows % ctx != 0) return CD360_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(tok_emb) & 15) || (reinterpret_cast<uintptr_t>(pos_emb) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 7))
    return CD360_ERR_ALIGN;
  const long long total = static_cast<long long>(rows) * (w >> 2);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  launch_ex(embed_tokens_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, ids, tok_emb,
            pos_emb, reinterpret_cast<__nv_bfloat16*>(out), rows, ctx, w, vocab);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_attention_causal_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                           int64_t ldv, void* out, int64_t ldo, int32_t batch, int32_t heads,
                                           int32_t n, cd360_stream_t stream_) {
  if (!q || !k || !v || !out) return CD360_ERR_NULL;
  if (batch <= 0 || heads <= 0 || n <= 0 || n > CA_MAXN) return CD360_ERR_SHAPE;
  if ((ldq & 1) || (ldk & 1) || (ldv & 1) || (ldo & 1) || ldq < heads * 64 || ldk < heads * 64 || ldv < heads * 64 ||
      ldo < heads * 64)
    return CD360_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(q) & 3) || (reinterpret_cast<uintptr_t>(k) & 3) ||
      (reinterpret_cast<uintptr_t>(v) & 3) || (reinterpret_cast<uintptr_t>(out) & 3))
    return CD360_ERR_ALIGN;
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  launch_ex(attention_causal_kernel, dim3(heads, batch), dim3(CA_WARPS * 32), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, reinterpret_cast<const __nv_bfloat16*>(q),
            static_cast<long long>(ldq), reinterpret_cast<const __nv_bfloat16*>(k), static_cast<long long>(ldk),
            reinterpret_cast<const __nv_bfloat16*>(v), static_cast<long long>(ldv),
            reinterpret_cast<__nv_bfloat16*>(out), static_cast<long long>(ldo), n, scale_log2e);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_gather_rows_bf16_f32(const void* x, int64_t ldx, const int32_t* idx, float* out, int32_t nrows,
                                          int32_t c, int64_t src_rows, cd360_stream_t stream_) {
  if (!x || !idx || !out) return CD360_ERR_NULL;
  if (nrows <= 0 || c <= 0 || src_rows <= 0 || ldx < c) return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(nrows) * c;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  launch_ex(gather_rows_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1,
            reinterpret_cast<const __nv_bfloat16*>(x), static_cast<long long>(ldx), idx, out, nrows, c,
            static_cast<long long>(src_rows));
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}
