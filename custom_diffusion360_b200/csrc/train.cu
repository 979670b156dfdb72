// train.cu — backward / loss / optimiser kernels of the training step (SURVEY.md §8 a20, a21;
// reference: sgm/models/diffusion.py:221-272 training_step, loss.py:140-216, and the autograd
// backward of the modules in sgm/modules/attention.py / openaimodel.py / nerfsd_pytorch3d.py).
//
// Only the pose weights train (diffusion.py:139-144, trainkeys 'pose'): pose_emb_layers and the
// FeatureNeRF MLPs of the 12 pose blocks.  The gradient therefore flows as ACTIVATION gradients
// through the frozen UNet downstream of the first pose block; the dense parts of that (dX of every
// Linear / conv) reuse the tcgen05 GEMM with transposed / flipped weight packs, and the kernels
// here supply everything else: normalisation / GEGLU / resampling backward, the FeatureNeRF
// gather / view-softmax / volume-rendering backward, operand transposes and column sums for the
// weight gradients, the loss and its gradient, antialiased resize of the supervision masks, AdamW.
// All are HBM-bound elementwise / reduction kernels: vectorised 16-byte accesses, fp32 arithmetic.
#include "cd360_common.cuh"

namespace cd360 {

__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
         d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
// d/dx [x sigmoid(x)]
__device__ __forceinline__ float silu_grad_f(float x) {
  const float s = sigmoid_f(x);
  return s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ float block_sum(float v, float* s_red) {  // blockDim.x multiple of 32, <= 1024
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // s_red reuse
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += s_red[i];  // fixed order: deterministic
  return t;
}

// =================================================================================================
// LayerNorm backward (nn.LayerNorm x3 per BasicTransformerBlock, attention.py:531-533), one warp per
// row:  g = dy*gamma,  dx = rstd (g - mean(g) - xhat mean(g xhat)) [+ add]
// =================================================================================================
constexpr int LNB_MAXV = 5;  // c <= 1280

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                     const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ add,
                     __nv_bfloat16* __restrict__ dx, int rows, int c, float eps) {
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = c >> 3;
  const long long base = static_cast<long long>(warp) * c;
  float v[LNB_MAXV][8], g[LNB_MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LNB_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      unpack8f(__ldg(reinterpret_cast<const uint4*>(x + base + vi * 8)), v[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += v[i][k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / c;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LNB_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[i][k] -= mean;
        sq = fmaf(v[i][k], v[i][k], sq);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / c + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < LNB_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      float d[8];
      unpack8f(__ldg(reinterpret_cast<const uint4*>(dy + base + vi * 8)), d);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[i][k] *= rstd;  // xhat
        g[i][k] = d[k] * gg[k];
        s1 += g[i][k];
        s2 = fmaf(g[i][k], v[i][k], s2);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float m1 = s1 / c, m2 = s2 / c;
#pragma unroll
  for (int i = 0; i < LNB_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rstd * (g[i][k] - m1 - v[i][k] * m2);
      if (add != nullptr) {
        float a[8];
        unpack8f(__ldg(reinterpret_cast<const uint4*>(add + base + vi * 8)), a);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += a[k];
      }
      *reinterpret_cast<uint4*>(dx + base + vi * 8) = pack8f(o);
    }
  }
}

// =================================================================================================
// GroupNorm(32) [+SiLU] backward over NHWC, input = virtual concat [x0 | x1] (GroupNorm32 + SiLU,
// openaimodel.py:280-283,315-317; Normalize, attention.py:118).
//   z = gamma xhat + beta, y = act(z);  dxhat = dy act'(z) gamma
//   dx = rstd (dxhat - mean_g(dxhat) - xhat mean_g(dxhat xhat)) [+ add]
// Kernel 1: one CTA per (group, image): mean / rstd of x, then the two group means -> ws[b][g][4].
// Kernel 2: elementwise.
// =================================================================================================
constexpr int GNB_GROUPS = 32;

__device__ __forceinline__ float2 ld_pair(const __nv_bfloat16* x0, int c0, const __nv_bfloat16* x1,
                                          int c1, long long row, int ch) {
  const __nv_bfloat16* src = ch < c0 ? x0 + row * c0 + ch : x1 + row * c1 + (ch - c0);
  return unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(src)));
}

// read a float at the same shared-memory offset in CTA `rank` of the cluster (distributed shared memory)
__device__ __forceinline__ float ld_cluster_f32(const float* p, uint32_t rank) {
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(p)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

// One (group, image) per CLUSTER of `splits` CTAs (1 = a plain CTA): each CTA owns a slice of the rows, the
// two dependent reductions (moments, then the two gradient sums) are combined through distributed shared
// memory in rank order — deterministic — so 32 x batch work items spread over up to 8 x as many SMs (one
// CTA per group walked 4096 x 10 channels alone: 36 us per launch on the critical backward chain).
__global__ void __launch_bounds__(512)
groupnorm_bwd_stats_kernel(const __nv_bfloat16* __restrict__ x0, int c0,
                           const __nv_bfloat16* __restrict__ x1, int c1,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const __nv_bfloat16* __restrict__ dy, float* __restrict__ ws, int hw,
                           float eps, int apply_silu, int splits) {
  pdl_wait();
  __shared__ float s_red[32];
  __shared__ float s_x[2], s_y[2], s_bc[2];
  const int g = blockIdx.x / splits, b = blockIdx.y;
  const uint32_t rank = splits > 1 ? cluster_ctarank() : 0u;
  const int ctot = c0 + c1;
  const int cg = ctot / GNB_GROUPS;
  const int pairs = cg >> 1;
  const long long img = static_cast<long long>(b) * hw;
  const int rows_per = (hw + splits - 1) / splits;
  const int r0 = static_cast<int>(rank) * rows_per;
  const int r1 = min(hw, r0 + rows_per);
  const int items = r1 > r0 ? (r1 - r0) * pairs : 0;
  float s = 0.f, ss = 0.f;
  // four independent loads in flight per thread (one dependent 4-byte load per iteration left the kernel
  // latency-bound: 22 us for 5 MB)
  for (int i = threadIdx.x; i < items; i += 4 * blockDim.x) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int iu = i + u * blockDim.x;
      const int r = iu / pairs, j = iu - r * pairs;
      v[u] = iu < items ? ld_pair(x0, c0, x1, c1, img + r0 + r, g * cg + 2 * j) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s += v[u].x + v[u].y;
      ss = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, ss));
    }
  }
  s = block_sum(s, s_red);
  ss = block_sum(ss, s_red);
  if (splits > 1) {
    if (threadIdx.x == 0) { s_x[0] = s; s_x[1] = ss; }
    cluster_sync_all();
    if (threadIdx.x == 0) {
      float t0 = 0.f, t1 = 0.f;
      for (int k = 0; k < splits; ++k) {  // rank order: every CTA of the cluster gets the same bits
        t0 += ld_cluster_f32(&s_x[0], k);
        t1 += ld_cluster_f32(&s_x[1], k);
      }
      s_bc[0] = t0; s_bc[1] = t1;
    }
    __syncthreads();
    s = s_bc[0]; ss = s_bc[1];
  }
  const float n = static_cast<float>(hw) * cg;
  const float mean = s / n;
  float var = ss / n - mean * mean;
  if (var < 0.f) var = 0.f;
  const float rstd = rsqrtf(var + eps);
  float a1 = 0.f, a2 = 0.f;
  for (int i = threadIdx.x; i < items; i += 4 * blockDim.x) {
    float2 v[4], d[4];
    int chs[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int iu = i + u * blockDim.x;
      const int r = iu / pairs, j = iu - r * pairs;
      chs[u] = g * cg + 2 * j;
      const bool ok = iu < items;
      v[u] = ok ? ld_pair(x0, c0, x1, c1, img + r0 + r, chs[u]) : make_float2(0.f, 0.f);
      d[u] = ok ? unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(dy + (img + r0 + r) * ctot + chs[u])))
                : make_float2(0.f, 0.f);            // dy = 0: the item adds nothing to either sum
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ch = chs[u];
      const float xh0 = (v[u].x - mean) * rstd, xh1 = (v[u].y - mean) * rstd;
      const float g0 = gamma[ch], g1 = gamma[ch + 1];
      float t0 = d[u].x * g0, t1 = d[u].y * g1;
      if (apply_silu) {
        t0 *= silu_grad_f(fmaf(g0, xh0, beta[ch]));
        t1 *= silu_grad_f(fmaf(g1, xh1, beta[ch + 1]));
      }
      a1 += t0 + t1;
      a2 = fmaf(t0, xh0, fmaf(t1, xh1, a2));
    }
  }
  a1 = block_sum(a1, s_red);
  a2 = block_sum(a2, s_red);
  if (splits > 1) {
    if (threadIdx.x == 0) { s_y[0] = a1; s_y[1] = a2; }
    cluster_sync_all();
    if (rank == 0 && threadIdx.x == 0) {
      a1 = 0.f; a2 = 0.f;
      for (int k = 0; k < splits; ++k) {
        a1 += ld_cluster_f32(&s_y[0], k);
        a2 += ld_cluster_f32(&s_y[1], k);
      }
    }
  }
  if (rank == 0 && threadIdx.x == 0) {
    float* o = ws + (static_cast<long long>(b) * GNB_GROUPS + g) * 4;
    o[0] = mean; o[1] = rstd; o[2] = a1 / n; o[3] = a2 / n;
  }
  if (splits > 1) cluster_sync_all();  // peers keep their shared memory alive until rank 0 has read it
}

__global__ void __launch_bounds__(256)
groupnorm_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x0, int c0,
                           const __nv_bfloat16* __restrict__ x1, int c1,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const __nv_bfloat16* __restrict__ dy,
                           const __nv_bfloat16* __restrict__ add0, long long ld_add0,
                           const __nv_bfloat16* __restrict__ add1, long long ld_add1,
                           __nv_bfloat16* __restrict__ dx0, __nv_bfloat16* __restrict__ dx1,
                           const float* __restrict__ ws, int batch, int hw, int apply_silu) {
  pdl_wait();
  const int ctot = c0 + c1;
  const int cg = ctot / GNB_GROUPS;
  const int pairs = ctot >> 1;
  const long long total = static_cast<long long>(batch) * hw * pairs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / pairs;
    const int ch = static_cast<int>(i - row * pairs) * 2;
    const int b = static_cast<int>(row / hw);
    const float2 v = ld_pair(x0, c0, x1, c1, row, ch);
    const float2 d = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(dy + row * ctot + ch)));
    float out[2];
    const float xs[2] = {v.x, v.y}, ds[2] = {d.x, d.y};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cc = ch + e;
      const float4 st = __ldg(reinterpret_cast<const float4*>(ws) + b * GNB_GROUPS + cc / cg);
      const float xh = (xs[e] - st.x) * st.y;
      const float gm = gamma[cc];
      float t = ds[e] * gm;
      if (apply_silu) t *= silu_grad_f(fmaf(gm, xh, beta[cc]));
      out[e] = st.y * (t - st.z - xh * st.w);
    }
    if (ch < c0) {
      if (add0 != nullptr) {
        const float2 a = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(add0 + row * ld_add0 + ch)));
        out[0] += a.x; out[1] += a.y;
      }
      *reinterpret_cast<uint32_t*>(dx0 + row * c0 + ch) = pack_bf16x2(out[0], out[1]);
    } else {
      const int cc = ch - c0;
      if (add1 != nullptr) {
        const float2 a = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(add1 + row * ld_add1 + cc)));
        out[0] += a.x; out[1] += a.y;
      }
      *reinterpret_cast<uint32_t*>(dx1 + row * c1 + cc) = pack_bf16x2(out[0], out[1]);
    }
  }
}

// =================================================================================================
// GEGLU backward (attention.py:94-96): h = a * gelu(gate);  da = dh gelu(gate),
// dgate = dh a (Phi(gate) + gate phi(gate)).  `raw` holds [a | gate] interleaved in blocks of `blk`
// columns (blk = F: plain chunk(2) layout; blk = the GEMM's pack block: the packed-weight layout).
// =================================================================================================
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const __nv_bfloat16* __restrict__ raw, const __nv_bfloat16* __restrict__ dh,
                 __nv_bfloat16* __restrict__ draw, long long rows, int F, int blk) {
  pdl_wait();
  const int nvec = F >> 3;
  const long long total = rows * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / nvec;
    const int j = static_cast<int>(i - row * nvec) * 8;
    const int t = j / blk, w = j - t * blk;
    const long long ca = row * 2 * F + static_cast<long long>(t) * 2 * blk + w;
    float a[8], gt[8], d[8], oa[8], og[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(raw + ca)), a);
    unpack8f(__ldg(reinterpret_cast<const uint4*>(raw + ca + blk)), gt);
    unpack8f(__ldg(reinterpret_cast<const uint4*>(dh + row * F + j)), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float x = gt[k];
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
      oa[k] = d[k] * x * cdf;
      og[k] = d[k] * a[k] * fmaf(x, pdf, cdf);
    }
    *reinterpret_cast<uint4*>(draw + ca) = pack8f(oa);
    *reinterpret_cast<uint4*>(draw + ca + blk) = pack8f(og);
  }
}

// GEGLU forward on a KEPT pre-activation (training: raw is saved for geglu_bwd instead of being
// recomputed): h = a * gelu(gate), same [a | gate] block layout as above, exact erf like the reference's
// F.gelu (attention.py:96).
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(const __nv_bfloat16* __restrict__ raw, __nv_bfloat16* __restrict__ h, long long rows, int F,
                 int blk) {
  pdl_wait();
  const int nvec = F >> 3;
  const long long total = rows * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / nvec;
    const int j = static_cast<int>(i - row * nvec) * 8;
    const int t = j / blk, w = j - t * blk;
    const long long ca = row * 2 * F + static_cast<long long>(t) * 2 * blk + w;
    float a[8], gt[8], o[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(raw + ca)), a);
    unpack8f(__ldg(reinterpret_cast<const uint4*>(raw + ca + blk)), gt);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      o[k] = a[k] * 0.5f * gt[k] * (1.0f + erff(gt[k] * 0.70710678118654752440f));
    *reinterpret_cast<uint4*>(h + row * F + j) = pack8f(o);
  }
}

// =================================================================================================
// small layout / reduction helpers
// =================================================================================================
// out = a + b (bf16), n multiple of 8
__global__ void __launch_bounds__(256)
add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                __nv_bfloat16* __restrict__ out, long long nvec) {
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(a) + i), x);
    unpack8f(__ldg(reinterpret_cast<const uint4*>(b) + i), y);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] += y[k];
    reinterpret_cast<uint4*>(out)[i] = pack8f(x);
  }
}

// out = dy * silu'(pre), fp32 (backward of the SiLU inside time_embed / label_emb / emb_layers,
// openaimodel.py:679-713, 270-276): silu'(x) = s (1 + x (1 - s)), s = sigmoid(x)
__global__ void __launch_bounds__(256)
silu_bwd_f32_kernel(const float* __restrict__ pre, const float* __restrict__ dy, float* __restrict__ out,
                    long long n) {
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = pre[i];
    const float s = 1.0f / (1.0f + __expf(-x));
    out[i] = dy[i] * s * (1.0f + x * (1.0f - s));
  }
}

// in [rows, cols] (bf16 or fp32, row stride ld_in) -> out bf16 [cols, ld_out], out[c][r] = in[r][c];
// columns r in [rows, ld_out) are zero filled (the transposed operand is the K-major input of a
// weight-gradient GEMM whose K = rows must be a multiple of 8).
template <typename T>
__global__ void __launch_bounds__(256)
transpose_to_bf16_kernel(const T* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
                         long long ld_out, int rows, int cols) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) v = static_cast<float>(in[static_cast<long long>(r) * ld_in + c]);
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (c < cols && r < ld_out) out[static_cast<long long>(c) * ld_out + r] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
  }
}

// out[c] += sum_r x[r][c]  (bias gradients).  grid (ceil(c/64), row splits); out must be zeroed.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out,
                   long long rows, int c) {
  pdl_wait();
  __shared__ float s_acc[8][64];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int ch = blockIdx.x * 64 + cx * 2;
  float a0 = 0.f, a1 = 0.f;
  if (ch < c) {
    for (long long r = static_cast<long long>(blockIdx.y) * 8 + ry; r < rows;
         r += static_cast<long long>(gridDim.y) * 8) {
      const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(x + r * ld + ch)));
      a0 += v.x;
      a1 += v.y;
    }
  }
  s_acc[ry][cx * 2] = a0;
  s_acc[ry][cx * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_acc[i][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < c) atomicAdd(out + cc, t);
  }
}

// Backward of the stride-2 pad-1 3x3 im2col (Downsample, openaimodel.py:215-222):
// dx[b, y, x, c] = sum over taps (ky, kx) with (y+1-ky), (x+1-kx) even and in range of
// dcol[b, (y+1-ky)/2, (x+1-kx)/2, (ky*3+kx)*C + c]
__global__ void __launch_bounds__(256)
col2im3x3_s2_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx,
                    int batch, int h, int w, int c) {
  pdl_wait();
  const int nvec = c >> 3;
  const int ho = h >> 1, wo = w >> 1;
  const long long total = static_cast<long long>(batch) * h * w * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long pix = i / nvec;
    const int x = static_cast<int>(pix % w);
    pix /= w;
    const int y = static_cast<int>(pix % h);
    const int b = static_cast<int>(pix / h);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1) || (ty >> 1) >= ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1) || (tx >> 1) >= wo) continue;
        const long long orow = (static_cast<long long>(b) * ho + (ty >> 1)) * wo + (tx >> 1);
        float f[8];
        unpack8f(__ldg(reinterpret_cast<const uint4*>(dcol + orow * 9 * c + (ky * 3 + kx) * c + v * 8)), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
    }
    *reinterpret_cast<uint4*>(dx + ((static_cast<long long>(b) * h + y) * w + x) * c + v * 8) = pack8f(acc);
  }
}

// Backward of nearest x2 upsampling (Upsample.forward, openaimodel.py:161): sum of the 2x2 block.
// g: [B, 2h, 2w, C] -> dx [B, h, w, C]
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dx, int batch,
                      int h, int w, int c) {
  pdl_wait();
  const int nvec = c >> 3;
  const long long total = static_cast<long long>(batch) * h * w * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long pix = i / nvec;
    const int x = static_cast<int>(pix % w);
    pix /= w;
    const int y = static_cast<int>(pix % h);
    const int b = static_cast<int>(pix / h);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        const long long row = (static_cast<long long>(b) * 2 * h + 2 * y + dy) * 2 * w + 2 * x + dxx;
        float f[8];
        unpack8f(__ldg(reinterpret_cast<const uint4*>(g + row * c + v * 8)), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8f(acc);
  }
}

// =================================================================================================
// FeatureNeRF backward
// =================================================================================================
// Volume rendering backward (VolRender.forward / get_weights, nerfsd_pytorch3d.py:170-231, with the
// trunc_exp / sigmoid activations applied in reference_attn, attention.py:590-594; _TruncExp
// backward clamps the exponent at 15, attention.py:201-205).  One warp per (b, ray).
//   w_s = alpha_s exp(-E_s), alpha_s = 1 - exp(-dd_s), E_s = sum_{t<s} dd_t, dd = dist * exp(raw_sigma)
//   dL/dw_s = <d_rendered, f_s> + dfg + <drgb, rgb_s>
//   dL/ddd_s = (dL/dw_s exp(-E_s) + dalpha_s) exp(-dd_s) - sum_{t>s} dL/dw_t w_t
// Outputs: dfeats bf16 [b,hw,d,c] = w_s d_rendered;  draw bf16 [b,hw,d,8] = (drgb_raw 3, dsigma_raw 1, 0 x4)
constexpr int VR_MAXD = 32;

__global__ void __launch_bounds__(256)
nerf_volrender_bwd_kernel(const __nv_bfloat16* __restrict__ feats, const float* __restrict__ raw,
                          const float* __restrict__ dists, const __nv_bfloat16* __restrict__ d_rendered,
                          const float* __restrict__ dfg, const float* __restrict__ dalphas,
                          const float* __restrict__ drgb, __nv_bfloat16* __restrict__ dfeats,
                          __nv_bfloat16* __restrict__ draw, int nb, int hw, int d, int c) {
  pdl_wait();
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<long long>(nb) * hw) return;
  const int ray = static_cast<int>(wid % hw);
  float dd = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, sig_raw = 0.f, dist = 0.f;
  if (lane < d) {
    const float4 rw = *reinterpret_cast<const float4*>(raw + (wid * d + lane) * 4);
    sig_raw = rw.w;
    dist = dists[ray * d + lane];
    dd = dist * expf(rw.w);
    r0 = sigmoid_f(rw.x);
    r1 = sigmoid_f(rw.y);
    r2 = sigmoid_f(rw.z);
  }
  float incl = dd;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const float trans = expf(-(incl - dd));
  const float e_dd = expf(-dd);
  float w = (1.f - e_dd) * trans;
  const bool w_bad = (w != w) || isinf(w);  // nan_to_num: no gradient through replaced values
  if (w != w) w = 0.f;
  else if (isinf(w)) w = w > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  if (lane >= d) w = 0.f;
  // per-sample <d_rendered, f_s> accumulated over this lane's channels, and dfeats = w_s d_rendered
  float dot[VR_MAXD];
#pragma unroll
  for (int s = 0; s < VR_MAXD; ++s) dot[s] = 0.f;
  const int nvec = c >> 3;
  // every lane runs every iteration (the shuffles below need the whole warp); `act` guards memory
  for (int v0 = 0; v0 < nvec; v0 += 32) {
    const int vi = v0 + lane;
    const bool act = vi < nvec;
    float g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = 0.f;
    if (act) unpack8f(__ldg(reinterpret_cast<const uint4*>(d_rendered + wid * c + vi * 8)), g);
#pragma unroll
    for (int s = 0; s < VR_MAXD; ++s) {
      if (s < d) {
        const float ws = __shfl_sync(0xffffffffu, w, s);
        if (act) {
          float f[8], o[8];
          unpack8f(__ldg(reinterpret_cast<const uint4*>(feats + (wid * d + s) * c + vi * 8)), f);
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            acc = fmaf(g[k], f[k], acc);
            o[k] = ws * g[k];
          }
          dot[s] += acc;
          *reinterpret_cast<uint4*>(dfeats + (wid * d + s) * c + vi * 8) = pack8f(o);
        }
      }
    }
  }
  float dw = 0.f;
#pragma unroll
  for (int s = 0; s < VR_MAXD; ++s) {
    float t = dot[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == s) dw = t;
  }
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (drgb != nullptr) {
    g0 = drgb[wid * 3 + 0];
    g1 = drgb[wid * 3 + 1];
    g2 = drgb[wid * 3 + 2];
  }
  dw += (dfg != nullptr ? dfg[wid] : 0.f) + g0 * r0 + g1 * r1 + g2 * r2;
  if (w_bad || lane >= d) dw = 0.f;
  // suffix sum over t > s of dw_t w_t
  const float dww = dw * w;
  float suf = dww;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += t;
  }
  suf -= dww;
  const float dal = (dalphas != nullptr && lane < d) ? dalphas[wid * d + lane] : 0.f;
  const float ddd = (dw * trans + dal) * e_dd - suf;
  if (lane < d) {
    const float dsig = ddd * dist * expf(fminf(sig_raw, 15.f));
    uint4 o;
    o.x = pack_bf16x2(w * g0 * r0 * (1.f - r0), w * g1 * r1 * (1.f - r1));
    o.y = pack_bf16x2(w * g2 * r2 * (1.f - r2), dsig);
    o.z = 0u;
    o.w = 0u;
    *reinterpret_cast<uint4*>(draw + (wid * d + lane) * 8) = o;
  }
}

// Backward of cd360_nerf_combine (nerf.cu): one warp per (b, ray*d + sample).
//   h_v = hpre_v + gather(G_v)[:c], s_v = silu(h_v), logit_v = vlogit_v + gather(G_v)[c], a = softmax_v
//   S = sum_v a_v s_v
// given dS: dhpre_v = a_v dS silu'(h_v); da_v = <dS, s_v>; dlogit_v = a_v (da_v - sum_u a_u da_u);
// dG (fp32, atomics through the 4 bilinear taps) receives dhpre_v in columns [0,c) and dlogit_v in
// column c.  grid_sample's own backward towards the feature map (zeros padding, align_corners) is
// exactly this scatter; there is no gradient towards the sampling grid (cameras are data).
constexpr int NB_MAX_VIEWS = 16;

__global__ void __launch_bounds__(256)
nerf_combine_bwd_kernel(const __nv_bfloat16* __restrict__ g, long long ldg,
                        const __nv_bfloat16* __restrict__ hpre, const int* __restrict__ gidx,
                        const float* __restrict__ gwgt, const float* __restrict__ vlogit,
                        const __nv_bfloat16* __restrict__ ds_in, __nv_bfloat16* __restrict__ dhpre,
                        float* __restrict__ dlogit, float* __restrict__ dg, int nb, int n, int hw,
                        int d, int c) {
  pdl_wait();
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long npts = static_cast<long long>(nb) * hw * d;
  if (wid >= npts) return;
  const int b = static_cast<int>(wid / (static_cast<long long>(hw) * d));
  const long long p = wid - static_cast<long long>(b) * hw * d;

  float logit = -INFINITY;
  int my_idx[4] = {-1, -1, -1, -1};
  float my_w[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane < n) {
    const long long point = ((static_cast<long long>(b) * n + lane) * hw) * d + p;
    const int4 gi = *reinterpret_cast<const int4*>(gidx + point * 4);
    const float4 gw = *reinterpret_cast<const float4*>(gwgt + point * 4);
    my_idx[0] = gi.x; my_idx[1] = gi.y; my_idx[2] = gi.z; my_idx[3] = gi.w;
    my_w[0] = gw.x; my_w[1] = gw.y; my_w[2] = gw.z; my_w[3] = gw.w;
    float acc = vlogit[point];
    const __nv_bfloat16* gb = g + (static_cast<long long>(b) * n + lane) * hw * ldg;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (my_idx[k] >= 0) acc += my_w[k] * __bfloat162float(gb[my_idx[k] * ldg + c]);
    logit = acc;
  }
  float mx = logit;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float e = (lane < n) ? expf(logit - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float a_mine = e / sum;

  float da[NB_MAX_VIEWS];
#pragma unroll
  for (int v = 0; v < NB_MAX_VIEWS; ++v) da[v] = 0.f;
  const int nvec = c >> 3;
  // every lane runs every iteration (the shuffles below need the whole warp); `act` guards memory
  for (int v0 = 0; v0 < nvec; v0 += 32) {
    const int vi = v0 + lane;
    const bool act = vi < nvec;
    float dsv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dsv[k] = 0.f;
    if (act) unpack8f(__ldg(reinterpret_cast<const uint4*>(ds_in + wid * c + vi * 8)), dsv);
#pragma unroll
    for (int v = 0; v < NB_MAX_VIEWS; ++v) {
      if (v < n) {
        const float a_v = __shfl_sync(0xffffffffu, a_mine, v);
        int idx[4];
        float wk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          idx[k] = __shfl_sync(0xffffffffu, my_idx[k], v);
          wk[k] = __shfl_sync(0xffffffffu, my_w[k], v);
        }
        if (!act) continue;
        const long long point = ((static_cast<long long>(b) * n + v) * hw) * d + p;
        float h[8];
        unpack8f(__ldg(reinterpret_cast<const uint4*>(hpre + point * c + vi * 8)), h);
        const long long gbase = (static_cast<long long>(b) * n + v) * hw;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (idx[k] >= 0) {
            float f[8];
            unpack8f(__ldg(reinterpret_cast<const uint4*>(g + (gbase + idx[k]) * ldg + vi * 8)), f);
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = fmaf(wk[k], f[q], h[q]);
          }
        float dp[8];
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float sg = sigmoid_f(h[q]);
          acc = fmaf(dsv[q], h[q] * sg, acc);                            // <dS, silu(h)>
          dp[q] = a_v * dsv[q] * sg * (1.0f + h[q] * (1.0f - sg));       // a_v dS silu'(h)
        }
        da[v] += acc;
        *reinterpret_cast<uint4*>(dhpre + point * c + vi * 8) = pack8f(dp);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (idx[k] >= 0) {
            float* dst = dg + (gbase + idx[k]) * ldg + vi * 8;
            atomicAdd(reinterpret_cast<float4*>(dst),
                      make_float4(wk[k] * dp[0], wk[k] * dp[1], wk[k] * dp[2], wk[k] * dp[3]));
            atomicAdd(reinterpret_cast<float4*>(dst + 4),
                      make_float4(wk[k] * dp[4], wk[k] * dp[5], wk[k] * dp[6], wk[k] * dp[7]));
          }
      }
    }
  }
  float da_mine = 0.f;
#pragma unroll
  for (int v = 0; v < NB_MAX_VIEWS; ++v) {
    float t = da[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == v) da_mine = t;
  }
  float mixed = (lane < n) ? a_mine * da_mine : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mixed += __shfl_xor_sync(0xffffffffu, mixed, o);
  if (lane < n) {
    const float dl = a_mine * (da_mine - mixed);
    const long long point = ((static_cast<long long>(b) * n + lane) * hw) * d + p;
    dlogit[point] = dl;
    const long long gbase = (static_cast<long long>(b) * n + lane) * hw;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (my_idx[k] >= 0) atomicAdd(dg + (gbase + my_idx[k]) * ldg + c, my_w[k] * dl);
  }
}

// Geometry part of the nviews weight gradient.  The columns of nviews.weight that multiply
// quantities shared by all views of a point (PE16(p_target) 96 | p_target 3) and the bias receive
// sum_v dlogit_v = 0 exactly (softmax); the per-view columns (origin of the reference camera in
// the target frame 3 | its PE16 96, nerfsd_pytorch3d.py:116-123,139-151) do not depend on the
// point, so their gradient is feat(b, v) * sum_p dlogit[b, v, p].  One CTA per (v, b);
// dw [198] must be zeroed by the caller.
__device__ __forceinline__ void geo_cam_center_in_target(const float* __restrict__ cams, int b, int n,
                                                         int v, float (&ot)[3]) {
  const float* tg = cams + static_cast<long long>(b) * (n + 1) * 16;
  const float* rc = tg + (1 + v) * 16;
  float oc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) oc[k] = -(rc[9] * rc[k * 3 + 0] + rc[10] * rc[k * 3 + 1] + rc[11] * rc[k * 3 + 2]);
#pragma unroll
  for (int j = 0; j < 3; ++j) ot[j] = oc[0] * tg[0 * 3 + j] + oc[1] * tg[1 * 3 + j] + oc[2] * tg[2 * 3 + j] + tg[9 + j];
}

__global__ void __launch_bounds__(256)
nerf_nviews_geo_bwd_kernel(const float* __restrict__ cams, const float* __restrict__ dlogit,
                           float* __restrict__ dw, int n, long long pts) {
  pdl_wait();
  __shared__ float s_red[32];
  const int v = blockIdx.x, b = blockIdx.y;
  const float* src = dlogit + (static_cast<long long>(b) * n + v) * pts;
  float s = 0.f;
  for (long long i = threadIdx.x; i < pts; i += blockDim.x) s += src[i];
  const float total = block_sum(s, s_red);
  float ot[3];
  geo_cam_center_in_target(cams, b, n, v, ot);
  const float kPiF = 3.14159274101257324f;
  for (int j = threadIdx.x; j < 99; j += blockDim.x) {
    float feat;
    if (j < 3) {
      feat = ot[j];
    } else {
      const int e = j - 3;           // PE16 layout: [sin band0 (3) ... sin band15 | cos band0 ...]
      const int is_cos = e >= 48;
      const int r = e - 48 * is_cos;
      const int fb = r / 3, k = r - fb * 3;
      const float arg = ot[k] * (exp2f(static_cast<float>(fb - 8)) * kPiF);
      feat = is_cos ? cosf(arg) : sinf(arg);
    }
    atomicAdd(dw + 99 + j, feat * total);
  }
}

// =================================================================================================
// Loss (StandardDiffusionLossImgRef.get_loss, loss.py:173-216, 'l2' branch) and its gradient.
// =================================================================================================
// Denoising term.  eps: UNet output, fp32 NHWC tokens [b*hw, 4]; x_noisy / target: fp32 NCHW
// [b, 4, hw]; model_output = x_noisy - sigma eps (EpsScaling c_out = -sigma, c_skip = 1,
// denoiser.py:44); w = sigma^-2 (EpsWeighting); loss_b = sum(w (mo - target)^2 mask) / (sum mask +
// 1e-6)  [mask: fp32 [b, hw] at latent resolution] or the plain mean when mask is NULL.
// Gradient of coef * sum_b loss_b w.r.t. eps, written as bf16 tokens [b*hw, ldd] (columns >= 4
// zero: the K-padded operand of the output convolution's data-gradient GEMM).  One CTA per image.
__global__ void __launch_bounds__(512)
diffusion_loss_kernel(const float* __restrict__ eps, const float* __restrict__ x_noisy,
                      const float* __restrict__ target, const float* __restrict__ sigma,
                      const float* __restrict__ mask, float coef, float* __restrict__ loss_out,
                      float* __restrict__ mask_sum_out, __nv_bfloat16* __restrict__ deps, int hw,
                      int ldd) {
  pdl_wait();
  __shared__ float s_red[32];
  const int b = blockIdx.x;
  const float sg = sigma[b];
  const float wgt = 1.0f / (sg * sg);
  float ms = 0.f;
  if (mask != nullptr)
    for (int i = threadIdx.x; i < hw; i += blockDim.x) ms += mask[static_cast<long long>(b) * hw + i];
  const float msum = mask != nullptr ? block_sum(ms, s_red) : 0.f;
  const float denom = mask != nullptr ? (msum + 1e-6f) : static_cast<float>(4 * hw);
  float acc = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const long long row = static_cast<long long>(b) * hw + i;
    const float4 e = *reinterpret_cast<const float4*>(eps + row * 4);
    const float ev[4] = {e.x, e.y, e.z, e.w};
    const float mk = mask != nullptr ? mask[row] : 1.0f;
    float gq[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const long long idx = (static_cast<long long>(b) * 4 + ch) * hw + i;
      const float diff = x_noisy[idx] - sg * ev[ch] - target[idx];
      acc = fmaf(wgt * diff * diff, mk, acc);
      gq[ch] = coef * 2.0f * wgt * diff * (-sg) * mk / denom;
    }
    __nv_bfloat16* drow = deps + row * ldd;
    *reinterpret_cast<uint2*>(drow) = make_uint2(pack_bf16x2(gq[0], gq[1]), pack_bf16x2(gq[2], gq[3]));
    for (int k = 4; k < ldd; k += 4) *reinterpret_cast<uint2*>(drow + k) = make_uint2(0u, 0u);
  }
  const float tot = block_sum(acc, s_red);
  if (threadIdx.x == 0) {
    loss_out[b] = tot / denom;
    if (mask_sum_out != nullptr) mask_sum_out[b] = msum;
  }
}

// FeatureNeRF supervision of one pose block (loss.py:183-206):
//   loss_fg_b  = mean_hw (clamp(fg, 0, 1) - op)^2
//   loss_bg_b  = mean_{hw,d} |alpha - op| (1 - op) [op < 0.1]
//   loss_rgb_b = sum_{3,hw} (tgt - rgb)^2 mask_s / (mask_sum_b + 1e-6)
// op [b, hw], mask_s [b, hw], tgt [b, 3, hw] are the (antialias-resized) supervision maps.
// wfg / wbg / wrgb: fp32 [b] weights of each term in the total (lambda * drop_im_b / (K (sum drop_im
// + 1e-12)), diffusion.py:221-236); gradients are written pre-multiplied by them.
__global__ void __launch_bounds__(256)
nerf_aux_loss_kernel(const float* __restrict__ fg, const float* __restrict__ alphas,
                     const float* __restrict__ rgb, const float* __restrict__ op,
                     const float* __restrict__ mask_s, const float* __restrict__ tgt,
                     const float* __restrict__ mask_sum, const float* __restrict__ wfg,
                     const float* __restrict__ wbg, const float* __restrict__ wrgb,
                     float* __restrict__ loss3, float* __restrict__ dfg, float* __restrict__ dalphas,
                     float* __restrict__ drgb, int hw, int d) {
  pdl_wait();
  __shared__ float s_red[32];
  const int b = blockIdx.x;
  const float w_fg = wfg[b], w_bg = wbg[b];
  const float w_rgb = rgb != nullptr ? wrgb[b] : 0.f;
  const float rden = rgb != nullptr ? 1.0f / (mask_sum[b] + 1e-6f) : 0.f;
  float lf = 0.f, lb = 0.f, lr = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const long long row = static_cast<long long>(b) * hw + i;
    const float o = op[row];
    const float f = fg[row];
    const float fc = fminf(fmaxf(f, 0.f), 1.f);
    const float df = fc - o;
    lf = fmaf(df, df, lf);
    dfg[row] = (f >= 0.f && f <= 1.f) ? w_fg * 2.f * df / hw : 0.f;
    const float bgw = (o < 0.1f) ? (1.f - o) : 0.f;
    for (int s = 0; s < d; ++s) {
      const float a = alphas[row * d + s];
      const float t = a - o;
      lb = fmaf(fabsf(t), bgw, lb);
      dalphas[row * d + s] = w_bg * (t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f)) * bgw / (static_cast<float>(hw) * d);
    }
    if (rgb != nullptr) {
      const float mk = mask_s[row];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float t = tgt[(static_cast<long long>(b) * 3 + k) * hw + i] - rgb[row * 3 + k];
        lr = fmaf(t * t, mk, lr);
        drgb[row * 3 + k] = w_rgb * (-2.f) * t * mk * rden;
      }
    }
  }
  lf = block_sum(lf, s_red);
  lb = block_sum(lb, s_red);
  lr = block_sum(lr, s_red);
  if (threadIdx.x == 0) {
    loss3[b * 3 + 0] = lf / hw;
    loss3[b * 3 + 1] = lb / (static_cast<float>(hw) * d);
    loss3[b * 3 + 2] = lr * rden;
  }
}

// torch.nn.functional.interpolate(mode='bilinear', antialias=True, align_corners=False) as used on
// the supervision maps (loss.py:186,199-200): separable triangle filter whose support grows with
// the downscale factor (ATen _upsample_bilinear2d_aa).  in [planes, ih, iw] -> out [planes, oh, ow];
// out = scale * resized + shift.  One thread per output element.
__device__ __forceinline__ void aa_window(int o, float scale, int in_size, int& lo, int& size,
                                          float& center, float& invscale, float& support) {
  support = scale >= 1.0f ? scale : 1.0f;         // interp_size / 2 * scale, interp_size = 2
  invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  center = scale * (static_cast<float>(o) + 0.5f);
  lo = max(static_cast<int>(center - support + 0.5f), 0);
  size = min(static_cast<int>(center + support + 0.5f), in_size) - lo;
}
// One WARP per output pixel: the window of a strong down-scale is large (512 -> 16: 64 x 64 taps) and
// there are few outputs (3 planes x 256), so a thread per output left 3 CTAs working for 0.5 ms.
// Lanes stride over the flattened window (x fastest: coalesced rows), fixed-order tree reduction.
__global__ void __launch_bounds__(256)
resize_bilinear_aa_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int ih,
                          int iw, int oh, int ow, float out_scale, float out_shift) {
  pdl_wait();
  const long long total = static_cast<long long>(planes) * oh * ow;
  const float sy = static_cast<float>(ih) / oh, sx = static_cast<float>(iw) / ow;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long i = warp0; i < total; i += nwarps) {
    const int ox = static_cast<int>(i % ow);
    const int oy = static_cast<int>((i / ow) % oh);
    const long long pl = i / (static_cast<long long>(ow) * oh);
    int ylo, ysz, xlo, xsz;
    float yc, yinv, ysup, xc, xinv, xsup;
    aa_window(oy, sy, ih, ylo, ysz, yc, yinv, ysup);
    aa_window(ox, sx, iw, xlo, xsz, xc, xinv, xsup);
    float wys = 0.f, wxs = 0.f;
    for (int j = 0; j < ysz; ++j) wys += fmaxf(0.f, 1.f - fabsf((j + ylo - yc + 0.5f) * yinv));
    for (int j = 0; j < xsz; ++j) wxs += fmaxf(0.f, 1.f - fabsf((j + xlo - xc + 0.5f) * xinv));
    const float* base = in + (pl * ih + ylo) * iw + xlo;
    const int taps = ysz * xsz;
    float acc = 0.f;
    for (int t = lane; t < taps; t += 32) {
      const int jy = t / xsz, jx = t - jy * xsz;
      const float wy = fmaxf(0.f, 1.f - fabsf((jy + ylo - yc + 0.5f) * yinv));
      const float wx = fmaxf(0.f, 1.f - fabsf((jx + xlo - xc + 0.5f) * xinv));
      acc = fmaf(wy * wx, __ldg(base + static_cast<long long>(jy) * iw + jx), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[i] = fmaf(out_scale, acc / (wys * wxs), out_shift);
  }
}

// =================================================================================================
// AdamW (torch.optim.AdamW semantics, the reference's default optimizer_config; diffusion.py:310-373)
// over flat fp32 buffers: p -= lr wd p; m, v updated; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// =================================================================================================
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
             float weight_decay, float bc1, float bc2_sqrt, float grad_scale) {
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.0f - lr * weight_decay);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
  }
}

static inline unsigned grid_for(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(kNumSMsB200) * 8;  // grid-stride loops cover the rest
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

}  // namespace cd360

using namespace cd360;

#define CD360_BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define CD360_BFW(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CD360_MISALIGNED(p) (reinterpret_cast<uintptr_t>(p) & 15)

extern "C" int cd360_layernorm_bwd_bf16(const void* x, const float* gamma, const void* dy,
                                        const void* add, void* dx, int32_t rows, int32_t c, float eps,
                                        cd360_stream_t stream_) {
  if (!x || !gamma || !dy || !dx) return CD360_ERR_NULL;
  if (rows <= 0 || c <= 0 || (c & 7) || c > LNB_MAXV * 256) return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(x) || CD360_MISALIGNED(dy) || CD360_MISALIGNED(dx) || CD360_MISALIGNED(gamma) ||
      (add && CD360_MISALIGNED(add)))
    return CD360_ERR_ALIGN;
  const int blocks = (rows + 7) / 8;
  launch_ex(layernorm_bwd_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1,
            CD360_BF(x), gamma, CD360_BF(dy), CD360_BF(add), CD360_BFW(dx), rows, c, eps);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int64_t cd360_groupnorm_bwd_workspace_floats(int32_t batch) {
  return static_cast<int64_t>(batch) * GNB_GROUPS * 4;
}

extern "C" int cd360_groupnorm_silu_bwd_bf16(const void* x0, int32_t c0, const void* x1, int32_t c1,
                                             const float* gamma, const float* beta, const void* dy,
                                             const void* add0, int64_t ld_add0, const void* add1,
                                             int64_t ld_add1, void* dx0, void* dx1, float* workspace,
                                             int32_t batch, int32_t hw, float eps, int32_t apply_silu,
                                             cd360_stream_t stream_) {
  if (!x0 || !gamma || !beta || !dy || !dx0 || !workspace) return CD360_ERR_NULL;
  if (c1 > 0 && (!x1 || !dx1)) return CD360_ERR_NULL;
  if (c1 < 0 || c0 <= 0 || batch <= 0 || hw <= 0 || batch > 65535) return CD360_ERR_SHAPE;
  const int ctot = c0 + c1;
  if ((c0 & 7) || (c1 & 7) || (ctot % GNB_GROUPS) != 0 || ((ctot / GNB_GROUPS) & 1)) return CD360_ERR_SHAPE;
  if ((add0 && (ld_add0 & 1)) || (add1 && (ld_add1 & 1))) return CD360_ERR_ALIGN;
  if (CD360_MISALIGNED(x0) || CD360_MISALIGNED(dy) || CD360_MISALIGNED(dx0) || CD360_MISALIGNED(workspace) ||
      (x1 && CD360_MISALIGNED(x1)) || (dx1 && CD360_MISALIGNED(dx1)) ||
      (add0 && (reinterpret_cast<uintptr_t>(add0) & 3)) || (add1 && (reinterpret_cast<uintptr_t>(add1) & 3)))
    return CD360_ERR_ALIGN;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  // rows of one (group, image) divided among a cluster of up to 8 CTAs (>= 64 rows each)
  static int split_ok = -1;
  if (split_ok < 0) {
    const char* e = getenv("CD360_GNB_CLUSTER");
    split_ok = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  int splits = 1;
  if (split_ok) while (splits < 8 && hw / (splits * 2) >= 64 && GNB_GROUPS * batch * splits * 2 <= 4 * kNumSMsB200) splits *= 2;
  launch_ex(groupnorm_bwd_stats_kernel, dim3(GNB_GROUPS * splits, batch), dim3(splits > 1 ? 256 : 512), 0, stream,
            splits, CD360_BF(x0), c0, CD360_BF(x1), c1, gamma, beta, CD360_BF(dy), workspace, hw, eps, apply_silu,
            splits);
  CD360_CHECK_LAUNCH();
  const long long items = static_cast<long long>(batch) * hw * (ctot / 2);
  launch_ex(groupnorm_bwd_apply_kernel, dim3(grid_for(items, 256)), dim3(256), 0, stream, 1, CD360_BF(x0),
            c0, CD360_BF(x1), c1, gamma, beta, CD360_BF(dy), CD360_BF(add0),
            static_cast<long long>(ld_add0), CD360_BF(add1), static_cast<long long>(ld_add1),
            CD360_BFW(dx0), CD360_BFW(dx1), static_cast<const float*>(workspace), batch, hw, apply_silu);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_geglu_bwd_bf16(const void* raw, const void* dh, void* draw, int64_t rows,
                                    int32_t f, int32_t block, cd360_stream_t stream_) {
  if (!raw || !dh || !draw) return CD360_ERR_NULL;
  if (rows <= 0 || f <= 0 || (f & 7) || block <= 0 || (block & 7) || (f % block) != 0) return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(raw) || CD360_MISALIGNED(dh) || CD360_MISALIGNED(draw)) return CD360_ERR_ALIGN;
  launch_ex(geglu_bwd_kernel, dim3(grid_for(rows * (f / 8), 256)), dim3(256), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, CD360_BF(raw), CD360_BF(dh), CD360_BFW(draw),
            static_cast<long long>(rows), f, block);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_geglu_fwd_bf16(const void* raw, void* h, int64_t rows, int32_t f, int32_t block,
                                    cd360_stream_t stream_) {
  if (!raw || !h) return CD360_ERR_NULL;
  if (rows <= 0 || f <= 0 || (f & 7) || block <= 0 || (block & 7) || (f % block) != 0) return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(raw) || CD360_MISALIGNED(h)) return CD360_ERR_ALIGN;
  launch_ex(geglu_fwd_kernel, dim3(grid_for(rows * (f / 8), 256)), dim3(256), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, CD360_BF(raw), CD360_BFW(h), static_cast<long long>(rows), f,
            block);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_add_bf16(const void* a, const void* b, void* out, int64_t n,
                              cd360_stream_t stream_) {
  if (!a || !b || !out) return CD360_ERR_NULL;
  if (n <= 0 || (n & 7)) return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(a) || CD360_MISALIGNED(b) || CD360_MISALIGNED(out)) return CD360_ERR_ALIGN;
  launch_ex(add_bf16_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_),
            1, CD360_BF(a), CD360_BF(b), CD360_BFW(out), static_cast<long long>(n / 8));
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_silu_bwd_f32(const float* pre, const float* dy, float* out, int64_t n,
                                  cd360_stream_t stream_) {
  if (!pre || !dy || !out) return CD360_ERR_NULL;
  if (n <= 0) return CD360_ERR_SHAPE;
  launch_ex(silu_bwd_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1,
            pre, dy, out, static_cast<long long>(n));
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_transpose_to_bf16(const void* in, int32_t in_is_fp32, int64_t ld_in, void* out,
                                       int64_t ld_out, int32_t rows, int32_t cols,
                                       cd360_stream_t stream_) {
  if (!in || !out) return CD360_ERR_NULL;
  if (rows <= 0 || cols <= 0 || ld_in < cols || ld_out < rows) return CD360_ERR_SHAPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const dim3 grid((cols + 31) / 32, static_cast<unsigned>((ld_out + 31) / 32));
  if (grid.y > 65535) return CD360_ERR_SHAPE;
  if (in_is_fp32)
    launch_ex(transpose_to_bf16_kernel<float>, grid, dim3(256), 0, stream, 1, static_cast<const float*>(in),
              static_cast<long long>(ld_in), CD360_BFW(out), static_cast<long long>(ld_out), rows, cols);
  else
    launch_ex(transpose_to_bf16_kernel<__nv_bfloat16>, grid, dim3(256), 0, stream, 1, CD360_BF(in),
              static_cast<long long>(ld_in), CD360_BFW(out), static_cast<long long>(ld_out), rows, cols);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_colsum_bf16(const void* x, int64_t ld, float* out, int64_t rows, int32_t c,
                                 cd360_stream_t stream_) {
  if (!x || !out) return CD360_ERR_NULL;
  if (rows <= 0 || c <= 0 || (c & 1) || (ld & 1) || ld < c) return CD360_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(x) & 3) return CD360_ERR_ALIGN;
  long long splits = (rows + 255) / 256;
  if (splits > 64) splits = 64;
  launch_ex(colsum_bf16_kernel, dim3((c + 63) / 64, static_cast<unsigned>(splits)), dim3(256), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, CD360_BF(x), static_cast<long long>(ld), out,
            static_cast<long long>(rows), c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_col2im3x3_s2_bf16(const void* dcol, void* dx, int32_t batch, int32_t h, int32_t w,
                                       int32_t c, cd360_stream_t stream_) {
  if (!dcol || !dx) return CD360_ERR_NULL;
  if (batch <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || c <= 0 || (c & 7)) return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(dcol) || CD360_MISALIGNED(dx)) return CD360_ERR_ALIGN;
  const long long items = static_cast<long long>(batch) * h * w * (c / 8);
  launch_ex(col2im3x3_s2_kernel, dim3(grid_for(items, 256)), dim3(256), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, CD360_BF(dcol), CD360_BFW(dx), batch, h, w, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_upsample_nearest2x_bwd_bf16(const void* g, void* dx, int32_t batch, int32_t h,
                                                 int32_t w, int32_t c, cd360_stream_t stream_) {
  if (!g || !dx) return CD360_ERR_NULL;
  if (batch <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 7)) return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(g) || CD360_MISALIGNED(dx)) return CD360_ERR_ALIGN;
  const long long items = static_cast<long long>(batch) * h * w * (c / 8);
  launch_ex(upsample2x_bwd_kernel, dim3(grid_for(items, 256)), dim3(256), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, CD360_BF(g), CD360_BFW(dx), batch, h, w, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nerf_volrender_bwd(const void* feats, const float* raw, const float* dists,
                                        const void* d_rendered, const float* dfg, const float* dalphas,
                                        const float* drgb, void* dfeats, void* draw, int32_t b,
                                        int32_t hw, int32_t d, int32_t c, cd360_stream_t stream_) {
  if (!feats || !raw || !dists || !d_rendered || !dfeats || !draw) return CD360_ERR_NULL;
  if (b <= 0 || hw <= 0 || d <= 0 || d > VR_MAXD || c <= 0 || (c & 7)) return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(feats) || CD360_MISALIGNED(raw) || CD360_MISALIGNED(d_rendered) ||
      CD360_MISALIGNED(dfeats) || CD360_MISALIGNED(draw))
    return CD360_ERR_ALIGN;
  const long long warps = static_cast<long long>(b) * hw;
  launch_ex(nerf_volrender_bwd_kernel, dim3(static_cast<unsigned>((warps + 7) / 8)), dim3(256), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, CD360_BF(feats), raw, dists, CD360_BF(d_rendered), dfg,
            dalphas, drgb, CD360_BFW(dfeats), CD360_BFW(draw), b, hw, d, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nerf_combine_bwd(const void* g, int64_t ldg, const void* hpre,
                                      const int32_t* gidx, const float* gwgt, const float* vlogit,
                                      const void* ds, void* dhpre, float* dlogit, float* dg, int32_t b,
                                      int32_t n, int32_t hw, int32_t d, int32_t c,
                                      cd360_stream_t stream_) {
  if (!g || !hpre || !gidx || !gwgt || !vlogit || !ds || !dhpre || !dlogit || !dg) return CD360_ERR_NULL;
  if (b <= 0 || n <= 0 || n > NB_MAX_VIEWS || hw <= 0 || d <= 0 || c <= 0 || (c & 7) || ldg < c + 1 ||
      (ldg & 7))
    return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(g) || CD360_MISALIGNED(hpre) || CD360_MISALIGNED(ds) || CD360_MISALIGNED(dhpre) ||
      CD360_MISALIGNED(dg) || CD360_MISALIGNED(gidx) || CD360_MISALIGNED(gwgt))
    return CD360_ERR_ALIGN;
  const long long warps = static_cast<long long>(b) * hw * d;
  launch_ex(nerf_combine_bwd_kernel, dim3(static_cast<unsigned>((warps + 7) / 8)), dim3(256), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, CD360_BF(g), static_cast<long long>(ldg),
            CD360_BF(hpre), gidx, gwgt, vlogit, CD360_BF(ds), CD360_BFW(dhpre), dlogit, dg, b, n, hw, d, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nerf_nviews_geo_bwd(const float* cams, const float* dlogit, float* dw, int32_t b,
                                         int32_t n, int64_t pts, cd360_stream_t stream_) {
  if (!cams || !dlogit || !dw) return CD360_ERR_NULL;
  if (b <= 0 || n <= 0 || pts <= 0 || b > 65535) return CD360_ERR_SHAPE;
  launch_ex(nerf_nviews_geo_bwd_kernel, dim3(n, b), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1,
            cams, dlogit, dw, n, static_cast<long long>(pts));
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_diffusion_loss(const float* eps, const float* x_noisy, const float* target,
                                    const float* sigma, const float* mask, float coef, float* loss,
                                    float* mask_sum, void* deps, int32_t batch, int32_t hw, int32_t ldd,
                                    cd360_stream_t stream_) {
  if (!eps || !x_noisy || !target || !sigma || !loss || !deps) return CD360_ERR_NULL;
  if (batch <= 0 || hw <= 0 || ldd < 4 || (ldd & 3)) return CD360_ERR_SHAPE;
  if (CD360_MISALIGNED(eps) || (reinterpret_cast<uintptr_t>(deps) & 7)) return CD360_ERR_ALIGN;
  launch_ex(diffusion_loss_kernel, dim3(batch), dim3(512), 0, reinterpret_cast<cudaStream_t>(stream_), 1, eps,
            x_noisy, target, sigma, mask, coef, loss, mask_sum, CD360_BFW(deps), hw, ldd);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nerf_aux_loss(const float* fg, const float* alphas, const float* rgb,
                                   const float* op, const float* mask_s, const float* tgt,
                                   const float* mask_sum, const float* wfg, const float* wbg,
                                   const float* wrgb, float* loss3, float* dfg, float* dalphas,
                                   float* drgb, int32_t batch, int32_t hw, int32_t d,
                                   cd360_stream_t stream_) {
  if (!fg || !alphas || !op || !wfg || !wbg || !loss3 || !dfg || !dalphas) return CD360_ERR_NULL;
  if (rgb && (!mask_s || !tgt || !mask_sum || !wrgb || !drgb)) return CD360_ERR_NULL;
  if (batch <= 0 || hw <= 0 || d <= 0) return CD360_ERR_SHAPE;
  launch_ex(nerf_aux_loss_kernel, dim3(batch), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, fg,
            alphas, rgb, op, mask_s, tgt, mask_sum, wfg, wbg, wrgb, loss3, dfg, dalphas, drgb, hw, d);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_resize_bilinear_aa(const float* in, float* out, int32_t planes, int32_t ih,
                                        int32_t iw, int32_t oh, int32_t ow, float out_scale,
                                        float out_shift, cd360_stream_t stream_) {
  if (!in || !out) return CD360_ERR_NULL;
  if (planes <= 0 || ih <= 0 || iw <= 0 || oh <= 0 || ow <= 0) return CD360_ERR_SHAPE;
  const long long items = static_cast<long long>(planes) * oh * ow;
  launch_ex(resize_bilinear_aa_kernel, dim3(grid_for(items * 32, 256)), dim3(256), 0,
            reinterpret_cast<cudaStream_t>(stream_), 1, in, out, planes, ih, iw, oh, ow, out_scale, out_shift);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int32_t step,
                                float grad_scale, cd360_stream_t stream_) {
  if (!p || !g || !m || !v) return CD360_ERR_NULL;
  if (n <= 0 || step <= 0) return CD360_ERR_SHAPE;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  launch_ex(adamw_kernel, dim3(grid_for(n, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, p,
            g, m, v, static_cast<long long>(n), lr, beta1, beta2, eps, weight_decay,
            static_cast<float>(bc1), static_cast<float>(sqrt(bc2)), grad_scale);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}
