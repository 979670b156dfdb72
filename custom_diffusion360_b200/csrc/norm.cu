// norm.cu — HBM-bound normalisation kernels (NHWC bf16 activations, fp32 statistics).
//
//  * GroupNorm(32) [+ SiLU]: GroupNorm32 (sgm/modules/diffusionmodules/util.py:309-311) followed
//    by nn.SiLU in ResBlock.in_layers/out_layers and UNetModel.out (openaimodel.py:280-283,
//    315-317, 968-971) and Normalize() in SpatialTransformer (sgm/modules/attention.py:118,748).
//    Pass 1 reduces per-(image, chunk) partial moments with 16-byte coalesced loads; pass 2
//    finalises mean/rstd (double), folds them with gamma/beta into per-channel scale/shift in
//    smem, and streams x -> y.  The input may be a virtual channel concat of two tensors.
//  * LayerNorm: nn.LayerNorm(dim) x3 per BasicTransformerBlock (attention.py:531-533);
//    one warp per token row, values held in registers, two-pass variance.
#include "cd360_common.cuh"

namespace cd360 {

constexpr int GN_MAX_CHUNKS = 128;
constexpr int GN_GROUPS = 32;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
         d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// partial[b][chunk][g][2] = (sum, sumsq) over the chunk's rows and the group's channels.
// Thread (rl, vec) walks rows rl, rl + rlanes, ... of the chunk with FOUR 16-byte loads in flight
// (a single dependent load per thread left the kernel latency-bound at ~1.5 TB/s).
__global__ void __launch_bounds__(512, 2)
groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ x0, int c0,
                       const __nv_bfloat16* __restrict__ x1, int c1, float* __restrict__ partial,
                       int hw, int rows_per_chunk, int nvec, int rlanes) {
  CD360_TL(2);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  // per-(row lane, channel) partial sums, reduced in a FIXED order below: results are bit-exact
  // run to run (no floating-point atomics anywhere on the path)
  extern __shared__ float s_part[];  // [2][rlanes][ctot]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int ctot = c0 + c1;
  const int cg = ctot / GN_GROUPS;
  const int vec = threadIdx.x % nvec;
  const int rl = threadIdx.x / nvec;
  if (rl < rlanes) {
    const int ch0 = vec * 8;
    const __nv_bfloat16* src;
    long long ld;
    int ch;
    if (ch0 < c0) { src = x0; ld = c0; ch = ch0; } else { src = x1; ld = c1; ch = ch0 - c0; }
    const int r_begin = chunk * rows_per_chunk;
    const int r_end = min(hw, r_begin + rows_per_chunk);
    src += static_cast<long long>(b) * hw * ld + ch;
    float s[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; ss[i] = 0.f; }
    for (int r = r_begin + rl; r < r_end; r += 4 * rlanes) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rj = r + j * rlanes;
        u[j] = rj < r_end ? __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(rj) * ld))
                          : make_uint4(0u, 0u, 0u, 0u);  // zeros add nothing to either moment
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack8(u[j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; ss[i] = fmaf(f[i], f[i], ss[i]); }
      }
    }
    float4* ps = reinterpret_cast<float4*>(s_part + static_cast<size_t>(rl) * ctot + ch0);
    float4* pss = reinterpret_cast<float4*>(s_part + static_cast<size_t>(rlanes + rl) * ctot + ch0);
    ps[0] = make_float4(s[0], s[1], s[2], s[3]);
    ps[1] = make_float4(s[4], s[5], s[6], s[7]);
    pss[0] = make_float4(ss[0], ss[1], ss[2], ss[3]);
    pss[1] = make_float4(ss[4], ss[5], ss[6], ss[7]);
  }
  __syncthreads();
  // 64 outputs (group, which), four threads each; fixed summation order
  const int sub = threadIdx.x & 3;
  for (int o = threadIdx.x >> 2; o < GN_GROUPS * 2; o += blockDim.x >> 2) {
    const int g = o >> 1, which = o & 1;
    const float* base = s_part + static_cast<size_t>(which) * rlanes * ctot + g * cg;
    float acc = 0.f;
    const int n = rlanes * cg;
    for (int i = sub; i < n; i += 4) {
      const int rr = i / cg, c = i - rr * cg;
      acc += base[static_cast<size_t>(rr) * ctot + c];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0)
      partial[(static_cast<long long>(b) * gridDim.x + chunk) * GN_GROUPS * 2 + o] = acc;
  }
}

__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x0, int c0,
                       const __nv_bfloat16* __restrict__ x1, int c1,
                       const float* __restrict__ partial, int nchunks,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       __nv_bfloat16* __restrict__ out, int hw, int rows_per_block, float eps,
                       int apply_silu) {
  CD360_TL(3);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_dyn[];  // scale[ctot] | shift[ctot]
  __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
  const int b = blockIdx.y;
  const int ctot = c0 + c1;
  const int cg = ctot / GN_GROUPS;
  float* s_scale = s_dyn;
  float* s_shift = s_dyn + ctot;
  {  // finalise the moments: 8 threads per group, fixed order (blockDim.x == 256)
    const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
    double su = 0.0, sq = 0.0;
    const float2* pp = reinterpret_cast<const float2*>(partial) +
                       static_cast<long long>(b) * nchunks * GN_GROUPS;
    for (int c = sub; c < nchunks; c += 8) {
      const float2 t = __ldg(pp + c * GN_GROUPS + g);
      su += t.x;
      sq += t.y;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      su += __shfl_xor_sync(0xffffffffu, su, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (sub == 0) {
      const double n = static_cast<double>(hw) * cg;
      const double mean = su / n;
      double var = sq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = static_cast<float>(mean);
      s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < ctot; c += blockDim.x) {
    const int g = c / cg;
    const float sc = s_rstd[g] * gamma[c];
    s_scale[c] = sc;
    s_shift[c] = beta[c] - s_mean[g] * sc;
  }
  __syncthreads();
  const int nvec = ctot / 8;
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(hw, r_begin + rows_per_block);
  // item i of this block = (row r_begin + i / nvec, vector i % nvec); threads stride the items by
  // blockDim.x and track (row, vec) incrementally (no division in the loop), 4 loads in flight
  const int dr = static_cast<int>(blockDim.x) / nvec, dv = static_cast<int>(blockDim.x) % nvec;
  int r = r_begin + static_cast<int>(threadIdx.x) / nvec;
  int v = static_cast<int>(threadIdx.x) % nvec;
  const long long img = static_cast<long long>(b) * hw;
  while (r < r_end) {
    uint4 u[4];
    int rr[4], vv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      rr[j] = r;
      vv[j] = v;
      if (r < r_end) {
        const int ch0 = v * 8;
        const long long grow = img + r;
        const __nv_bfloat16* src =
            (ch0 < c0) ? (x0 + grow * c0 + ch0) : (x1 + grow * c1 + (ch0 - c0));
        u[j] = __ldg(reinterpret_cast<const uint4*>(src));
      }
      r += dr;
      v += dv;
      if (v >= nvec) { v -= nvec; ++r; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (rr[j] < r_end) {
        const int ch0 = vv[j] * 8;
        float f[8];
        unpack8(u[j], f);
        const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + ch0);
        const float4 sc1 = *reinterpret_cast<const float4*>(s_scale + ch0 + 4);
        const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + ch0);
        const float4 sh1 = *reinterpret_cast<const float4*>(s_shift + ch0 + 4);
        const float scs[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
        const float shs[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float y = fmaf(f[k], scs[k], shs[k]);
          f[k] = apply_silu ? silu_f(y) : y;
        }
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]);
        o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]);
        o.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(out + (img + rr[j]) * ctot + ch0) = o;
      }
    }
  }
}

constexpr int LN_MAXV = 8;  // c <= 2048

__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, int rows, int c,
                 float eps) {
  CD360_TL(4);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = c >> 3;
  const __nv_bfloat16* src = x + static_cast<long long>(warp) * c;
  float v[LN_MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + vi * 8));
      unpack8(u, v[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += v[i][k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / c;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = v[i][k] - mean;
        sq = fmaf(d, d, sq);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / c + eps);
  __nv_bfloat16* dst = out + static_cast<long long>(warp) * c;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float y[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = fmaf((v[i][k] - mean) * rstd, gg[k], bb[k]);
      uint4 o;
      o.x = pack_bf16x2(y[0], y[1]);
      o.y = pack_bf16x2(y[2], y[3]);
      o.z = pack_bf16x2(y[4], y[5]);
      o.w = pack_bf16x2(y[6], y[7]);
      *reinterpret_cast<uint4*>(dst + vi * 8) = o;
    }
  }
}

static int gn_chunks(int batch, int hw) {
  int chunks = (2 * kNumSMsB200) / batch;  // two resident CTAs per SM, one wave
  if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
  const int max_by_rows = (hw + 7) / 8;
  if (chunks > max_by_rows) chunks = max_by_rows;
  if (chunks < 1) chunks = 1;
  return chunks;
}

}  // namespace cd360

using namespace cd360;

extern "C" int64_t cd360_groupnorm_workspace_floats(int32_t batch, int32_t hw) {
  (void)hw;
  return static_cast<int64_t>(batch) * GN_MAX_CHUNKS * GN_GROUPS * 2;
}

extern "C" int cd360_groupnorm_silu_bf16(const void* x0, int32_t c0, const void* x1, int32_t c1,
                                         const float* gamma, const float* beta, void* out,
                                         float* workspace, int32_t batch, int32_t hw, float eps,
                                         int32_t apply_silu, cd360_stream_t stream_) {
  if (!x0 || !gamma || !beta || !out || !workspace) return CD360_ERR_NULL;
  if (c1 > 0 && !x1) return CD360_ERR_NULL;
  if (c1 < 0 || c0 <= 0 || batch <= 0 || hw <= 0 || batch > 65535) return CD360_ERR_SHAPE;
  const int ctot = c0 + c1;
  if ((c0 & 7) || (c1 & 7) || (ctot % GN_GROUPS) != 0) return CD360_ERR_SHAPE;
  const int nvec = ctot / 8;
  if ((reinterpret_cast<uintptr_t>(x0) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (x1 && (reinterpret_cast<uintptr_t>(x1) & 15)))
    return CD360_ERR_ALIGN;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int chunks = gn_chunks(batch, hw);
  const int rows_per_chunk = (hw + chunks - 1) / chunks;
  if (nvec > 512) return CD360_ERR_SHAPE;
  int rlanes = 512 / nvec;
  if (rlanes > rows_per_chunk) rlanes = rows_per_chunk;
  if (rlanes < 1) rlanes = 1;
  int threads = ((nvec * rlanes + 31) / 32) * 32;
  if (threads < 256) threads = 256;
  const size_t stats_smem = static_cast<size_t>(2) * rlanes * ctot * sizeof(float);  // <= 40 KiB
  launch_ex(groupnorm_stats_kernel, dim3(chunks, batch), dim3(threads), stats_smem, stream, 1,
      reinterpret_cast<const __nv_bfloat16*>(x0), c0, reinterpret_cast<const __nv_bfloat16*>(x1),
      c1, workspace, hw, rows_per_chunk, nvec, rlanes);
  CD360_CHECK_LAUNCH();
  int row_blocks = (4 * kNumSMsB200) / batch;  // four resident CTAs per SM, one wave
  if (row_blocks < 1) row_blocks = 1;
  if (row_blocks > hw) row_blocks = hw;
  const int rows_per_block = (hw + row_blocks - 1) / row_blocks;
  row_blocks = (hw + rows_per_block - 1) / rows_per_block;
  const size_t smem = static_cast<size_t>(ctot) * 2 * sizeof(float);
  launch_ex(groupnorm_apply_kernel, dim3(row_blocks, batch), dim3(256), smem, stream, 1, 
      reinterpret_cast<const __nv_bfloat16*>(x0), c0, reinterpret_cast<const __nv_bfloat16*>(x1),
      c1, workspace, chunks, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), hw,
      rows_per_block, eps, apply_silu);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_layernorm_bf16(const void* x, const float* gamma, const float* beta, void* out,
                                    int32_t rows, int32_t c, float eps, cd360_stream_t stream_) {
  if (!x || !gamma || !beta || !out) return CD360_ERR_NULL;
  if (rows <= 0 || c <= 0 || (c & 7) || c > LN_MAXV * 256) return CD360_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(gamma) & 15) || (reinterpret_cast<uintptr_t>(beta) & 15))
    return CD360_ERR_ALIGN;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int warps_per_block = 8;
  const int blocks = (rows + warps_per_block - 1) / warps_per_block;
  launch_ex(layernorm_kernel, dim3(blocks), dim3(warps_per_block * 32), 0, stream, 1, 
      reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(out),
      rows, c, eps);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

CD360_TL_SETTER(norm)
