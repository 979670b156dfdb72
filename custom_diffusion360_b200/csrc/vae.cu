// vae.cu — the two kernels the VAE decode needs beyond the UNet's:
//   * pointwise_conv_nchw: the 4 -> 4 channel 1x1 `post_quant_conv` with the 1/scale_factor of
//     decode_first_stage folded in (sgm/models/diffusion.py:207-212, autoencoder.py:313-316), fp32;
//   * softmax_rows: row softmax of the single-head mid-block attention (model.py:231-266) whose
//     head width (512) is outside the 64-wide flash kernel: scores and P.V are two tcgen05 GEMMs
//     (cd360_gemm_bf16), this kernel sits between them.  HBM-bound: reads fp32 scores twice (the
//     second pass hits L2: one row is <= 64 KB), writes bf16 weights once.
#include "cd360_common.cuh"

namespace cd360 {

constexpr int PW_MAXC = 8;

__global__ void __launch_bounds__(256)
pointwise_conv_nchw_kernel(const float* __restrict__ x, const float* __restrict__ w,
                           const float* __restrict__ bias, float* __restrict__ out, int cin,
                           int cout, long long hw, float scale, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_w[PW_MAXC * PW_MAXC], s_b[PW_MAXC];
  if (threadIdx.x < cout * cin) s_w[threadIdx.x] = w[threadIdx.x] * scale;
  if (threadIdx.x < cout) s_b[threadIdx.x] = bias != nullptr ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, pixel)
  if (i >= total) return;
  const long long b = i / hw, p = i - b * hw;
  float v[PW_MAXC];
#pragma unroll
  for (int c = 0; c < PW_MAXC; ++c) v[c] = c < cin ? x[(b * cin + c) * hw + p] : 0.f;
  for (int o = 0; o < cout; ++o) {
    float acc = s_b[o];
#pragma unroll
    for (int c = 0; c < PW_MAXC; ++c)
      if (c < cin) acc = fmaf(s_w[o * cin + c], v[c], acc);
    out[(b * cout + o) * hw + p] = acc;
  }
}

// one CTA per row; online (max, sum) in pass 1, normalised bf16 weights in pass 2
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, long long lds, __nv_bfloat16* __restrict__ out,
                    long long ldo, int n, float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  const float* row = s + static_cast<long long>(blockIdx.x) * lds;
  __nv_bfloat16* orow = out + static_cast<long long>(blockIdx.x) * ldo;
  const int nvec = n >> 2;
  float m = -INFINITY, l = 0.f;
  for (int i = threadIdx.x; i < nvec; i += 4 * blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ij = i + j * blockDim.x;
      v[j] = ij < nvec ? __ldg(reinterpret_cast<const float4*>(row) + ij)
                       : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    float mx = m;
#pragma unroll
    for (int j = 0; j < 4; ++j) mx = fmaxf(mx, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    if (mx == -INFINITY) continue;
    float acc = l * exp2f((m - mx) * scale_log2);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc += exp2f((v[j].x - mx) * scale_log2) + exp2f((v[j].y - mx) * scale_log2) +
             exp2f((v[j].z - mx) * scale_log2) + exp2f((v[j].w - mx) * scale_log2);
    m = mx;
    l = acc;
  }
  // block reduction of (m, l)
  __shared__ float s_m[8], s_l[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float mx = fmaxf(m, m2);
    l = (m == -INFINITY ? 0.f : l * exp2f((m - mx) * scale_log2)) +
        (m2 == -INFINITY ? 0.f : l2 * exp2f((m2 - mx) * scale_log2));
    m = mx;
  }
  if ((threadIdx.x & 31) == 0) { s_m[threadIdx.x >> 5] = m; s_l[threadIdx.x >> 5] = l; }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int wdx = 0; wdx < 8; ++wdx) M = fmaxf(M, s_m[wdx]);
  float L = 0.f;
#pragma unroll
  for (int wdx = 0; wdx < 8; ++wdx)
    if (s_m[wdx] != -INFINITY) L += s_l[wdx] * exp2f((s_m[wdx] - M) * scale_log2);
  const float inv = 1.f / L;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
    const float a = exp2f((v.x - M) * scale_log2) * inv, b = exp2f((v.y - M) * scale_log2) * inv,
                c = exp2f((v.z - M) * scale_log2) * inv, d = exp2f((v.w - M) * scale_log2) * inv;
    uint2 pk;
    pk.x = pack_bf16x2(a, b);
    pk.y = pack_bf16x2(c, d);
    *reinterpret_cast<uint2*>(orow + 4 * static_cast<long long>(i)) = pk;
  }
}

}  // namespace cd360

using namespace cd360;

extern "C" int cd360_pointwise_conv_nchw_f32(const float* x, const float* w, const float* bias,
                                             float* out, int32_t batch, int32_t cin, int32_t cout,
                                             int64_t hw, float scale, cd360_stream_t stream_) {
  if (!x || !w || !out) return CD360_ERR_NULL;
  if (batch <= 0 || hw <= 0 || cin <= 0 || cout <= 0 || cin > PW_MAXC || cout > PW_MAXC)
    return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(batch) * hw;
  const long long blocks = (total + 255) / 256;
  if (blocks > 2147483647LL) return CD360_ERR_SHAPE;
  if (launch_ex(pointwise_conv_nchw_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
                reinterpret_cast<cudaStream_t>(stream_), 1, x, w, bias, out, cin, cout,
                static_cast<long long>(hw), scale, total) != cudaSuccess)
    return CD360_ERR_LAUNCH;
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_softmax_rows_f32_bf16(const float* s, int64_t lds, void* out, int64_t ldo,
                                           int64_t rows, int32_t n, float scale,
                                           cd360_stream_t stream_) {
  if (!s || !out) return CD360_ERR_NULL;
  if (rows <= 0 || rows > 2147483647LL || n <= 0 || (n & 3) || lds < n || ldo < n)
    return CD360_ERR_SHAPE;
  if ((lds & 3) || (ldo & 3) || (reinterpret_cast<uintptr_t>(s) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 7))
    return CD360_ERR_ALIGN;
  if (launch_ex(softmax_rows_kernel, dim3(static_cast<unsigned>(rows)), dim3(256), 0,
                reinterpret_cast<cudaStream_t>(stream_), 1, s, static_cast<long long>(lds),
                reinterpret_cast<__nv_bfloat16*>(out), static_cast<long long>(ldo), n,
                scale * 1.4426950408889634f) != cudaSuccess)
    return CD360_ERR_LAUNCH;
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}
