// elementwise.cu — small HBM/latency-bound kernels around the tensor-core path:
// timestep embedding, the small-M embedding linears, im2col / resampling layout helpers, dtype
// casts, and the fused denoiser-scaling + CFG + Euler update of the sampler.
// Reference lines are cited per kernel.
#include "cd360_common.cuh"

namespace cd360 {

// timestep_embedding (sgm/modules/diffusionmodules/util.py:206-230): [cos(t f_k) | sin(t f_k)]
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out,
                                          int batch, int dim) {
  CD360_TL(6);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * half) return;
  const int b = i / half, k = i - b * half;
  // freqs = exp(-ln(10000) * k / half), computed in fp32 like the reference
  const float freq = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
  const float arg = t[b] * freq;
  float s, c;
  sincosf(arg, &s, &c);
  out[static_cast<long long>(b) * dim + k] = c;
  out[static_cast<long long>(b) * dim + half + k] = s;
  if ((dim & 1) && k == 0) out[static_cast<long long>(b) * dim + dim - 1] = 0.f;
}

// out[b, n] = act_out(sum_k act_in(x[b,k]) W[n,k] + bias[n]) + add[b,n]
// (time_embed / label_emb, openaimodel.py:679-713,1026-1031; ResBlock.emb_layers, :307-313,363)
constexpr int SL_BCHUNK = 4;
constexpr int SL_OUT_PER_WARP = 2;
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                    const float* __restrict__ bias, const float* __restrict__ add,
                    float* __restrict__ out, int batch, int n, int k, int act_in, int act_out) {
  CD360_TL(5);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_x[];  // [SL_BCHUNK][k]
  const int b0 = blockIdx.y * SL_BCHUNK;
  const int nb = min(SL_BCHUNK, batch - b0);
  for (int i = threadIdx.x; i < SL_BCHUNK * k; i += blockDim.x) {
    const int bi = i / k, kk = i - bi * k;
    float v = 0.f;
    if (bi < nb) {
      v = x[static_cast<long long>(b0 + bi) * k + kk];
      if (act_in == CD360_ACT_SILU) v = v / (1.0f + expf(-v));
    }
    s_x[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = k >> 3;
  // a warp owns SL_OUT_PER_WARP consecutive output columns and walks their weight rows TOGETHER:
  // SL_OUT_PER_WARP x 2 independent 16-byte loads in flight per lane (the M = batch <= 4 product is
  // pure weight streaming; one dependent load per lane left it latency-bound at ~0.1 TB/s)
  const int col0 = (blockIdx.x * 8 + warp) * SL_OUT_PER_WARP;
  if (col0 < n) {
    float acc[SL_OUT_PER_WARP][SL_BCHUNK];
#pragma unroll
    for (int oi = 0; oi < SL_OUT_PER_WARP; ++oi)
#pragma unroll
      for (int bi = 0; bi < SL_BCHUNK; ++bi) acc[oi][bi] = 0.f;
    for (int v0 = lane; v0 < nvec; v0 += 64) {
      uint4 u[SL_OUT_PER_WARP][2];
#pragma unroll
      for (int oi = 0; oi < SL_OUT_PER_WARP; ++oi) {
        const int col = min(col0 + oi, n - 1);  // clamped: duplicates are discarded below
        const __nv_bfloat16* wr = w + static_cast<long long>(col) * k;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int v = v0 + 32 * h;
          u[oi][h] = v < nvec ? __ldg(reinterpret_cast<const uint4*>(wr + v * 8))
                              : make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int v = min(v0 + 32 * h, nvec - 1);  // zero weights make the clamped x harmless
#pragma unroll
        for (int oi = 0; oi < SL_OUT_PER_WARP; ++oi) {
          const uint4 uu = u[oi][h];
          const float2 w0 = unpack_bf16x2(uu.x), w1 = unpack_bf16x2(uu.y), w2 = unpack_bf16x2(uu.z),
                       w3 = unpack_bf16x2(uu.w);
          const float wf[8] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y, w3.x, w3.y};
#pragma unroll
          for (int bi = 0; bi < SL_BCHUNK; ++bi) {
            const float* xr = s_x + bi * k + v * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[oi][bi] = fmaf(wf[j], xr[j], acc[oi][bi]);
          }
        }
      }
    }
#pragma unroll
    for (int oi = 0; oi < SL_OUT_PER_WARP; ++oi) {
#pragma unroll
      for (int bi = 0; bi < SL_BCHUNK; ++bi) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          acc[oi][bi] += __shfl_xor_sync(0xffffffffu, acc[oi][bi], o);
      }
      const int col = col0 + oi;
      if (lane == 0 && col < n) {
        for (int bi = 0; bi < nb; ++bi) {
          float r = acc[oi][bi] + (bias ? bias[col] : 0.f);
          if (act_out == CD360_ACT_SILU) r = r / (1.0f + expf(-r));
          const long long o = static_cast<long long>(b0 + bi) * n + col;
          if (add) r += add[o];
          out[o] = r;
        }
      }
    }
  }
}

// input conv im2col: x fp32 NCHW * scale[b] -> bf16 [B*H*W, kpad], k = (ky*3+kx)*Cin + c
// (UNetModel.input_blocks[0], openaimodel.py:719-721; c_in scaling denoiser.py:42-43)
__global__ void im2col3x3_nchw_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                      __nv_bfloat16* __restrict__ out, int batch, int src_batch,
                                      int cin, int h, int w, int kpad) {
  CD360_TL(7);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(batch) * h * w * kpad;
  if (i >= total) return;
  const int kk = static_cast<int>(i % kpad);
  const long long pix = i / kpad;
  const int px = static_cast<int>(pix % w);
  const int py = static_cast<int>((pix / w) % h);
  const int b = static_cast<int>(pix / (static_cast<long long>(w) * h));
  float v = 0.f;
  if (kk < 9 * cin) {
    const int tap = kk / cin, c = kk - tap * cin;
    const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      // batch row b reads image b % src_batch: the CFG replication `torch.cat([x] * rows)`
      // (guiders.py:133) is folded into the load
      v = x[((static_cast<long long>(b % src_batch) * cin + c) * h + yy) * w + xx];
      if (scale) v *= scale[b];
    }
  }
  out[i] = __float2bfloat16_rn(v);
}

// stride-2 pad-1 3x3 im2col of an NHWC bf16 tensor, 8 channels per thread
// (Downsample.op, openaimodel.py:215-222)
__global__ void im2col3x3_s2_kernel(const __nv_bfloat16* __restrict__ x,
                                    __nv_bfloat16* __restrict__ out, int batch, int h, int w,
                                    int c) {
  CD360_TL(8);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const int ho = h / 2, wo = w / 2, cv = c / 8;
  const long long total = static_cast<long long>(batch) * ho * wo * 9 * cv;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % cv);
  long long r = i / cv;
  const int tap = static_cast<int>(r % 9);
  r /= 9;
  const int ox = static_cast<int>(r % wo);
  const int oy = static_cast<int>((r / wo) % ho);
  const int b = static_cast<int>(r / (static_cast<long long>(wo) * ho));
  const int yy = oy * 2 + tap / 3 - 1, xx = ox * 2 + tap % 3 - 1;
  uint4 u = make_uint4(0, 0, 0, 0);
  if (yy >= 0 && yy < h && xx >= 0 && xx < w)
    u = __ldg(reinterpret_cast<const uint4*>(
        x + ((static_cast<long long>(b) * h + yy) * w + xx) * c + v * 8));
  *reinterpret_cast<uint4*>(out + (r * 9 + tap) * c + v * 8) = u;
}

// nearest x2 (Upsample.forward, openaimodel.py:161), NHWC, 8 channels per thread
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x,
                                  __nv_bfloat16* __restrict__ out, int batch, int h, int w, int c) {
  CD360_TL(9);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const int cv = c / 8;
  const long long total = static_cast<long long>(batch) * (2 * h) * (2 * w) * cv;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % cv);
  long long r = i / cv;
  const int ox = static_cast<int>(r % (2 * w));
  const int oy = static_cast<int>((r / (2 * w)) % (2 * h));
  const int b = static_cast<int>(r / (static_cast<long long>(4) * w * h));
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(
      x + ((static_cast<long long>(b) * h + oy / 2) * w + ox / 2) * c + v * 8));
  *reinterpret_cast<uint4*>(out + r * c + v * 8) = u;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                     long long n) {
  CD360_TL(10);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(x[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                     long long n) {
  CD360_TL(11);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(x[i]);
}

// [B, hw, C] (bf16 or fp32) -> NCHW fp32 [B, C, hw]
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x, int x_is_fp32,
                                    float* __restrict__ out, int batch, int hw, int c) {
  CD360_TL(12);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(batch) * hw * c;
  if (i >= total) return;
  const int p = static_cast<int>(i % hw);
  const int ch = static_cast<int>((i / hw) % c);
  const int b = static_cast<int>(i / (static_cast<long long>(hw) * c));
  const long long src = (static_cast<long long>(b) * hw + p) * c + ch;
  out[i] = x_is_fp32 ? reinterpret_cast<const float*>(x)[src]
                     : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[src]);
}

// NCHW fp32 [B, C, hw] -> [B, hw, C] bf16 (module inputs at the sgm boundary)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                    int batch, int hw, int c) {
  CD360_TL(13);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(batch) * hw * c;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  const int p = static_cast<int>((i / c) % hw);
  const int b = static_cast<int>(i / (static_cast<long long>(hw) * c));
  out[i] = __float2bfloat16_rn(x[(static_cast<long long>(b) * c + ch) * hw + p]);
}

// Fused EpsScaling denoiser output + CFG combine + Euler step.  See include/cd360.h.
__global__ void cfg_euler_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                 float* __restrict__ denoised_out, int n_img, int g, int hw,
                                 float sigma_q, float sigma, float sigma_next, float scale,
                                 float scale_im, const float* __restrict__ sig_dev) {
  CD360_TL(14);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  if (sig_dev != nullptr) {  // (sigma_q, sigma, sigma_next) live in device memory: graph-replayable
    sigma_q = sig_dev[0];
    sigma = sig_dev[1];
    sigma_next = sig_dev[2];
  }
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_img) * 4 * hw;
  if (i >= total) return;
  const int p = static_cast<int>(i % hw);
  const int ch = static_cast<int>((i / hw) % 4);
  const int img = static_cast<int>(i / (static_cast<long long>(hw) * 4));
  const float xv = x[i];
  float d[3];
  for (int r = 0; r < g; ++r) {
    const float e = eps[((static_cast<long long>(r) * n_img + img) * hw + p) * 4 + ch];
    d[r] = e * (-sigma_q) + xv;  // predict * c_out + input * c_skip
  }
  float den;
  if (g == 3) {
    den = d[0] + scale * (d[2] - d[1]) + scale_im * (d[1] - d[0]);
  } else if (g == 2) {
    den = d[0] + scale * (d[1] - d[0]);
  } else {
    den = d[0];
  }
  if (denoised_out) denoised_out[i] = den;
  const float dd = (xv - den) / sigma;  // to_d
  x[i] = xv + (sigma_next - sigma) * dd;
}

static inline int blocks_for(long long n, int t) { return static_cast<int>((n + t - 1) / t); }

}  // namespace cd360

using namespace cd360;

extern "C" int cd360_timestep_embedding(const float* t, float* out, int32_t batch, int32_t dim,
                                        cd360_stream_t stream_) {
  if (!t || !out) return CD360_ERR_NULL;
  if (batch <= 0 || dim < 2) return CD360_ERR_SHAPE;
  const int n = batch * (dim / 2);
  launch_ex(timestep_embedding_kernel, dim3(blocks_for(n, 128)), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      t, out, batch, dim);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_small_linear(const float* x, const void* w, const float* bias,
                                  const float* add, float* out, int32_t batch, int32_t n,
                                  int32_t k, int32_t act_in, int32_t act_out,
                                  cd360_stream_t stream_) {
  if (!x || !w || !out) return CD360_ERR_NULL;
  if (batch <= 0 || n <= 0 || k <= 0 || (k & 7)) return CD360_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(w) & 15) return CD360_ERR_ALIGN;
  const size_t smem = static_cast<size_t>(SL_BCHUNK) * k * sizeof(float);
  if (smem > 48 * 1024) return CD360_ERR_SHAPE;
  dim3 grid((n + 8 * SL_OUT_PER_WARP - 1) / (8 * SL_OUT_PER_WARP),
            (batch + SL_BCHUNK - 1) / SL_BCHUNK);
  launch_ex(small_linear_kernel, dim3(grid), dim3(256), smem, reinterpret_cast<cudaStream_t>(stream_), 1, 
      x, reinterpret_cast<const __nv_bfloat16*>(w), bias, add, out, batch, n, k, act_in, act_out);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_im2col3x3_nchw_f32(const float* x, const float* scale, void* out,
                                        int32_t batch, int32_t src_batch, int32_t cin, int32_t h,
                                        int32_t w, int32_t kpad, cd360_stream_t stream_) {
  if (!x || !out) return CD360_ERR_NULL;
  if (src_batch <= 0) src_batch = batch;
  if (batch <= 0 || cin <= 0 || h <= 0 || w <= 0 || kpad < 9 * cin || (kpad & 7))
    return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(batch) * h * w * kpad;
  launch_ex(im2col3x3_nchw_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      x, scale, reinterpret_cast<__nv_bfloat16*>(out), batch, src_batch, cin, h, w, kpad);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_im2col3x3_s2_bf16(const void* x, void* out, int32_t batch, int32_t h,
                                       int32_t w, int32_t c, cd360_stream_t stream_) {
  if (!x || !out) return CD360_ERR_NULL;
  if (batch <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || c <= 0 || (c & 7))
    return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(batch) * (h / 2) * (w / 2) * 9 * (c / 8);
  launch_ex(im2col3x3_s2_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), batch, h, w,
      c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_upsample_nearest2x_bf16(const void* x, void* out, int32_t batch, int32_t h,
                                             int32_t w, int32_t c, cd360_stream_t stream_) {
  if (!x || !out) return CD360_ERR_NULL;
  if (batch <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 7)) return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(batch) * 4 * h * w * (c / 8);
  launch_ex(upsample2x_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), batch, h, w,
      c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_cast_f32_to_bf16(const float* x, void* out, int64_t n,
                                      cd360_stream_t stream_) {
  if (!x || !out) return CD360_ERR_NULL;
  if (n <= 0) return CD360_ERR_SHAPE;
  launch_ex(cast_f32_bf16_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      x, reinterpret_cast<__nv_bfloat16*>(out), n);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_cast_bf16_to_f32(const void* x, float* out, int64_t n,
                                      cd360_stream_t stream_) {
  if (!x || !out) return CD360_ERR_NULL;
  if (n <= 0) return CD360_ERR_SHAPE;
  launch_ex(cast_bf16_f32_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      reinterpret_cast<const __nv_bfloat16*>(x), out, n);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

namespace cd360 {
// Finish of a split-K GEMM (cd360_gemm_bf16 with k_splits > 1): one thread per 4 columns, slices
// summed in ascending order (deterministic).
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float* __restrict__ ws, long long ldw, long long split_stride, int slices,
                     const float* __restrict__ bias, const __nv_bfloat16* __restrict__ residual,
                     long long ldr, void* __restrict__ out, long long ldo, int out_fp32, long long M,
                     int N) {
  pdl_launch_dependents();
  pdl_wait();
  const int nv = N >> 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * nv) return;
  const long long m = i / nv;
  const int n = static_cast<int>(i - m * nv) * 4;
  const float* src = ws + m * ldw + n;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  int s = 0;
  for (; s + 4 <= slices; s += 4) {  // four loads in flight
    float4 t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = __ldcs(reinterpret_cast<const float4*>(src + (s + j) * split_stride));
#pragma unroll
    for (int j = 0; j < 4; ++j) { v.x += t[j].x; v.y += t[j].y; v.z += t[j].z; v.w += t[j].w; }
  }
  for (; s < slices; ++s) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(src + s * split_stride));
    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  }
  if (bias != nullptr) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (residual != nullptr) {
    const uint2 r = *reinterpret_cast<const uint2*>(residual + m * ldr + n);
    const float2 r0 = unpack_bf16x2(r.x), r1 = unpack_bf16x2(r.y);
    v.x += r0.x; v.y += r0.y; v.z += r1.x; v.w += r1.y;
  }
  if (out_fp32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + m * ldo + n) = v;
  } else {
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + m * ldo + n) = o;
  }
}
}  // namespace cd360

extern "C" int cd360_splitk_slices(int32_t k0, int32_t k1, int32_t k_splits) {
  const int nkb = (k0 + 63) / 64 + (k1 + 63) / 64;
  if (nkb <= 0) return CD360_ERR_SHAPE;
  if (k_splits <= 1 || nkb <= 1) return 1;
  const int want = k_splits < nkb ? k_splits : nkb;
  const int per = (nkb + want - 1) / want;
  return (nkb + per - 1) / per;
}

extern "C" int cd360_splitk_finish(const float* ws, int64_t ldw, int64_t split_stride, int32_t slices,
                                   const float* bias, const void* residual, int64_t ldr, void* out,
                                   int64_t ldo, int32_t out_fp32, int64_t M, int32_t N,
                                   cd360_stream_t stream_) {
  if (!ws || !out) return CD360_ERR_NULL;
  if (M <= 0 || N <= 0 || (N & 3) || ldw < N || ldo < N || slices <= 0) return CD360_ERR_SHAPE;
  if ((ldw & 3) || (ldo & 3) || (split_stride & 3) || (residual && (ldr & 3)) ||
      (reinterpret_cast<uintptr_t>(ws) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & (out_fp32 ? 15 : 7)) ||
      (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) ||
      (residual && (reinterpret_cast<uintptr_t>(residual) & 7)))
    return CD360_ERR_ALIGN;
  const long long total = static_cast<long long>(M) * (N >> 2);
  if (launch_ex(cd360::splitk_finish_kernel, dim3(blocks_for(total, 256)), dim3(256), 0,
                reinterpret_cast<cudaStream_t>(stream_), 1, ws, static_cast<long long>(ldw),
                static_cast<long long>(split_stride), slices, bias,
                reinterpret_cast<const __nv_bfloat16*>(residual), static_cast<long long>(ldr), out,
                static_cast<long long>(ldo), out_fp32, static_cast<long long>(M), N) != cudaSuccess)
    return CD360_ERR_LAUNCH;
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nhwc_to_nchw_f32(const void* x, int32_t x_is_fp32, float* out, int32_t batch,
                                      int32_t hw, int32_t c, cd360_stream_t stream_) {
  if (!x || !out) return CD360_ERR_NULL;
  if (batch <= 0 || hw <= 0 || c <= 0) return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(batch) * hw * c;
  launch_ex(nhwc_to_nchw_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      x, x_is_fp32, out, batch, hw, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nchw_f32_to_nhwc_bf16(const float* x, void* out, int32_t batch, int32_t hw,
                                           int32_t c, cd360_stream_t stream_) {
  if (!x || !out) return CD360_ERR_NULL;
  if (batch <= 0 || hw <= 0 || c <= 0) return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(batch) * hw * c;
  launch_ex(nchw_to_nhwc_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      x, reinterpret_cast<__nv_bfloat16*>(out), batch, hw, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_cfg_euler_step(float* x, const float* eps, float* denoised_out, int32_t n_img,
                                    int32_t guidance_rows, int32_t hw, float sigma_q, float sigma,
                                    float sigma_next, float scale, float scale_im,
                                    cd360_stream_t stream_) {
  if (!x || !eps) return CD360_ERR_NULL;
  if (n_img <= 0 || hw <= 0 || guidance_rows < 1 || guidance_rows > 3) return CD360_ERR_SHAPE;
  if (!(sigma > 0.f)) return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(n_img) * 4 * hw;
  launch_ex(cfg_euler_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      x, eps, denoised_out, n_img, guidance_rows, hw, sigma_q, sigma, sigma_next, scale, scale_im,
      nullptr);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_cfg_euler_step_dev(float* x, const float* eps, float* denoised_out,
                                        int32_t n_img, int32_t guidance_rows, int32_t hw,
                                        const float* sigmas3, float scale, float scale_im,
                                        cd360_stream_t stream_) {
  if (!x || !eps || !sigmas3) return CD360_ERR_NULL;
  if (n_img <= 0 || hw <= 0 || guidance_rows < 1 || guidance_rows > 3) return CD360_ERR_SHAPE;
  const long long total = static_cast<long long>(n_img) * 4 * hw;
  launch_ex(cfg_euler_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      x, eps, denoised_out, n_img, guidance_rows, hw, 0.f, 1.f, 0.f, scale, scale_im, sigmas3);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_abi_version(void) { return 7; }

extern "C" const char* cd360_strerror(int code) {
  switch (code) {
    case CD360_OK: return "ok";
    case CD360_ERR_SHAPE: return "unsupported or inconsistent shape";
    case CD360_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case CD360_ERR_UNSUPPORTED: return "dtype or mode not built";
    case CD360_ERR_LAUNCH: return "CUDA launch or driver error";
    case CD360_ERR_NULL: return "required pointer is NULL";
    default: return "unknown error";
  }
}

CD360_TL_SETTER(elementwise)
