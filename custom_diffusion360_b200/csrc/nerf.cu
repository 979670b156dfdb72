// nerf.cu — FeatureNeRF geometry / gather / view-softmax / volume-rendering kernels.
//
// The reference (sgm/modules/nerfsd_pytorch3d.py FeatureNeRFEncoding.forward :53-161) materialises,
// per (batch row, reference view, ray, depth sample), a (c+198)-wide input, runs a 2-layer MLP on
// it, softmaxes over the views and sums.  We restructure algebraically (exact in real arithmetic):
//   * the first Linear is split W1 = [W1f | W1p]; because grid_sample with zero padding is linear,
//     W1f . bilinear(F) == bilinear(W1f . F): W1f is applied ONCE to the n*hw reference tokens
//     (tensor-core GEMM, "G"), then gathered;
//   * the same holds for the feature columns of the `nviews` logit (one extra column of G);
//   * the second Linear commutes with the view-weighted sum (softmax weights sum to one) and is
//     applied after it.
// What remains here is geometry in fp32 registers (projection, positional encodings, Plücker
// coordinates), the 4-tap gather of G, SiLU, the softmax over views and the volume-rendering scan.
// Camera conventions follow PyTorch3D (row vectors, X_cam = X_world R + T, NDC +X left / +Y up;
// SURVEY.md §8c restates the pinned-dependency semantics).
#include <algorithm>

#include "cd360_common.cuh"

namespace cd360 {

struct Cam {
  float R[9];
  float T[3];
  float f[2];
  float pp[2];
};

__device__ __forceinline__ Cam load_cam(const float* __restrict__ p) {
  Cam c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.R[i] = p[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) c.T[i] = p[9 + i];
  c.f[0] = p[12]; c.f[1] = p[13]; c.pp[0] = p[14]; c.pp[1] = p[15];
  return c;
}
// row-vector transform: out_j = sum_k v_k R[k][j] (+ T_j)
__device__ __forceinline__ void xform_point(const Cam& c, const float (&v)[3], float (&o)[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
    o[j] = v[0] * c.R[0 * 3 + j] + v[1] * c.R[1 * 3 + j] + v[2] * c.R[2 * 3 + j] + c.T[j];
}
__device__ __forceinline__ void xform_dir(const Cam& c, const float (&v)[3], float (&o)[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
    o[j] = v[0] * c.R[0 * 3 + j] + v[1] * c.R[1 * 3 + j] + v[2] * c.R[2 * 3 + j];
}
// camera centre C = -T R^T ; C_k = -sum_j T_j R[k][j]   (get_camera_center)
__device__ __forceinline__ void cam_center(const Cam& c, float (&o)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k)
    o[k] = -(c.T[0] * c.R[k * 3 + 0] + c.T[1] * c.R[k * 3 + 1] + c.T[2] * c.R[k * 3 + 2]);
}
// normalised world-space direction of the ray through NDC (x, y) at depth 1
// (get_directional_raybundle, utils_cameraray.py:73-88: unproject - centre, normalised)
__device__ __forceinline__ void ray_dir(const Cam& c, float x, float y, float (&d)[3]) {
  const float vc[3] = {(x - c.pp[0]) / c.f[0], (y - c.pp[1]) / c.f[1], 1.0f};
  // world = (v_cam - T) R^T ; origin = -T R^T  => world - origin = v_cam R^T
  float w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = vc[0] * c.R[k * 3 + 0] + vc[1] * c.R[k * 3 + 1] + vc[2] * c.R[k * 3 + 2];
  const float inv = 1.0f / sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = w[k] * inv;
}

constexpr float kPi = 3.14159274101257324f;  // fl32(pi), as in `2.0**arange * np.pi` on a float32 tensor

// positional_encoding (utils_cameraray.py:222-242): bands 2^(-nf/2 .. nf/2-1) * pi,
// output order [sin band0 (dim comps), sin band1, ..., cos band0, ...]
template <int DIM, int NF, typename F>
__device__ __forceinline__ void pos_enc(const float (&v)[DIM], F&& emit) {
#pragma unroll 1
  for (int fb = 0; fb < NF; ++fb) {
    const float freq = exp2f(static_cast<float>(fb - NF / 2)) * kPi;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      float s, c;
      sincosf(v[k] * freq, &s, &c);
      emit(fb * DIM + k, s);
      emit(NF * DIM + fb * DIM + k, c);
    }
  }
}

// One thread per (b, ray, sample); loops over the n reference views.  The 198 positional features of a
// point are staged in shared memory ([128 points][kpe] bf16, the block's rows are contiguous in `pe`)
// and written with coalesced 16-byte stores: a thread storing its own 400-byte row two bytes at a
// time touched 32 sectors per store instruction and left the kernel at a few hundred GB/s.
constexpr int NP_THREADS = 128;
__global__ void __launch_bounds__(NP_THREADS)
nerf_points_kernel(const float* __restrict__ cams, const float* __restrict__ xy,
                   const float* __restrict__ depths, const float* __restrict__ w_nv_geo,
                   const float* __restrict__ b_nv, __nv_bfloat16* __restrict__ pe, int* __restrict__ gidx,
                   float* __restrict__ gwgt, float* __restrict__ vlogit, int nb, int n, int res,
                   int d, int kpe) {
  CD360_TL(16);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float s_w[];  // w_nv_geo[198] | per-view origin logit [16] | pe tile
  float* s_ol = s_w + 200;
  __nv_bfloat16* s_pe = reinterpret_cast<__nv_bfloat16*>(s_w + 200 + 16);  // [NP_THREADS][kpe]
  const int hw = res * res;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 198; i += blockDim.x) s_w[i] = w_nv_geo[i];
  __syncthreads();
  const Cam tgt = load_cam(cams + (static_cast<long long>(b) * (n + 1)) * 16);
  // per-view part of the nviews logit: origin of the reference camera in the target frame
  // (convert_to_target_space(pose, rays[:, 1:])[..., :3], nerfsd_pytorch3d.py:116-123) and its PE16
  for (int v = threadIdx.x; v < n; v += blockDim.x) {
    const Cam rc = load_cam(cams + (static_cast<long long>(b) * (n + 1) + 1 + v) * 16);
    float oc[3], ot[3];
    cam_center(rc, oc);
    xform_point(tgt, oc, ot);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc += s_w[99 + k] * ot[k];
    pos_enc<3, 16>(ot, [&](int idx, float val) { acc += s_w[102 + idx] * val; });
    s_ol[v] = acc;
  }
  __syncthreads();

  const int pid0 = blockIdx.x * blockDim.x;                // first point of this block
  const bool valid = pid0 + static_cast<int>(threadIdx.x) < hw * d;
  const int pid = valid ? pid0 + threadIdx.x : hw * d - 1;  // ray * d + sample (tail threads idle along)
  const int rows_here = min(static_cast<int>(blockDim.x), hw * d - pid0);
  const int ray = pid / d;
  const float depth = depths[pid];
  const float x = xy[ray * 2], y = xy[ray * 2 + 1];
  float o_t[3], d_t[3], pw[3];
  cam_center(tgt, o_t);
  ray_dir(tgt, x, y, d_t);
#pragma unroll
  for (int k = 0; k < 3; ++k) pw[k] = o_t[k] + depth * d_t[k];  // ray_bundle_to_ray_points

  // shared-over-views part of the logit: PE16(p in target view frame) (96) | p_tgt (3)
  float p_t[3];
  xform_point(tgt, pw, p_t);
  float logit_pt = b_nv != nullptr ? __ldg(b_nv) : 0.f;
  pos_enc<3, 16>(p_t, [&](int idx, float val) { logit_pt += s_w[idx] * val; });
#pragma unroll
  for (int k = 0; k < 3; ++k) logit_pt += s_w[96 + k] * p_t[k];

  for (int v = 0; v < n; ++v) {
    const Cam rc = load_cam(cams + (static_cast<long long>(b) * (n + 1) + 1 + v) * 16);
    const long long point = ((static_cast<long long>(b) * n + v) * hw) * d + pid;
    // ---- projection into the reference view (transform_points_ndc) and grid_sample taps ----
    float pv[3];
    xform_point(rc, pw, pv);
    float gx = -(rc.f[0] * pv[0] / pv[2] + rc.pp[0]);
    float gy = -(rc.f[1] * pv[1] / pv[2] + rc.pp[1]);
    // torch.nan_to_num then clip(+-1.2): NaN -> 0, +-inf -> +-FLT_MAX -> +-1.2
    gx = (gx != gx) ? 0.f : fminf(fmaxf(gx, -1.2f), 1.2f);
    gy = (gy != gy) ? 0.f : fminf(fmaxf(gy, -1.2f), 1.2f);
    const float ix = (gx + 1.f) * 0.5f * static_cast<float>(res - 1);
    const float iy = (gy + 1.f) * 0.5f * static_cast<float>(res - 1);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
    const float tx = ix - fx0, ty = iy - fy0;
    const int xs[2] = {x0, x0 + 1}, ys[2] = {y0, y0 + 1};
    const float wx[2] = {1.f - tx, tx}, wy[2] = {1.f - ty, ty};
    int4 gi;
    float4 gw;
    int* gip = reinterpret_cast<int*>(&gi);
    float* gwp = reinterpret_cast<float*>(&gw);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const bool in = xs[c] >= 0 && xs[c] < res && ys[a] >= 0 && ys[a] < res;
        gip[a * 2 + c] = in ? ys[a] * res + xs[c] : -1;
        gwp[a * 2 + c] = in ? wy[a] * wx[c] : 0.f;
      }
    if (valid) {
      *reinterpret_cast<int4*>(gidx + point * 4) = gi;
      *reinterpret_cast<float4*>(gwgt + point * 4) = gw;
      vlogit[point] = logit_pt + s_ol[v];
    }

    // ---- the 198 positional features of plane_coefs' input ----
    __nv_bfloat16* row = s_pe + static_cast<int>(threadIdx.x) * kpe;
    // PE16(p_view) 96 | p_view 3
    pos_enc<3, 16>(pv, [&](int idx, float val) { row[idx] = __float2bfloat16_rn(val); });
#pragma unroll
    for (int k = 0; k < 3; ++k) row[96 + k] = __float2bfloat16_rn(pv[k]);
    // target ray in the reference view frame (convert_to_view_space), Plücker (dir, o x dir), PE8
    float ov[3], dv[3];
    xform_point(rc, o_t, ov);
    xform_dir(rc, d_t, dv);
    const float inv = 1.0f / sqrtf(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
    const float dn[3] = {dv[0] * inv, dv[1] * inv, dv[2] * inv};
    const float pl[6] = {dn[0], dn[1], dn[2], ov[1] * dn[2] - ov[2] * dn[1],
                         ov[2] * dn[0] - ov[0] * dn[2], ov[0] * dn[1] - ov[1] * dn[0]};
    pos_enc<6, 8>(pl, [&](int idx, float val) { row[99 + idx] = __float2bfloat16_rn(val); });
#pragma unroll
    for (int k = 0; k < 3; ++k) row[195 + k] = __float2bfloat16_rn(dv[k]);
    for (int k = 198; k < kpe; ++k) row[k] = __float2bfloat16_rn(0.f);
    __syncthreads();
    {  // coalesced copy-out of the block's rows_here x kpe tile (kpe % 8 == 0: whole 16-byte vectors)
      const long long first = ((static_cast<long long>(b) * n + v) * hw) * d + pid0;
      uint4* dst = reinterpret_cast<uint4*>(pe + first * kpe);
      const uint4* src = reinterpret_cast<const uint4*>(s_pe);
      const int nvec = rows_here * (kpe >> 3);
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();  // the tile is rewritten for the next view
  }
}

__device__ __forceinline__ void fma8(float (&acc)[8], const uint4& u, float w) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
         d = unpack_bf16x2(u.w);
  acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]);
  acc[2] = fmaf(w, b.x, acc[2]); acc[3] = fmaf(w, b.y, acc[3]);
  acc[4] = fmaf(w, c.x, acc[4]); acc[5] = fmaf(w, c.y, acc[5]);
  acc[6] = fmaf(w, d.x, acc[6]); acc[7] = fmaf(w, d.y, acc[7]);
}

constexpr int NERF_MAX_VIEWS = 16;

// One warp per (b, ray*d + sample).
__global__ void __launch_bounds__(256)
nerf_combine_kernel(const __nv_bfloat16* __restrict__ g, long long ldg,
                    const __nv_bfloat16* __restrict__ hpre, const int* __restrict__ gidx,
                    const float* __restrict__ gwgt, const float* __restrict__ vlogit,
                    __nv_bfloat16* __restrict__ s_out, float* __restrict__ view_softmax, int nb,
                    int n, int hw, int d, int c) {
  CD360_TL(17);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long npts = static_cast<long long>(nb) * hw * d;
  if (wid >= npts) return;
  const int b = static_cast<int>(wid / (static_cast<long long>(hw) * d));
  const long long p = wid - static_cast<long long>(b) * hw * d;

  // --- view softmax: lane v owns view v ---
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
  float e = (lane < n) ? expf(logit - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float a_mine = e / sum;
  if (lane < n) view_softmax[(static_cast<long long>(b) * n + lane) * hw * d + p] = a_mine;

  // --- weighted sum over views of SiLU(hpre + gather(G)) ---
  const int nvec = c >> 3;
  // every lane runs every iteration: the shuffles below name the whole warp, and their source
  // lanes (views) must be converged with it even when c/8 is not a multiple of 32 (c = 640)
  for (int v0 = 0; v0 < nvec; v0 += 32) {
    const int vi = v0 + lane;
    const bool act = vi < nvec;
    float out[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = 0.f;
    for (int v = 0; v < n; ++v) {
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
      {
        // hpre is read exactly once: stream it (evict-first) so that the G rows the 4-tap gathers keep
        // re-reading stay resident in L1 / L2 (the gathers are 4x the bytes of hpre, all cache hits)
        const uint4 u = __ldcs(reinterpret_cast<const uint4*>(hpre + point * c + vi * 8));
        float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z),
               a3 = unpack_bf16x2(u.w);
        h[0] = a0.x; h[1] = a0.y; h[2] = a1.x; h[3] = a1.y;
        h[4] = a2.x; h[5] = a2.y; h[6] = a3.x; h[7] = a3.y;
      }
      const __nv_bfloat16* gb = g + (static_cast<long long>(b) * n + v) * hw * ldg + vi * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (idx[k] >= 0)
          fma8(h, __ldg(reinterpret_cast<const uint4*>(gb + idx[k] * ldg)), wk[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) out[k] = fmaf(a_v, silu_f(h[k]), out[k]);
    }
    if (!act) continue;
    uint4 o;
    o.x = pack_bf16x2(out[0], out[1]);
    o.y = pack_bf16x2(out[2], out[3]);
    o.z = pack_bf16x2(out[4], out[5]);
    o.w = pack_bf16x2(out[6], out[7]);
    __stcs(reinterpret_cast<uint4*>(s_out + wid * c + vi * 8), o);
  }
}

// One warp per (b, ray): lanes 0..d-1 own the samples for the scan, then all lanes stride channels.
__global__ void __launch_bounds__(256)
nerf_volrender_kernel(const __nv_bfloat16* __restrict__ feats, const float* __restrict__ raw,
                      const float* __restrict__ dists, __nv_bfloat16* __restrict__ rendered,
                      float* __restrict__ fg, float* __restrict__ alphas, float* __restrict__ rgb,
                      int nb, int hw, int d, int c) {
  CD360_TL(18);  // tools/step_timeline.py; empty in the product build
  pdl_launch_dependents();
  pdl_wait();
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<long long>(nb) * hw) return;
  const int ray = static_cast<int>(wid % hw);
  float dd = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
  if (lane < d) {
    const float4 rw = *reinterpret_cast<const float4*>(raw + (wid * d + lane) * 4);
    const float sigma = expf(rw.w);  // trunc_exp forward == exp (attention.py:196-199)
    dd = dists[ray * d + lane] * sigma;
    r0 = 1.f / (1.f + expf(-rw.x));
    r1 = 1.f / (1.f + expf(-rw.y));
    r2 = 1.f / (1.f + expf(-rw.z));
  }
  // exclusive prefix sum of delta*density (get_weights, nerfsd_pytorch3d.py:179-189)
  float incl = dd;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const float excl = incl - dd;
  const float alpha = 1.f - expf(-dd);
  float w = alpha * expf(-excl);
  // torch.nan_to_num: NaN -> 0, +-inf -> +-FLT_MAX
  if (w != w) w = 0.f;
  else if (isinf(w)) w = w > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  if (lane >= d) w = 0.f;
  if (lane < d) alphas[wid * d + lane] = alpha;
  float sw = w, s0 = w * r0, s1 = w * r1, s2 = w * r2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    fg[wid] = sw;
    rgb[wid * 3 + 0] = s0;
    rgb[wid * 3 + 1] = s1;
    rgb[wid * 3 + 2] = s2;
  }
  const int nvec = c >> 3;
  for (int v0 = 0; v0 < nvec; v0 += 32) {  // whole warp in every iteration (see nerf_combine_kernel)
    const int vi = v0 + lane;
    const bool act = vi < nvec;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int s = 0; s < d; ++s) {
      const float ws = __shfl_sync(0xffffffffu, w, s);
      if (act) fma8(acc, __ldg(reinterpret_cast<const uint4*>(feats + (wid * d + s) * c + vi * 8)), ws);
    }
    if (!act) continue;
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]);
    o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]);
    o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(rendered + wid * c + vi * 8) = o;
  }
}

// Reference padding masks (FeatureNeRFEncoding.forward, nerfsd_pytorch3d.py:61-70):
//   mask = F.interpolate(mask_ref[(b n), 1, H, W], size=[res, res], mode="nearest");  xref = xref * mask
// One pass over the reference tokens: row (bn, y, x) is scaled by mask[bn, floor(y*H/res), floor(x*W/res)]
// (torch's legacy "nearest": src = min(floor(dst * scale), in - 1) with scale = in / out in fp32).
__global__ void __launch_bounds__(256)
nerf_mask_ref_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mask,
                     __nv_bfloat16* __restrict__ out, long long rows, int res, int mh, int mw, int c) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = c >> 3;
  const long long total = rows * nvec;
  const float sy = static_cast<float>(mh) / static_cast<float>(res);
  const float sx = static_cast<float>(mw) / static_cast<float>(res);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / nvec;
    const int hw = res * res;
    const long long bn = row / hw;
    const int pix = static_cast<int>(row - bn * hw);
    const int y = pix / res, xx = pix - y * res;
    const int my = min(static_cast<int>(floorf(static_cast<float>(y) * sy)), mh - 1);
    const int mx = min(static_cast<int>(floorf(static_cast<float>(xx) * sx)), mw - 1);
    const float m = __ldg(mask + (bn * mh + my) * mw + mx);
    const uint4 u = *reinterpret_cast<const uint4*>(x + i * 8);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    uint4 o;
    o.x = pack_bf16x2(a.x * m, a.y * m);
    o.y = pack_bf16x2(b.x * m, b.y * m);
    o.z = pack_bf16x2(cc.x * m, cc.y * m);
    o.w = pack_bf16x2(d.x * m, d.y * m);
    *reinterpret_cast<uint4*>(out + i * 8) = o;
  }
}

}  // namespace cd360

using namespace cd360;

extern "C" int cd360_nerf_mask_ref(const void* x, const float* mask, void* out, int64_t bn, int32_t res,
                                   int32_t mh, int32_t mw, int32_t c, cd360_stream_t stream_) {
  if (!x || !mask || !out) return CD360_ERR_NULL;
  if (bn <= 0 || res <= 0 || mh <= 0 || mw <= 0 || c <= 0 || (c & 7)) return CD360_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return CD360_ERR_ALIGN;
  const long long rows = static_cast<long long>(bn) * res * res;
  const long long total = rows * (c >> 3);
  const unsigned blocks = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148LL * 16));
  launch_ex(nerf_mask_ref_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1,
            reinterpret_cast<const __nv_bfloat16*>(x), mask, reinterpret_cast<__nv_bfloat16*>(out), rows, res,
            mh, mw, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nerf_points(const float* cams, const float* xy, const float* depths,
                                 const float* w_nv_geo, const float* b_nv, void* pe, int32_t* gidx,
                                 float* gwgt, float* vlogit, int32_t b, int32_t n, int32_t res,
                                 int32_t d, int32_t kpe, cd360_stream_t stream_) {
  if (!cams || !xy || !depths || !w_nv_geo || !pe || !gidx || !gwgt || !vlogit)
    return CD360_ERR_NULL;
  if (b <= 0 || n <= 0 || n > NERF_MAX_VIEWS || res <= 1 || d <= 0 || kpe < 198 || (kpe & 7))
    return CD360_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(gidx) & 15) || (reinterpret_cast<uintptr_t>(gwgt) & 15) ||
      (reinterpret_cast<uintptr_t>(pe) & 15))
    return CD360_ERR_ALIGN;
  if (kpe > 256) return CD360_ERR_SHAPE;
  const int pts = res * res * d;
  dim3 grid((pts + NP_THREADS - 1) / NP_THREADS, b);
  const size_t smem = (200 + 16) * sizeof(float) + static_cast<size_t>(NP_THREADS) * kpe * 2;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(nerf_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (200 + 16) * 4 + NP_THREADS * 256 * 2) != cudaSuccess)
      return CD360_ERR_LAUNCH;
    attr_set = true;
  }
  launch_ex(nerf_points_kernel, dim3(grid), dim3(NP_THREADS), smem, reinterpret_cast<cudaStream_t>(stream_), 1, 
      cams, xy, depths, w_nv_geo, b_nv, reinterpret_cast<__nv_bfloat16*>(pe), gidx, gwgt, vlogit, b,
      n, res, d, kpe);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nerf_combine(const void* g, int64_t ldg, const void* hpre,
                                  const int32_t* gidx, const float* gwgt, const float* vlogit,
                                  void* s, float* view_softmax, int32_t b, int32_t n, int32_t hw,
                                  int32_t d, int32_t c, cd360_stream_t stream_) {
  if (!g || !hpre || !gidx || !gwgt || !vlogit || !s || !view_softmax) return CD360_ERR_NULL;
  if (b <= 0 || n <= 0 || n > NERF_MAX_VIEWS || hw <= 0 || d <= 0 || c <= 0 || (c & 7) ||
      ldg < c + 1 || (ldg & 7))
    return CD360_ERR_SHAPE;
  const long long warps = static_cast<long long>(b) * hw * d;
  const long long blocks = (warps + 7) / 8;
  static bool carveout_set = false;
  if (!carveout_set) {  // no shared memory in this kernel: give everything to L1 (gather reuse)
    cudaFuncSetAttribute(nerf_combine_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    carveout_set = true;
  }
  launch_ex(nerf_combine_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      reinterpret_cast<const __nv_bfloat16*>(g), ldg, reinterpret_cast<const __nv_bfloat16*>(hpre),
      gidx, gwgt, vlogit, reinterpret_cast<__nv_bfloat16*>(s), view_softmax, b, n, hw, d, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

extern "C" int cd360_nerf_volrender(const void* feats, const float* raw, const float* dists,
                                    void* rendered, float* fg, float* alphas, float* rgb,
                                    int32_t b, int32_t hw, int32_t d, int32_t c,
                                    cd360_stream_t stream_) {
  if (!feats || !raw || !dists || !rendered || !fg || !alphas || !rgb) return CD360_ERR_NULL;
  if (b <= 0 || hw <= 0 || d <= 0 || d > 32 || c <= 0 || (c & 7)) return CD360_ERR_SHAPE;
  const long long warps = static_cast<long long>(b) * hw;
  const long long blocks = (warps + 7) / 8;
  launch_ex(nerf_volrender_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), 1, 
      reinterpret_cast<const __nv_bfloat16*>(feats), raw, dists,
      reinterpret_cast<__nv_bfloat16*>(rendered), fg, alphas, rgb, b, hw, d, c);
  CD360_CHECK_LAUNCH();
  return CD360_OK;
}

CD360_TL_SETTER(nerf)
