// attention_bwd.cu — C-ABI entry points of the attention backward (training step).
//
// The reference gets these gradients from autograd through xformers.ops.memory_efficient_attention
// (sgm/modules/attention.py:406).  The kernels live in attention_bwd_tcgen05.cu (tcgen05 / TMEM / TMA:
// a dQ + statistics kernel and a dK/dV kernel over 128 x 128 tiles); this file validates the arguments,
// dispatches, and converts the fp32 partial sums of the query-split dK/dV variant.  (The first two
// generations — nvcuda::wmma, then mma.sync m16n8k16 with the scores in registers — were removed once the
// tcgen05 kernels passed the same parity tests: 54.8 -> 46.3 ms per training step, and no legacy HMMA path
// is left in the library.)
#include "cd360_common.cuh"

namespace cd360 {

// attention_bwd_tcgen05.cu.  which: bit 0 = dQ + statistics, bit 1 = dK / dV (nsplit > 1: partial sums
// into kv_acc).
int attention_bwd_tcgen05(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                          const void* o, long long ldo, const void* dout, long long lddo, void* dq, long long lddq,
                          void* dk, long long lddk, void* dv, long long lddv, float* lse, float* dsum, float* kv_acc,
                          int batch, int heads, int nq, int nkv, int nsplit, int which, cudaStream_t stream);

constexpr int AB_T = 64;   // head dim

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
  return attention_bwd_tcgen05(q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, dq, lddq, dk, lddk, dv, lddv, lse, dsum,
                               nullptr, batch, heads, nq, nkv, 1, dk != nullptr ? 3 : 1,
                               reinterpret_cast<cudaStream_t>(stream_));
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
  const int rc = attention_bwd_tcgen05(q, ldq, k, ldk, v, ldv, nullptr, 0, dout, lddo, nullptr, 0, nullptr, 0, nullptr,
                                       0, const_cast<float*>(lse), const_cast<float*>(dsum), kv_acc, batch, heads, nq,
                                       nkv, nsplit, 2, stream);
  if (rc != CD360_OK) return rc;
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
