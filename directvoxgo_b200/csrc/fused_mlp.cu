// fused_mlp.cu -- the rgbnet MLP (lib/dvgo.py:123-131, applied at :524-539) on the 5th-generation
// tensor cores: tcgen05.mma kind::f16 (FP16 operands, fp32 accumulators in TMEM), operands staged in
// shared memory in the canonical no-swizzle core-matrix layout (tc_common.cuh).
//
// Numerics: fp32 master weights / features / outputs; GEMM operands rounded to FP16 (10-bit mantissa,
// the same as TF32 -- what the reference's nn.Linear uses on any tensor-core GPU under its pinned
// PyTorch 1.8.1, where allow_tf32 defaulted to True; README.md:32) on the way into shared memory;
// products accumulated in fp32.  Backward operands are scaled by a power of two so that small
// gradients stay in FP16's normal range, and unscaled exactly in the fp32 epilogue.
// Stated tolerance vs the exact-fp32 path: 2e-3 abs on rgb (tests/test_gpu_mlp.py).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/dvgo_b200_fused.h"

namespace dvgo {

using namespace tc;

// ----------------------------------------------------------------------------------------------------
// Self test: D[128,N] = A[128,K] * B[N,K]^T for every operand orientation the MLP kernels use.
//   a_mn = 0: A given as [128][K] (K contiguous)   -> K-major tile (rows = m, cols = k)
//   a_mn = 1: A given as [K][128] (M contiguous)   -> MN-major tile (rows = k, cols = m)
//   b_mn = 0: B given as [N][K]                    -> K-major tile (rows = n, cols = k)
//   b_mn = 1: B given as [K][N]                    -> MN-major tile (rows = k, cols = n)
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A,
                                                          const float* __restrict__ B,
                                                          float* __restrict__ D, int N, int K, int a_mn,
                                                          int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a_rows = a_mn ? K : 128, a_cols = a_mn ? 128 : K;
  const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  uint8_t* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // 1 KiB-aligned operand tiles
  uint8_t* sB = sA + ((tile_bytes(a_rows, a_cols) + 1023) / 1024) * 1024;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
  for (int i = tid; i < a_rows * a_cols; i += blockDim.x) {
    const int r = i / a_cols, c = i % a_cols;
    *reinterpret_cast<__half*>(sA + tile_off(r, c, a_cols)) = __float2half_rn(A[i]);
  }
  for (int i = tid; i < b_rows * b_cols; i += blockDim.x) {
    const int r = i / b_cols, c = i % b_cols;
    *reinterpret_cast<__half*>(sB + tile_off(r, c, b_cols)) = __float2half_rn(B[i]);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, a_mn, b_mn);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint32_t a_step = a_mn ? 2 * group_stride(a_cols) : 256u;
    const uint32_t b_step = b_mn ? 2 * group_stride(b_cols) : 256u;
    for (int k = 0; k < K / kMmaK; ++k) {
      const uint64_t da = a_mn ? desc_mnmajor(a0 + k * a_step, a_cols) : desc_kmajor(a0 + k * a_step, a_cols);
      const uint64_t db = b_mn ? desc_mnmajor(b0 + k * b_step, b_cols) : desc_kmajor(b0 + k * b_step, b_cols);
      mma_f16(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Descriptor probe (debug aid): A is staged K-major (validated orientation); the B region of shared
// memory is filled VERBATIM from Braw (word i -> byte 4*i) and described with caller-chosen
// LBO / SBO / per-k-step advance, so with A = unit vectors and Braw = ramp the output D reveals which
// shared-memory word the tensor core reads for each logical B[n][k].
__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ A,
                                                       const float* __restrict__ Braw, float* __restrict__ D,
                                                       int N, int K, int b_mn, uint32_t lbo, uint32_t sbo,
                                                       uint32_t kstep, int nwords) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* sB = sA + ((tile_bytes(128, K) + 1023) / 1024) * 1024;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
  for (int i = tid; i < 128 * K; i += blockDim.x)
    *reinterpret_cast<__half*>(sA + tile_off(i / K, i % K, K)) = __float2half_rn(A[i]);
  for (int i = tid; i < nwords; i += blockDim.x) reinterpret_cast<__half*>(sB)[i] = __float2half_rn(Braw[i]);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, 0, b_mn);
    for (int k = 0; k < K / kMmaK; ++k)
      mma_f16(tmem, desc_kmajor(smem_u32(sA) + k * 256u, K), make_desc(smem_u32(sB) + k * kstep, lbo, sbo),
               idesc, k > 0 ? 1u : 0u);
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn,
                              dvgo_stream_t stream) {
  if (!A || !B || !D || N < 16 || N > 256 || N % 16 || K < 16 || K > 128 || K % 16) return DVGO_EINVAL;
  const int a_rows = a_mn ? K : 128, a_cols = a_mn ? 128 : K;
  const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  const size_t bytes = ((tc::tile_bytes(a_rows, a_cols) + 1023) / 1024) * 1024 + tc::tile_bytes(b_rows, b_cols) + 2048;
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_selftest_kernel<<<1, 128, bytes, as_stream(stream)>>>(A, B, D, N, K, a_mn, b_mn);
  return launch_status();
}

DVGO_API int dvgo_tc_probe(const float* A, const float* Braw, float* D, int N, int K, int b_mn, int lbo, int sbo,
                           int kstep, int nwords, dvgo_stream_t stream) {
  if (!A || !Braw || !D || N < 16 || N > 256 || N % 16 || K < 16 || K > 128 || K % 16 || nwords > 32768)
    return DVGO_EINVAL;
  const size_t bytes = ((tc::tile_bytes(128, K) + 1023) / 1024) * 1024 + nwords * 4 + 2048;
  cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  tc_probe_kernel<<<1, 128, bytes, as_stream(stream)>>>(A, Braw, D, N, K, b_mn, lbo, sbo, kstep, nwords);
  return launch_status();
}
