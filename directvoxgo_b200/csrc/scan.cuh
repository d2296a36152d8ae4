// scan.cuh -- single-CTA inclusive scan of int64 counts (N_steps -> N_steps_cumsum of sample_pts_on_rays,
// reference lib/cuda/render_utils_kernel.cu:204).  Header so that the float32 (ray_ops.cu) and the float64
// (f64_ops.cu) translation units each launch their own copy (no relocatable device code in this library).
#pragma once
#include "common.cuh"

namespace dvgo {

// Single-CTA inclusive scan of int64 counts.  n_rays is 8192 per training step (64 Ki at most in
// the configs), i.e. 2-16 trips of a 1024-thread CTA: cheaper than a multi-kernel device scan.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
static __global__ void __launch_bounds__(kScanThreads) inclusive_scan_i64_kernel(
    const int64_t* __restrict__ in, int n, int64_t* __restrict__ out) {
  __shared__ int64_t warp_sums[kScanThreads / kWarp];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kScanThreads * kScanItems) {
    int64_t v[kScanItems];
    int64_t local = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const int i = base + tid * kScanItems + k;
      v[k] = (i < n) ? in[i] : 0;
      local += v[k];
    }
    int64_t incl = local;  // warp inclusive scan of the per-thread sums
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int64_t up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += up;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int64_t ws = warp_sums[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int64_t up = __shfl_up_sync(0xffffffffu, ws, off);
        if (lane >= off) ws += up;
      }
      warp_sums[lane] = ws;  // inclusive over warps
    }
    __syncthreads();
    const int64_t carry = carry_s;
    int64_t run = carry + (wid ? warp_sums[wid - 1] : 0) + (incl - local);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      const int i = base + tid * kScanItems + k;
      run += v[k];
      if (i < n) out[i] = run;
    }
    __syncthreads();
    if (tid == kScanThreads - 1) carry_s = run;
    __syncthreads();
  }
}

}  // namespace dvgo
