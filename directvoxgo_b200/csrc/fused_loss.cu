// fused_loss.cu -- survivor-stream kernels between the two march passes: direct colour, compositing,
// per-ray loss / gradient seeds, per-sample gradient seeds.  All of them read the survivor count
// from device memory (counters[0]) so the host never synchronises on the data-dependent M4.
//
// Reference: lib/dvgo.py:512-514 (sigmoid colour), :554-576 (segment_coo compositing),
// run.py:377-386 (loss), SURVEY.md appendix C (gradient flow).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/dvgo_b200_fused.h"

namespace dvgo {

__device__ __forceinline__ int64_t survivor_count(const int32_t* counters, int64_t cap) {
  const int64_t n = counters[0];
  return n < cap ? n : cap;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) rgb_direct_kernel(const float* __restrict__ feat,
                                                         const int32_t* __restrict__ counters,
                                                         int64_t cap, float* __restrict__ rgb) {
  const int64_t n = survivor_count(counters, cap) * 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    rgb[i] = sigmoidf(feat[i]);
}

__global__ void __launch_bounds__(256) rgb_direct_bwd_kernel(const float* __restrict__ rgb,
                                                             const float* __restrict__ d_rgb,
                                                             const int32_t* __restrict__ counters,
                                                             int64_t cap, float* __restrict__ d_feat) {
  const int64_t n = survivor_count(counters, cap) * 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float s = rgb[i];
    d_feat[i] = d_rgb[i] * s * (1.f - s);
  }
}

// Segmented sums by ray.  Survivors of one warp-chunk of march_fwd are contiguous and share a ray,
// so runs of equal ray ids are long: reduce runs inside the warp with shuffles, one atomic per run.
__global__ void __launch_bounds__(256) composite_kernel(
    const float* __restrict__ rgb, const float* __restrict__ s_weight,
    const int32_t* __restrict__ s_ray, const int32_t* __restrict__ s_slot,
    const int32_t* __restrict__ ray_off, const int32_t* __restrict__ counters, int64_t cap,
    float* __restrict__ rgb_acc, float* __restrict__ depth_acc) {
  const int lane = threadIdx.x & 31;
  const int64_t n = survivor_count(counters, cap);
  const int64_t n_round = (n + 31) / 32 * 32;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < n_round;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool valid = p < n;
    const int key = valid ? s_ray[p] : -1;
    const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head_any = lane == 0 || key_prev != key;
    const bool head = valid && head_any;
    // Run identity, not key equality: chunks of one ray get their stream positions from an atomicAdd in
    // march_fwd, so a window may read A.. B.. A.. and the two A runs must stay separate partial sums.
    const unsigned heads = __ballot_sync(0xffffffffu, head_any);
    const int seg = __popc(heads & (0xffffffffu >> (31 - lane)));
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const float w = s_weight[p];
      v[0] = w * rgb[3 * p]; v[1] = w * rgb[3 * p + 1]; v[2] = w * rgb[3 * p + 2];  // lib/dvgo.py:555
      if (depth_acc) v[3] = w * static_cast<float>(s_slot[p] - ray_off[key]);       // :572 w * step_id
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int sd = __shfl_down_sync(0xffffffffu, seg, off);
      const bool take = (lane + off < 32) && (sd == seg);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float vd = __shfl_down_sync(0xffffffffu, v[c], off);
        if (take) v[c] += vd;
      }
    }
    if (head) {
      atomicAdd(rgb_acc + 3 * key, v[0]);
      atomicAdd(rgb_acc + 3 * key + 1, v[1]);
      atomicAdd(rgb_acc + 3 * key + 2, v[2]);
      if (depth_acc) atomicAdd(depth_acc + key, v[3]);
    }
  }
}

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sm[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.f;
  if (wid == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  }
  return v;  // valid in thread 0
}

__global__ void __launch_bounds__(256) ray_finish_kernel(
    float* __restrict__ rgb_acc, const float* __restrict__ alphainv_last,
    const float* __restrict__ target, float bg, int n_rays, int n_global, float w_main, float w_ent,
    float* __restrict__ G, float* __restrict__ g_last, float* __restrict__ loss_acc) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (r < n_rays) {
    const float last = alphainv_last[r];
    float m[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      m[c] = rgb_acc[3 * r + c] + last * bg;  // lib/dvgo.py:559
      rgb_acc[3 * r + c] = m[c];
    }
    if (target) {
      const float inv3n = 1.f / (3.f * static_cast<float>(n_global));
      float gsum = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = m[c] - target[3 * r + c];
        loss += w_main * d * d * inv3n;                      // run.py:377 mse mean over 3N
        const float gc = w_main * 2.f * d * inv3n;
        G[3 * r + c] = gc;
        gsum += gc;
      }
      float gl = bg * gsum;
      if (w_ent > 0.f) {                                     // run.py:379-382
        const float p = fminf(fmaxf(last, 1e-6f), 1.f - 1e-6f);
        const float lp = logf(p), lq = logf(1.f - p);
        loss += -w_ent * (p * lp + (1.f - p) * lq) / static_cast<float>(n_global);
        if (last >= 1e-6f && last <= 1.f - 1e-6f)            // clamp passes gradient inside its range
          gl += -w_ent * (lp - lq) / static_cast<float>(n_global);
      }
      g_last[r] = gl;
    }
  }
  if (target && loss_acc) {
    const float s = block_sum(loss);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(loss_acc, s);
  }
}

__global__ void __launch_bounds__(256) sample_grad_kernel(
    const float* __restrict__ rgb, const float* __restrict__ s_weight,
    const int32_t* __restrict__ s_ray, const float* __restrict__ G, const float* __restrict__ target,
    const int32_t* __restrict__ counters, int64_t cap, int n_global, float w_per,
    float* __restrict__ d_rgb, float* __restrict__ d_w, float* __restrict__ loss_acc, uint8_t* __restrict__ dzt,
    float grad_scale) {
  const int64_t n = survivor_count(counters, cap);
  const int64_t rows = dzt ? (n + 255) / 256 * 256 : n;   // the dZ3 tiles are zero up to the end of the last tile pair
  const float inv_n = 1.f / static_cast<float>(n_global);
  float loss = 0.f;
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < rows;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    uint8_t* t = dzt ? dzt + (p >> 7) * tc::tile_bytes(128, 16) + tc::tile_off(static_cast<int>(p & 127), 0, 16) : nullptr;
    if (p >= n) {   // (columns 8..15 of every row are never written: the buffer is allocated zeroed)
      *reinterpret_cast<uint4*>(t) = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const int r = s_ray[p];
    const float w = s_weight[p];
    float dw = 0.f;
    float z[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = rgb[3 * p + c];
      const float g = G[3 * r + c];
      float dr = w * g;
      if (w_per > 0.f) {
        const float d = x - target[3 * r + c];
        dr += w_per * 2.f * w * d * inv_n;        // run.py:384-386 (weights detached)
        loss += w_per * w * d * d * inv_n;
      }
      if (d_rgb) d_rgb[3 * p + c] = dr;
      z[c] = dr * x * (1.f - x);                  // through the sigmoid: what the rgbnet backward starts from
      dw += g * x;
    }
    d_w[p] = dw;
    if (t) {   // [row][0..2] = S * dZ3 as saturated fp16, [3..15] = 0
      const uint2 h = tc::pack4(make_float4(z[0] * grad_scale, z[1] * grad_scale, z[2] * grad_scale, 0.f));
      *reinterpret_cast<uint4*>(t) = make_uint4(h.x, h.y, 0u, 0u);
    }
  }
  if (loss_acc && w_per > 0.f) {
    const float s = block_sum(loss);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(loss_acc, s);
  }
}

__global__ void __launch_bounds__(256) zero_words_kernel(uint32_t* __restrict__ p, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = 0u;
}

// Start of a fused step: fold the previous call's counters into the running statistics, then zero the block.
// zblock = counters[2] (survivor count, overflow flag) | accumulators...; stats (int64): [0] += survivors of the
// previous call, [1] += 1 (calls), [2] |= overflow flag, [3] = max survivors of a call.  No host synchronisation:
// the host reads `stats` only where it synchronises anyway (bench roofline, sync_to_model, checkpoints).
__global__ void __launch_bounds__(256) step_begin_kernel(uint32_t* __restrict__ z, int64_t n,
                                                         long long* __restrict__ stats) {
  const int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i0 == 0) {   // words 0 and 1 (the counters) belong to thread 0 alone: read, fold, then zero, in program order
    if (stats) {
      const long long surv = static_cast<int32_t>(z[0]);
      stats[0] += surv;
      stats[1] += 1;
      stats[2] |= static_cast<long long>(z[1]);   // bit 0: capacity overflow, bit 1: non-finite rgbnet value
      if (surv > stats[3]) stats[3] = surv;
    }
    z[0] = 0u;
    z[1] = 0u;
  }
  for (int64_t i = i0; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    if (i >= 2) z[i] = 0u;
}

// Survivor-stream kernels size their grid for the capacity (the count lives on the device).
static inline int stream_grid(int64_t cap, int threads) {
  const int64_t want = (cap + threads - 1) / threads;
  const int64_t lim = static_cast<int64_t>(kNumSMs) * 16;
  return static_cast<int>(want < lim ? (want > 0 ? want : 1) : lim);
}

// View-direction positional encoding for the tensor-core rgbnet, written straight into its padded table
// (lib/dvgo.py:524-525: emb = (viewdirs[..., None] * viewfreq).flatten(-2); cat[viewdirs, sin(emb), cos(emb)]):
// columns [0,3) viewdirs, [3, 3+3F) sin, [3+3F, 3+6F) cos, column P = 3+6F holds 1 (carries b1 through the first
// GEMM), the rest up to `stride` is 0.  One launch instead of ~8 elementwise torch kernels per step / render chunk.
__global__ void __launch_bounds__(256) view_embedding_kernel(const float* __restrict__ viewdirs,
                                                             const float* __restrict__ freq, int F, int64_t n,
                                                             int stride, float* __restrict__ out,
                                                             __half* __restrict__ rows16, int C, int K1) {
  const int P = 3 + 6 * F;
  const int64_t total = n * stride;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / stride;
    const int c = static_cast<int>(i - r * stride);
    float v;
    if (c < 3) {
      v = viewdirs[3 * r + c];
    } else if (c < P) {
      const int k = (c - 3) % (3 * F);          // index into emb = [x*f0..x*fF-1, y*f0.., z*f0..]
      const float e = fmul(viewdirs[3 * r + k / F], __ldg(freq + k % F));
      v = (c - 3) < 3 * F ? sinf(e) : cosf(e);
    } else {
      v = c == P ? 1.f : 0.f;
    }
    out[i] = v;
    // the same value as the tail of the ray's X~ row (fp16, columns [C, C + stride) of K1): what k0_gather_tiles copies
    if (rows16) rows16[r * K1 + C + c] = __low2half(tc::pack2_sat(v, 0.f));
  }
}

}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_fused_rgb_direct(const float* feat, const int32_t* counters, int64_t surv_cap,
                                   float* rgb, dvgo_stream_t stream) {
  if (surv_cap < 0 || !feat || !counters || !rgb) return DVGO_EINVAL;
  rgb_direct_kernel<<<stream_grid(surv_cap * 3, 256), 256, 0, as_stream(stream)>>>(feat, counters,
                                                                                   surv_cap, rgb);
  return launch_status();
}

DVGO_API int dvgo_fused_rgb_direct_bwd(const float* rgb, const float* d_rgb, const int32_t* counters,
                                       int64_t surv_cap, float* d_feat, dvgo_stream_t stream) {
  if (surv_cap < 0 || !rgb || !d_rgb || !counters || !d_feat) return DVGO_EINVAL;
  rgb_direct_bwd_kernel<<<stream_grid(surv_cap * 3, 256), 256, 0, as_stream(stream)>>>(
      rgb, d_rgb, counters, surv_cap, d_feat);
  return launch_status();
}

DVGO_API int dvgo_fused_composite(const float* rgb, const float* s_weight, const int32_t* s_ray,
                                  const int32_t* s_slot, const int32_t* ray_off,
                                  const int32_t* counters, int64_t surv_cap, float* rgb_acc,
                                  float* depth_acc, dvgo_stream_t stream) {
  if (surv_cap < 0 || !rgb || !s_weight || !s_ray || !counters || !rgb_acc ||
      (depth_acc && (!s_slot || !ray_off)))
    return DVGO_EINVAL;
  composite_kernel<<<stream_grid(surv_cap, 256), 256, 0, as_stream(stream)>>>(
      rgb, s_weight, s_ray, s_slot, ray_off, counters, surv_cap, rgb_acc, depth_acc);
  return launch_status();
}

DVGO_API int dvgo_fused_ray_finish(float* rgb_acc, const float* alphainv_last, const float* target,
                                   float bg, int n_rays, int n_global, float weight_main,
                                   float weight_entropy_last, float* G, float* g_last,
                                   float* loss_acc, dvgo_stream_t stream) {
  if (n_rays < 0 || n_global <= 0) return DVGO_EINVAL;
  if (n_rays == 0) return 0;
  if (!rgb_acc || !alphainv_last || (target && (!G || !g_last))) return DVGO_EINVAL;
  ray_finish_kernel<<<blocks_for(n_rays, 256), 256, 0, as_stream(stream)>>>(
      rgb_acc, alphainv_last, target, bg, n_rays, n_global, weight_main, weight_entropy_last, G,
      g_last, loss_acc);
  return launch_status();
}

DVGO_API int dvgo_fused_sample_grad(const float* rgb, const float* s_weight, const int32_t* s_ray,
                                    const float* G, const float* target, const int32_t* counters,
                                    int64_t surv_cap, int n_global, float weight_rgbper, float* d_rgb,
                                    float* d_w, float* loss_acc, void* dzt, float grad_scale,
                                    dvgo_stream_t stream) {
  if (surv_cap < 0 || n_global <= 0 || !rgb || !s_weight || !s_ray || !G || !counters || (!d_rgb && !dzt) ||
      !d_w || (weight_rgbper > 0.f && !target) || (dzt && !(grad_scale > 0.f)))
    return DVGO_EINVAL;   // (d_rgb may be NULL when the dZ3 tiles are the only consumer: the tensor-core rgbnet)
  sample_grad_kernel<<<stream_grid(surv_cap, 256), 256, 0, as_stream(stream)>>>(
      rgb, s_weight, s_ray, G, target, counters, surv_cap, n_global, weight_rgbper, d_rgb, d_w,
      loss_acc, static_cast<uint8_t*>(dzt), grad_scale);
  return launch_status();
}

DVGO_API int dvgo_fused_zero(void* ptr, int64_t n_words, dvgo_stream_t stream) {
  if (n_words < 0) return DVGO_EINVAL;
  if (n_words == 0) return 0;
  if (!ptr) return DVGO_EINVAL;
  zero_words_kernel<<<stream_grid(n_words, 256), 256, 0, as_stream(stream)>>>(
      static_cast<uint32_t*>(ptr), n_words);
  return launch_status();
}

DVGO_API int dvgo_fused_step_begin(void* zblock, int64_t n_words, long long* stats, dvgo_stream_t stream) {
  if (n_words < 2 || !zblock) return DVGO_EINVAL;
  step_begin_kernel<<<stream_grid(n_words, 256), 256, 0, as_stream(stream)>>>(static_cast<uint32_t*>(zblock), n_words,
                                                                              stats);
  return launch_status();
}

DVGO_API int dvgo_view_embedding(const float* viewdirs, const float* freq, int n_freq, int64_t n_rays, int stride,
                                 float* out, void* rows16, int C, dvgo_stream_t stream) {
  if (n_rays < 0 || n_freq < 0 || stride < 3 + 6 * n_freq + 1 || !out) return DVGO_EINVAL;
  if (rows16 && (C < 0 || C + stride > 64)) return DVGO_EINVAL;
  const int K1 = ((C + stride + 15) / 16) * 16;
  if (n_rays == 0) return 0;
  if (!viewdirs || (n_freq > 0 && !freq)) return DVGO_EINVAL;
  const int64_t want = (n_rays * stride + 255) / 256;
  const int blocks = static_cast<int>(want < kNumSMs * 8 ? want : kNumSMs * 8);
  view_embedding_kernel<<<blocks, 256, 0, as_stream(stream)>>>(viewdirs, freq, n_freq, n_rays, stride, out,
                                                               static_cast<__half*>(rows16), C, K1);
  return launch_status();
}
