// optim_ops.cu -- total-variation gradient and the three Adam updates on the reference's
// [1,C,X,Y,Z] tensors (rows a12, a13).  Reference: lib/cuda/total_variation_kernel.cu:13-67,
// lib/cuda/adam_upd_kernel.cu:8-132.
//
// Both are pure HBM streaming sweeps (TV: 1 param read + grad RMW, neighbours come from L1/L2;
// Adam: 4 reads + 3 writes per element).  The drop-in versions here keep the reference's one
// element per thread semantics but move 16 bytes per thread where alignment allows; the fused
// TV+Adam sweep of the trainer lives in fused_sweep.cu.
#include "common.cuh"

namespace dvgo {

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

// One TV term: w * clamp(p - p_n, -1, 1) as a plain product (the reference SASS shows predicated
// FMUL + a left-to-right FADD chain, no contraction into the running sum).
__device__ __forceinline__ float tv_term(float w, float p, float pn) {
  return fmul(w, clamp1(fsub(p, pn)));
}

template <bool kDense>
__global__ void __launch_bounds__(256) tv_add_grad_kernel(const float* __restrict__ param,
                                                          float* __restrict__ grad, float wy,
                                                          float wz, int64_t sz_i, int64_t sz_j,
                                                          int64_t sz_k, int64_t N) {
  const int64_t sjk = sz_j * sz_k;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < N;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float g0 = grad[idx];
    if (!kDense && g0 == 0.f) continue;  // :21 sparse mode gates on grad != 0 BEFORE the add
    const int64_t k = idx % sz_k;
    const int64_t j = idx / sz_k % sz_j;
    const int64_t i = idx / sjk % sz_i;
    const float p = param[idx];
    float g = 0.f;  // :25-32, same order; i-axis uses wz (reference quirk, :31-32)
    g = fadd(g, k == 0 ? 0.f : tv_term(wz, p, __ldg(param + idx - 1)));
    g = fadd(g, k == sz_k - 1 ? 0.f : tv_term(wz, p, __ldg(param + idx + 1)));
    g = fadd(g, j == 0 ? 0.f : tv_term(wy, p, __ldg(param + idx - sz_k)));
    g = fadd(g, j == sz_j - 1 ? 0.f : tv_term(wy, p, __ldg(param + idx + sz_k)));
    g = fadd(g, i == 0 ? 0.f : tv_term(wz, p, __ldg(param + idx - sjk)));
    g = fadd(g, i == sz_i - 1 ? 0.f : tv_term(wz, p, __ldg(param + idx + sjk)));
    grad[idx] = fadd(g0, g);  // :33
  }
}

// ---- Adam ----------------------------------------------------------------------------------------
// mode 0: adam_upd (:8-23), 1: masked_adam_upd (:25-40), 2: adam_upd_with_perlr (:42-58).
// Expression tree as in the reference SASS: m = fma(b1, m, (1-b1)*g); v = fma(b2, v, ((1-b2)*g)*g);
// p -= (step_size [* perlr]) * m / (sqrt(v) + eps), IEEE sqrt and divide.
template <int kMode>
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float perlr,
                                         float step_size, float beta1, float beta2, float eps) {
  if (kMode == 1 && g == 0.f) return;  // :35
  m = fma_(beta1, m, fmul(fsub(1.f, beta1), g));
  v = fma_(beta2, v, fmul(fmul(fsub(1.f, beta2), g), g));
  const float num = (kMode == 2) ? fmul(fmul(step_size, perlr), m) : fmul(step_size, m);
  p = fsub(p, fdiv(num, fadd(sqrtf(v), eps)));
}

template <int kMode>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ param,
                                                   const float* __restrict__ grad,
                                                   float* __restrict__ exp_avg,
                                                   float* __restrict__ exp_avg_sq,
                                                   const float* __restrict__ perlr, int64_t N,
                                                   float step_size, float beta1, float beta2,
                                                   float eps, int vec_ok) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n4 = vec_ok ? N / 4 : 0;
  for (int64_t q = tid; q < n4; q += nthreads) {
    float4 p = reinterpret_cast<float4*>(param)[q];
    const float4 g = reinterpret_cast<const float4*>(grad)[q];
    if (kMode == 1 && g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    float4 m = reinterpret_cast<float4*>(exp_avg)[q];
    float4 v = reinterpret_cast<float4*>(exp_avg_sq)[q];
    float4 l = make_float4(1.f, 1.f, 1.f, 1.f);
    if (kMode == 2) l = reinterpret_cast<const float4*>(perlr)[q];
    adam_one<kMode>(p.x, g.x, m.x, v.x, l.x, step_size, beta1, beta2, eps);
    adam_one<kMode>(p.y, g.y, m.y, v.y, l.y, step_size, beta1, beta2, eps);
    adam_one<kMode>(p.z, g.z, m.z, v.z, l.z, step_size, beta1, beta2, eps);
    adam_one<kMode>(p.w, g.w, m.w, v.w, l.w, step_size, beta1, beta2, eps);
    reinterpret_cast<float4*>(param)[q] = p;
    reinterpret_cast<float4*>(exp_avg)[q] = m;
    reinterpret_cast<float4*>(exp_avg_sq)[q] = v;
  }
  for (int64_t i = n4 * 4 + tid; i < N; i += nthreads) {
    float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    const float g = grad[i];
    if (kMode == 1 && g == 0.f) continue;
    adam_one<kMode>(p, g, m, v, kMode == 2 ? perlr[i] : 1.f, step_size, beta1, beta2, eps);
    param[i] = p;
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
  }
}

static inline int grid_for(int64_t n, int threads) {
  const int64_t want = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 16;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Host-side bias correction, reference adam_upd_kernel.cu:72 (float pow/sqrt on the host).
static inline float adam_step_size(int step, float beta1, float beta2, float lr) {
  return lr * sqrtf(1.f - powf(beta2, static_cast<float>(step))) /
         (1.f - powf(beta1, static_cast<float>(step)));
}

template <int kMode>
static int adam_launch(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       const float* perlr, int64_t N, int step, float beta1, float beta2, float lr,
                       float eps, dvgo_stream_t stream) {
  if (N < 0) return DVGO_EINVAL;
  if (N == 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq || (kMode == 2 && !perlr)) return DVGO_EINVAL;
  const float step_size = adam_step_size(step, beta1, beta2, lr);
  const int vec_ok = aligned16(param) && aligned16(grad) && aligned16(exp_avg) &&
                     aligned16(exp_avg_sq) && (kMode != 2 || aligned16(perlr));
  const int64_t work = vec_ok ? (N + 3) / 4 : N;
  adam_kernel<kMode><<<grid_for(work, 256), 256, 0, as_stream(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, perlr, N, step_size, beta1, beta2, eps, vec_ok);
  return launch_status();
}

}  // namespace dvgo

using namespace dvgo;

DVGO_API int dvgo_total_variation_add_grad(const float* param, float* grad, float wx, float wy,
                                           float wz, int dense_mode, int64_t N, int64_t sz_i,
                                           int64_t sz_j, int64_t sz_k, dvgo_stream_t stream) {
  (void)wx;  // unused by the reference as well (total_variation_kernel.cu:31-32)
  if (N < 0 || sz_i <= 0 || sz_j <= 0 || sz_k <= 0) return DVGO_EINVAL;
  if (N == 0) return 0;
  if (!param || !grad) return DVGO_EINVAL;
  wy /= 6;  // :45-47
  wz /= 6;
  const int blocks = grid_for(N, 256);
  if (dense_mode)
    tv_add_grad_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(param, grad, wy, wz, sz_i, sz_j,
                                                                    sz_k, N);
  else
    tv_add_grad_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(param, grad, wy, wz, sz_i,
                                                                     sz_j, sz_k, N);
  return launch_status();
}

DVGO_API int dvgo_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                           int64_t N, int step, float beta1, float beta2, float lr, float eps,
                           dvgo_stream_t stream) {
  return adam_launch<0>(param, grad, exp_avg, exp_avg_sq, nullptr, N, step, beta1, beta2, lr, eps,
                        stream);
}

DVGO_API int dvgo_masked_adam_upd(float* param, const float* grad, float* exp_avg,
                                  float* exp_avg_sq, int64_t N, int step, float beta1, float beta2,
                                  float lr, float eps, dvgo_stream_t stream) {
  return adam_launch<1>(param, grad, exp_avg, exp_avg_sq, nullptr, N, step, beta1, beta2, lr, eps,
                        stream);
}

DVGO_API int dvgo_adam_upd_with_perlr(float* param, const float* grad, float* exp_avg,
                                      float* exp_avg_sq, const float* perlr, int64_t N, int step,
                                      float beta1, float beta2, float lr, float eps,
                                      dvgo_stream_t stream) {
  return adam_launch<2>(param, grad, exp_avg, exp_avg_sq, perlr, N, step, beta1, beta2, lr, eps,
                        stream);
}
