// mlp_binding.cpp -- torch adaptor for the tensor-core rgbnet kernels (fused_mlp.cu).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../include/dvgo_b200_fused.h"

namespace {
using torch::Tensor;

inline dvgo_stream_t cur_stream() {
  return reinterpret_cast<dvgo_stream_t>(at::cuda::getCurrentCUDAStream().stream());
}
inline void chkf(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kFloat32, name,
              " must be a contiguous float32 CUDA tensor");
}
inline void rc_check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "dvgo_b200 mlp: ", what, " failed with code ", rc);
}

Tensor tc_selftest(Tensor A, Tensor B, int N, int K, bool a_mn, bool b_mn) {
  chkf(A, "A"); chkf(B, "B");
  TORCH_CHECK(A.numel() == 128 * K && B.numel() == (int64_t)N * K, "A must hold 128*K, B N*K elements");
  const c10::cuda::CUDAGuard guard(A.device());
  auto D = torch::empty({128, N}, A.options());
  rc_check(dvgo_tc_selftest(A.data_ptr<float>(), B.data_ptr<float>(), D.data_ptr<float>(), N, K, a_mn, b_mn,
                            cur_stream()), "tc_selftest");
  return D;
}

Tensor tc_probe(Tensor A, Tensor Braw, int N, int K, bool b_mn, int lbo, int sbo, int kstep) {
  chkf(A, "A"); chkf(Braw, "Braw");
  const c10::cuda::CUDAGuard guard(A.device());
  auto D = torch::empty({128, N}, A.options());
  rc_check(dvgo_tc_probe(A.data_ptr<float>(), Braw.data_ptr<float>(), D.data_ptr<float>(), N, K, b_mn, lbo, sbo,
                         kstep, static_cast<int>(Braw.numel()), cur_stream()), "tc_probe");
  return D;
}

std::vector<Tensor> tc_ts_probe(Tensor A, Tensor B, int N, int K, bool b_mn, int reps) {
  chkf(A, "A"); chkf(B, "B");
  TORCH_CHECK(A.numel() == 128 * K && B.numel() == (int64_t)N * K, "A must hold 128*K, B N*K elements");
  const c10::cuda::CUDAGuard guard(A.device());
  auto D = torch::empty({128, N}, A.options());
  auto cyc = torch::zeros({2}, torch::TensorOptions().dtype(torch::kInt64).device(A.device()));
  rc_check(dvgo_tc_ts_probe(A.data_ptr<float>(), B.data_ptr<float>(), D.data_ptr<float>(), N, K, b_mn, reps,
                            reinterpret_cast<long long*>(cyc.data_ptr<int64_t>()), cur_stream()), "tc_ts_probe");
  return {D, cyc};
}

Tensor tc_contention(int mma_mode, int n_mma, int simt_mode, int reps) {
  auto out = torch::zeros({4}, torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA));
  auto gbuf = torch::ones({1 << 18}, torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA));
  rc_check(dvgo_tc_contention(mma_mode, n_mma, simt_mode, reps, gbuf.data_ptr(),
                              reinterpret_cast<long long*>(out.data_ptr<int64_t>()), cur_stream()), "tc_contention");
  return out;
}

Tensor tc_ldtm_rate(int nwarps, int reps, int mode) {
  auto out = torch::zeros({3}, torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA));
  rc_check(dvgo_tc_ldtm_rate(nwarps, reps, mode, reinterpret_cast<long long*>(out.data_ptr<int64_t>()), cur_stream()),
           "tc_ldtm_rate");
  return out;
}

Tensor tc_rate(int ctas, int N, int ksteps, int reps, bool a_mn, bool b_mn, int a_lbo, int a_sbo, int a_kstep, int b_lbo,
               int b_sbo, int b_kstep, int layout, int n_accum) {
  auto out = torch::zeros({ctas}, torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA));
  rc_check(dvgo_tc_rate(ctas, N, ksteps, reps, a_mn, b_mn, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, layout, n_accum,
                        reinterpret_cast<long long*>(out.data_ptr<int64_t>()), cur_stream()), "tc_rate");
  return out;
}

// C < 0: the fp32 table only.  C >= 0: also the rays' share of the X~ rows as fp16 ([N][K1] halves, zero elsewhere).
std::vector<Tensor> view_embedding(Tensor viewdirs, Tensor freq, int stride, int C) {
  chkf(viewdirs, "viewdirs"); chkf(freq, "freq");
  TORCH_CHECK(viewdirs.dim() == 2 && viewdirs.size(1) == 3, "viewdirs must be [N,3]");
  const c10::cuda::CUDAGuard guard(viewdirs.device());
  auto out = torch::empty({viewdirs.size(0), stride}, viewdirs.options());
  Tensor rows16;
  if (C >= 0) rows16 = torch::zeros({viewdirs.size(0), ((C + stride + 15) / 16) * 16}, viewdirs.options().dtype(torch::kHalf));
  rc_check(dvgo_view_embedding(viewdirs.data_ptr<float>(), freq.numel() ? freq.data_ptr<float>() : nullptr,
                               static_cast<int>(freq.numel()), viewdirs.size(0), stride, out.data_ptr<float>(),
                               C >= 0 ? rows16.data_ptr() : nullptr, C >= 0 ? C : 0, cur_stream()), "view_embedding");
  if (C >= 0) return {out, rows16};
  return {out};
}

inline void chki(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kInt32, name,
              " must be a contiguous int32 CUDA tensor");
}

// params / grads: flat fp32 buffers laid out [W1 (width*d_in) | b1 | W2 | b2 | W3 (3*width) | b3]
struct Offsets { int64_t W1, b1, W2, b2, W3, b3, total; };
inline Offsets offsets(int d_in, int width) {
  Offsets o;
  o.W1 = 0; o.b1 = o.W1 + (int64_t)width * d_in; o.W2 = o.b1 + width; o.b2 = o.W2 + (int64_t)width * width;
  o.W3 = o.b2 + width; o.b3 = o.W3 + 3 * width; o.total = o.b3 + 3;
  return o;
}

// params (fp32 masters) -> wpack (fp16 operand tiles, uint8 tensor of mlp_wpack_bytes(C, pe_stride) bytes)
int64_t mlp_wpack_bytes(int C, int pe_stride) { return dvgo_mlp_wpack_bytes(C, pe_stride); }

void mlp_pack(Tensor params, int C, int P, int pe_stride, int width, Tensor wpack) {
  chkf(params, "params");
  const Offsets o = offsets(C + P, width);
  TORCH_CHECK(params.numel() == o.total, "params has the wrong size");
  TORCH_CHECK(wpack.is_cuda() && wpack.is_contiguous() && wpack.scalar_type() == torch::kUInt8 &&
              wpack.numel() >= dvgo_mlp_wpack_bytes(C, pe_stride), "wpack: uint8 CUDA tensor of mlp_wpack_bytes bytes");
  const c10::cuda::CUDAGuard guard(params.device());
  const float* p = params.data_ptr<float>();
  rc_check(dvgo_mlp_pack_weights(C, P, pe_stride, p + o.W1, p + o.b1, p + o.W2, p + o.b2, p + o.W3, p + o.b3, width,
                                 wpack.data_ptr(), cur_stream()), "mlp_pack_weights");
}

// wpack = None: pack into a temporary (tests / one-off calls); the trainer keeps a persistent pack per step
Tensor ensure_pack(c10::optional<Tensor> wpack, const Tensor& params, int C, int P, int pe_stride, int width) {
  if (wpack.has_value()) return *wpack;
  auto t = torch::empty({dvgo_mlp_wpack_bytes(C, pe_stride)}, params.options().dtype(torch::kUInt8));
  mlp_pack(params, C, P, pe_stride, width, t);
  return t;
}

inline void chku8(const Tensor& t, int64_t bytes, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == torch::kUInt8 && t.numel() >= bytes &&
              reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0, name, ": 16-byte aligned uint8 CUDA tensor of >= ", bytes, " bytes");
}

// ---- survivor tiles (the kernels' input format; include/dvgo_b200_fused.h) ----
int64_t mlp_xtile_bytes(int64_t cap, int C, int pe_stride) { return dvgo_mlp_xtile_bytes(cap, C, pe_stride); }
int64_t mlp_dztile_bytes(int64_t cap) { return dvgo_mlp_dztile_bytes(cap); }

void mlp_pack_x(Tensor feat, Tensor s_ray, Tensor pe, int P, Tensor counters, Tensor xt) {
  chkf(feat, "feat"); chki(s_ray, "s_ray"); chkf(pe, "pe"); chki(counters, "counters");
  const int C = feat.size(1), pe_stride = pe.size(1);
  const int64_t cap = s_ray.numel();
  TORCH_CHECK(feat.size(0) >= cap, "feat shorter than the stream capacity");
  chku8(xt, dvgo_mlp_xtile_bytes(cap, C, pe_stride), "xt");
  const c10::cuda::CUDAGuard guard(feat.device());
  rc_check(dvgo_mlp_pack_x(feat.data_ptr<float>(), C, s_ray.data_ptr<int32_t>(), pe.data_ptr<float>(), P, pe_stride,
                           counters.data_ptr<int32_t>(), cap, xt.data_ptr(), cur_stream()), "mlp_pack_x");
}

void mlp_pack_dz(Tensor rgb, Tensor d_rgb, double grad_scale, Tensor counters, Tensor dzt) {
  chkf(rgb, "rgb"); chkf(d_rgb, "d_rgb"); chki(counters, "counters");
  const int64_t cap = rgb.numel() / 3;
  TORCH_CHECK(d_rgb.numel() >= cap * 3, "d_rgb shorter than rgb");
  chku8(dzt, dvgo_mlp_dztile_bytes(cap), "dzt");
  const c10::cuda::CUDAGuard guard(rgb.device());
  rc_check(dvgo_mlp_pack_dz(rgb.data_ptr<float>(), d_rgb.data_ptr<float>(), static_cast<float>(grad_scale),
                            counters.data_ptr<int32_t>(), cap, dzt.data_ptr(), cur_stream()), "mlp_pack_dz");
}

void mlp_fwd_tiles(Tensor xt, int C, int P, int pe_stride, Tensor counters, int64_t cap, Tensor wpack, Tensor rgb,
                   c10::optional<Tensor> timeline) {
  chki(counters, "counters"); chkf(rgb, "rgb");
  chku8(xt, dvgo_mlp_xtile_bytes(cap, C, pe_stride), "xt");
  chku8(wpack, dvgo_mlp_wpack_bytes(C, pe_stride), "wpack");
  TORCH_CHECK(rgb.numel() >= cap * 3, "rgb too small");
  const c10::cuda::CUDAGuard guard(xt.device());
  rc_check(dvgo_mlp_fwd_timed(xt.data_ptr(), C, P, pe_stride, counters.data_ptr<int32_t>(), cap, wpack.data_ptr(),
                              rgb.data_ptr<float>(),
                              timeline.has_value() ? reinterpret_cast<long long*>(timeline->data_ptr<int64_t>()) : nullptr,
                              cur_stream()), "mlp_fwd");
}

void mlp_bwd_tiles(Tensor xt, Tensor dzt, int C, int P, int pe_stride, Tensor counters, int64_t cap, Tensor wpack,
                   double grad_scale, Tensor d_feat, Tensor grads, c10::optional<Tensor> timeline) {
  chki(counters, "counters"); chkf(d_feat, "d_feat"); chkf(grads, "grads");
  chku8(xt, dvgo_mlp_xtile_bytes(cap, C, pe_stride), "xt");
  chku8(dzt, dvgo_mlp_dztile_bytes(cap), "dzt");
  chku8(wpack, dvgo_mlp_wpack_bytes(C, pe_stride), "wpack");
  const Offsets o = offsets(C + P, 128);
  TORCH_CHECK(grads.numel() == o.total, "grads has the wrong size");
  TORCH_CHECK(d_feat.numel() >= cap * C, "d_feat too small");
  const c10::cuda::CUDAGuard guard(xt.device());
  float* g = grads.data_ptr<float>();
  rc_check(dvgo_mlp_bwd_timed(xt.data_ptr(), dzt.data_ptr(), C, P, pe_stride, counters.data_ptr<int32_t>(), cap,
                              wpack.data_ptr(), static_cast<float>(grad_scale), d_feat.data_ptr<float>(), g + o.W1,
                              g + o.b1, g + o.W2, g + o.b2, g + o.W3, g + o.b3,
                              timeline.has_value() ? reinterpret_cast<long long*>(timeline->data_ptr<int64_t>()) : nullptr,
                              cur_stream()), "mlp_bwd");
}

// ---- the same on fp32 streams: tiles built into temporaries first (tests, tools, one-off calls) ----
Tensor tiles_from_streams(const Tensor& feat, const Tensor& s_ray, const Tensor& pe, int P, const Tensor& counters) {
  auto xt = torch::zeros({dvgo_mlp_xtile_bytes(s_ray.numel(), feat.size(1), pe.size(1))},
                         feat.options().dtype(torch::kUInt8));
  mlp_pack_x(feat, s_ray, pe, P, counters, xt);
  return xt;
}

void mlp_fwd(Tensor feat, Tensor s_ray, Tensor pe, int P, Tensor counters, Tensor params, int width, Tensor rgb,
             c10::optional<Tensor> wpack) {
  chkf(params, "params");
  const int C = feat.size(1), pe_stride = pe.size(1);
  Tensor xt = tiles_from_streams(feat, s_ray, pe, P, counters);
  Tensor wp = ensure_pack(wpack, params, C, P, pe_stride, width);
  mlp_fwd_tiles(xt, C, P, pe_stride, counters, s_ray.numel(), wp, rgb, c10::nullopt);
}

Tensor mlp_fwd_timeline(Tensor feat, Tensor s_ray, Tensor pe, int P, Tensor counters, Tensor params, int width, Tensor rgb) {
  const int C = feat.size(1), pe_stride = pe.size(1);
  auto tl = torch::zeros({128}, feat.options().dtype(torch::kInt64));
  Tensor xt = tiles_from_streams(feat, s_ray, pe, P, counters);
  Tensor wp = ensure_pack(c10::nullopt, params, C, P, pe_stride, width);
  mlp_fwd_tiles(xt, C, P, pe_stride, counters, s_ray.numel(), wp, rgb, tl);
  return tl;
}

void mlp_bwd(Tensor feat, Tensor s_ray, Tensor pe, int P, Tensor counters, Tensor params, int width, Tensor rgb, Tensor d_rgb,
             double grad_scale, Tensor d_feat, Tensor grads, c10::optional<Tensor> wpack, c10::optional<Tensor> timeline) {
  chkf(params, "params");
  const int C = feat.size(1), pe_stride = pe.size(1);
  const int64_t cap = s_ray.numel();
  TORCH_CHECK(rgb.numel() == cap * 3, "rgb must be [surv_cap,3]");
  Tensor xt = tiles_from_streams(feat, s_ray, pe, P, counters);
  auto dzt = torch::zeros({dvgo_mlp_dztile_bytes(cap)}, feat.options().dtype(torch::kUInt8));
  mlp_pack_dz(rgb, d_rgb, grad_scale, counters, dzt);
  Tensor wp = ensure_pack(wpack, params, C, P, pe_stride, width);
  mlp_bwd_tiles(xt, dzt, C, P, pe_stride, counters, cap, wp, grad_scale, d_feat, grads, timeline);
}

Tensor mlp_bwd_timeline(Tensor feat, Tensor s_ray, Tensor pe, int P, Tensor counters, Tensor params, int width, Tensor rgb,
                        Tensor d_rgb, double grad_scale, Tensor d_feat, Tensor grads) {
  auto tl = torch::zeros({128}, feat.options().dtype(torch::kInt64));
  mlp_bwd(feat, s_ray, pe, P, counters, params, width, rgb, d_rgb, grad_scale, d_feat, grads, c10::nullopt, tl);
  return tl;
}
}  // namespace

void dvgo_bind_mlp(pybind11::module_& m) {
  m.def("tc_selftest", &tc_selftest);
  m.def("tc_probe", &tc_probe);
  m.def("tc_rate", &tc_rate);
  m.def("tc_ts_probe", &tc_ts_probe);
  m.def("tc_ldtm_rate", &tc_ldtm_rate);
  m.def("tc_contention", &tc_contention);
  m.def("view_embedding", &view_embedding, pybind11::arg("viewdirs"), pybind11::arg("freq"), pybind11::arg("stride"),
        pybind11::arg("C") = -1);
  m.def("mlp_wpack_bytes", &mlp_wpack_bytes);
  m.def("mlp_pack", &mlp_pack);
  m.def("mlp_fwd", &mlp_fwd, pybind11::arg("feat"), pybind11::arg("s_ray"), pybind11::arg("pe"), pybind11::arg("P"),
        pybind11::arg("counters"), pybind11::arg("params"), pybind11::arg("width"), pybind11::arg("rgb"),
        pybind11::arg("wpack") = pybind11::none());
  m.def("mlp_fwd_timeline", &mlp_fwd_timeline);
  m.def("mlp_bwd", &mlp_bwd, pybind11::arg("feat"), pybind11::arg("s_ray"), pybind11::arg("pe"), pybind11::arg("P"),
        pybind11::arg("counters"), pybind11::arg("params"), pybind11::arg("width"), pybind11::arg("rgb"),
        pybind11::arg("d_rgb"), pybind11::arg("grad_scale"), pybind11::arg("d_feat"), pybind11::arg("grads"),
        pybind11::arg("wpack") = pybind11::none(), pybind11::arg("timeline") = pybind11::none());
  m.def("mlp_xtile_bytes", &mlp_xtile_bytes);
  m.def("mlp_dztile_bytes", &mlp_dztile_bytes);
  m.def("mlp_pack_x", &mlp_pack_x);
  m.def("mlp_pack_dz", &mlp_pack_dz);
  m.def("mlp_fwd_tiles", &mlp_fwd_tiles, pybind11::arg("xt"), pybind11::arg("C"), pybind11::arg("P"),
        pybind11::arg("pe_stride"), pybind11::arg("counters"), pybind11::arg("cap"), pybind11::arg("wpack"),
        pybind11::arg("rgb"), pybind11::arg("timeline") = pybind11::none());
  m.def("mlp_bwd_tiles", &mlp_bwd_tiles, pybind11::arg("xt"), pybind11::arg("dzt"), pybind11::arg("C"), pybind11::arg("P"),
        pybind11::arg("pe_stride"), pybind11::arg("counters"), pybind11::arg("cap"), pybind11::arg("wpack"),
        pybind11::arg("grad_scale"), pybind11::arg("d_feat"), pybind11::arg("grads"),
        pybind11::arg("timeline") = pybind11::none());
  m.def("mlp_bwd_timeline", &mlp_bwd_timeline);
}
