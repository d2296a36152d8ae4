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
}  // namespace

void dvgo_bind_mlp(pybind11::module_& m) {
  m.def("tc_selftest", &tc_selftest);
  m.def("tc_probe", &tc_probe);
}
