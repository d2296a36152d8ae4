// mlp_binding.cpp -- torch adaptor for the tensor-core rgbnet kernels (fused_mlp.cu).
#include <torch/extension.h>

void dvgo_bind_mlp(pybind11::module_& m) { (void)m; }
