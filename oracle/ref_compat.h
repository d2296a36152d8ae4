// TEST INFRASTRUCTURE ONLY (oracle/_ref build recipe) -- never part of the product path.
//
// Force-included (-include) in front of the UNMODIFIED reference sources
// /root/reference/lib/cuda/*.cu so that they compile against torch 2.11:
// the reference writes AT_DISPATCH_FLOATING_TYPES(x.type(), ...) (e.g.
// lib/cuda/render_utils_kernel.cu:86), and torch >= 2.x no longer converts
// at::DeprecatedTypeProperties to c10::ScalarType implicitly.  We pull in
// torch/extension.h first (its include guard turns the reference's own
// #include into a no-op) and re-define the dispatch macro so that both
// spellings are accepted.  No reference source is copied or edited.
#pragma once
#include <torch/extension.h>

namespace dvgo_ref_compat {
inline c10::ScalarType to_scalar(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
inline c10::ScalarType to_scalar(c10::ScalarType t) { return t; }
}  // namespace dvgo_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(::dvgo_ref_compat::to_scalar(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
