// TEST INFRASTRUCTURE ONLY (oracle/): force-included (-include) when the recipe
// oracle/build_ref.py compiles the reference's OWN models/csrc sources *where
// they lie* under /root/reference, unmodified, into oracle/_ref/vren_ref*.so.
//
// The reference targets a 2022 PyTorch; with torch 2.11 its sources fail on two
// mechanical points:
//   (1) AT_DISPATCH_*(<tensor>.type(), ...) - the macros now want a ScalarType,
//       .type() returns at::DeprecatedTypeProperties (16 sites);
//   (2) thrust::device / thrust::reduce used without their headers.
// Instead of editing a copy of the sources we re-define the three dispatch
// macros the reference uses so that they accept either type, and pre-include the
// thrust headers.  No kernel code is touched, so the compiled kernels are the
// reference's arithmetic bit for bit.
#pragma once
#include <torch/extension.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>

namespace ncn_ref_compat {
inline c10::ScalarType st(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
inline c10::ScalarType st(c10::ScalarType t) { return t; }
}  // namespace ncn_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(ncn_ref_compat::st(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
#undef AT_DISPATCH_FLOATING_TYPES_AND_HALF
#define AT_DISPATCH_FLOATING_TYPES_AND_HALF(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(ncn_ref_compat::st(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES_AND_HALF(__VA_ARGS__))
#undef AT_DISPATCH_INTEGRAL_TYPES
#define AT_DISPATCH_INTEGRAL_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(ncn_ref_compat::st(TYPE), NAME, AT_DISPATCH_CASE_INTEGRAL_TYPES(__VA_ARGS__))
