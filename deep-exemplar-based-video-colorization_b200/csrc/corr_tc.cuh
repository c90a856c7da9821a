// tcgen05 (5th-gen tensor core) implementation of K7; see corr_tc.cu.
#pragma once
#include <string>

#include "dvc_internal.cuh"

namespace dvc {
// math = DVC_MATH_TF32X3 or DVC_MATH_BF16X3.  Returns 0 on success, non-zero with *err set otherwise.
// cluster: 2 = CTA pairs (tcgen05.mma.cta_group::2) on adjacent query-row tiles, 1 = single CTAs
int launch_corr_tc(const CorrParams& p, int math, int cluster, cudaStream_t s, std::string* err);
}  // namespace dvc
