// ORACLE / TEST INFRASTRUCTURE.  The reference's curope.cpp declares rope_2d_cuda (defined in kernels.cu, which does not compile against
// this image's torch: SURVEY.md section 8c) and dispatches to it for CUDA tensors.  oracle/_ref only needs the CPU branch (rope_2d_cpu,
// curope.cpp:11-47), so the CUDA symbol is satisfied by a stub that refuses to run.
#include <torch/extension.h>

void rope_2d_cuda(torch::Tensor, const torch::Tensor, const float, const float) {
    TORCH_CHECK(false, "oracle/_ref/curope_ref: CPU build of the reference's curope.cpp; the CUDA branch is not part of it");
}
