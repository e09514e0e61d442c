// fused_fast.cu -- register-resident specialisation of the fused hot path
// (size 256 / 512).  Placeholder until the 16x16 register FFT lands: reports
// "not handled" so the generic shared-memory kernel runs.
#include "common.cuh"
#include "fused.cuh"

namespace sep {

int fused_fast_try(const sep_plan *, const FusedArgs &, int, int, double *, double *, Scratch &,
                   cudaStream_t, bool *handled) {
  *handled = false;
  return SEP_OK;
}

}  // namespace sep
