// pgtg_inst_philox.cu -- the PGTG_RNG_PHILOX instantiations of the tick and map-generation kernels
// (pgtg_tick_kernels.cuh); one translation unit per random-number source so that they build in parallel.
#include "pgtg_tick_kernels.cuh"

int pgtg_launch_mode_philox(pgtg_env* e, int mode, const uint8_t* mask, const int64_t* seeds, const void* actions, int action_bytes, void* stream) {
  return launch_mode<PGTG_RNG_PHILOX>(e, mode, mask, seeds, actions, action_bytes, (cudaStream_t)stream);
}
