// Axis-pass launchers and kernels for 1-byte labels (see edt_passes.cuh).
#include "edt_passes.cuh"
EDT_INSTANTIATE_PASSES(1)
