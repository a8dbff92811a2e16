// Axis-pass launchers and kernels for 4-byte labels (see edt_passes.cuh).
#include "edt_passes.cuh"
EDT_INSTANTIATE_PASSES(4)
