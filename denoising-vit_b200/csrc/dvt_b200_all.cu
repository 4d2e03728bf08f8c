// Unity translation unit: one device link unit (no -rdc), one shared object.
#include "runtime.cu"
#include "gemm.cu"
#include "api.cu"
