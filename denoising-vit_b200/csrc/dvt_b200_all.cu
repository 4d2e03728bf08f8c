// Unity translation unit: one device link unit (no -rdc), one shared object.
#include "runtime.cu"
#include "gemm.cu"
#include "vit_kernels.cu"
#include "attention.cu"
#include "vit.cu"
#include "fit.cu"
#include "views.cu"
#include "api.cu"
