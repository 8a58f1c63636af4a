// fishgym_cuda.cu — the product library: include/fishgym.h over CUDA for sm_100a.
// Build: see Makefile (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo).  CUDA runtime only.
#include "lbm_core.cuh"
#include "dev_cuda.cuh"
#define FG_DEV fg::CudaDev
#if defined(FG_POP16)
#define FG_BACKEND_NAME "cuda-sm100a-f16"    // 16-bit population storage, fp32 arithmetic (opt-in build)
#else
#define FG_BACKEND_NAME "cuda-sm100a"
#endif
#include "abi_impl.hpp"
