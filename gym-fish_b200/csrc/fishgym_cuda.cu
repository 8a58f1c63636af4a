// fishgym_cuda.cu — the product library: include/fishgym.h over CUDA for sm_100a.
// Build: see Makefile (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo).  CUDA runtime only.
#include "lbm_core.cuh"
#include "dev_cuda.cuh"
#define FG_DEV fg::CudaDev
#define FG_BACKEND_NAME "cuda-sm100a"
#include "abi_impl.hpp"
