// fishgym_emu.cpp — TEST INFRASTRUCTURE: the product's kernel bodies and host runtime compiled with g++
// and executed on the CPU (see dev_host.hpp).  Exports the ABI of include/fishgym.h so the tests can
// compare it with the fp64 oracle without a GPU.  Not a product backend, not a fallback.
#include "../../gym-fish_b200/csrc/lbm_core.cuh"
#include "dev_host.hpp"
#define FG_DEV fg::HostDev
#if defined(FG_POP16)
#define FG_BACKEND_NAME "emu-host-f16"
#else
#define FG_BACKEND_NAME "emu-host-fp32"
#endif
#include "../../gym-fish_b200/csrc/abi_impl.hpp"
