// fx8010_families.h — the kernel families, one translation unit each (they compile in parallel).
//
// Every family keeps its own copy of the decoded program in __constant__ memory (c_prog is `static` in
// fx8010_kernel.cuh), so the host uploads an encoding to the family whose kernel it is about to launch.
#pragma once

#include <cuda_runtime.h>

#include "fx8010_kernel.cuh"
#include "fx8010_stateless.cuh"

namespace fxk {

typedef void (*KernelFn)(const Params);
typedef void (*SLKernelFn)(const SLParams);

enum Family { FAM_GENERIC = 0, FAM_SHORT, FAM_SL1, FAM_SL2, FAM_SL4, FAM_COUNT };

// fx_interp_kernel<K, SKIP, EXT>  (k_generic.cu)
KernelFn generic_kernel(int K, bool skip, bool ext);
// fx_short_kernel<K, EXT, NI>  (k_short.cu); ni = exact number of encoded instructions, 1..SH_MAX_NI
KernelFn short_kernel(int K, bool ext, int ni);
// fx_stateless_kernel<K, TRAM>  (k_sl1.cu, k_sl2.cu, k_sl4.cu)
SLKernelFn sl_kernel(int K, bool tram);
// copies `bytes` of an encoding to word `word_off` of the family's constant-memory arena
cudaError_t upload_program(Family f, const uint4* src, size_t bytes, int word_off, cudaStream_t st);
inline Family sl_family(int K) { return K == 4 ? FAM_SL4 : (K == 2 ? FAM_SL2 : FAM_SL1); }

constexpr int SH_MAX_NI_HOST = 4;   // == SH_MAX_NI of fx8010_short.cuh (static_assert there)

// per-family pieces (defined in the family's translation unit)
cudaError_t upload_generic(const uint4*, size_t, int, cudaStream_t);
cudaError_t upload_short(const uint4*, size_t, int, cudaStream_t);
cudaError_t upload_sl1(const uint4*, size_t, int, cudaStream_t);
cudaError_t upload_sl2(const uint4*, size_t, int, cudaStream_t);
cudaError_t upload_sl4(const uint4*, size_t, int, cudaStream_t);
SLKernelFn sl_kernel_1(bool tram);
SLKernelFn sl_kernel_2(bool tram);
SLKernelFn sl_kernel_4(bool tram);

}  // namespace fxk
