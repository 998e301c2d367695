// k_short.cu — instantiations of the short-program kernel fx_short_kernel<K, EXT, NI>.
#include "fx8010_families.h"
#include "fx8010_short.cuh"

namespace fxk {

static_assert(SH_MAX_NI == SH_MAX_NI_HOST, "keep fx8010_families.h in step");

template <int K, bool EXT> static KernelFn pick(int ni) {
    switch (ni) {
    case 1: return fx_short_kernel<K, EXT, 1>;
    case 2: return fx_short_kernel<K, EXT, 2>;
    case 3: return fx_short_kernel<K, EXT, 3>;
    default: return fx_short_kernel<K, EXT, 4>;
    }
}
template <int K> static KernelFn pick(bool ext, int ni) { return ext ? pick<K, true>(ni) : pick<K, false>(ni); }
KernelFn short_kernel(int K, bool ext, int ni) {
    return K == 4 ? pick<4>(ext, ni) : (K == 2 ? pick<2>(ext, ni) : pick<1>(ext, ni));
}
cudaError_t upload_short(const uint4* src, size_t bytes, int word_off, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_prog, src, bytes, sizeof(uint4) * (size_t)word_off, cudaMemcpyHostToDevice, st);
}

}  // namespace fxk
