// k_generic.cu — instantiations of the general interpreter kernel fx_interp_kernel<K, SKIP, EXT>.
#include "fx8010_families.h"

namespace fxk {

template <int K> static KernelFn pick(bool skip, bool ext) {
    if (skip) return ext ? fx_interp_kernel<K, true, true> : fx_interp_kernel<K, true, false>;
    return ext ? fx_interp_kernel<K, false, true> : fx_interp_kernel<K, false, false>;
}
KernelFn generic_kernel(int K, bool skip, bool ext) {
    return K == 4 ? pick<4>(skip, ext) : (K == 2 ? pick<2>(skip, ext) : pick<1>(skip, ext));
}
cudaError_t upload_generic(const uint4* src, size_t bytes, int word_off, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_prog, src, bytes, sizeof(uint4) * (size_t)word_off, cudaMemcpyHostToDevice, st);
}

}  // namespace fxk
