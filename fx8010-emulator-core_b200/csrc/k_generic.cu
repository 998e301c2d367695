// k_generic.cu — instantiations of the general interpreter kernel fx_interp_kernel<K, SKIP, EXT, NI>.
#include "fx8010_families.h"

namespace fxk {

template <int K, int NI> static KernelFn pick(bool skip, bool ext) {
    if (skip) return ext ? fx_interp_kernel<K, true, true, NI> : fx_interp_kernel<K, true, false, NI>;
    return ext ? fx_interp_kernel<K, false, true, NI> : fx_interp_kernel<K, false, false, NI>;
}
template <int K> static KernelFn pick(bool skip, bool ext, bool shortp) {
    (void)shortp;                        // short programs the other kernels cannot take run the fetch loop like any other
    return pick<K, 0>(skip, ext);
}
KernelFn generic_kernel(int K, bool skip, bool ext, bool shortp) {
    return K == 4 ? pick<4>(skip, ext, shortp) : (K == 2 ? pick<2>(skip, ext, shortp) : pick<1>(skip, ext, shortp));
}
cudaError_t upload_generic(const uint4* src, size_t bytes, int slot, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_prog, src, bytes, sizeof(uint4) * (size_t)SLOT_WORDS * slot, cudaMemcpyHostToDevice, st);
}

}  // namespace fxk
