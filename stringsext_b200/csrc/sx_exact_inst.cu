// sx_exact_inst.cu -- the kernels of ONE decoder per translation unit (-DSX_INST=<ENC_* value>), so the library builds
// in parallel: sx_exact_kernel<Dec>, sx_range_carry_kernel<Dec> and, for the decoders with a mask engine, the
// sparse-list pipeline (sx_sparse_utf8.cuh).
#include "sx_exact.cuh"
#include "sx_sparse_utf8.cuh"

namespace sx {
#if SX_INST == 0
#define SX_DEC DecXud
#elif SX_INST == 1
#define SX_DEC DecUtf8
#elif SX_INST == 2
#define SX_DEC DecUtf16<false>
#elif SX_INST == 3
#define SX_DEC DecUtf16<true>
#elif SX_INST == 4
#define SX_DEC DecSb
#elif SX_INST == 5
#define SX_DEC DecUtf32<false>
#elif SX_INST == 6
#define SX_DEC DecUtf32<true>
#elif SX_INST == 7
#define SX_DEC DecBig5
#elif SX_INST == 8
#define SX_DEC DecEucJp
#else
#error "SX_INST must be 0..8"
#endif
#define SX_CAT2(a, b) a##b
#define SX_CAT(a, b) SX_CAT2(a, b)

cudaError_t SX_CAT(launch_exact_, SX_INST)(const ScanParams& P, const ScanOut& O, const ExactCfg& X, unsigned grid, cudaStream_t st) {
    sx_exact_kernel<SX_DEC><<<grid, kThreads, 0, st>>>(P, O, X);
    return cudaGetLastError();
}
cudaError_t SX_CAT(launch_range_carry_, SX_INST)(const ScanParams& P, const RangeCarryArgs& A, cudaStream_t st) {
    sx_range_carry_kernel<SX_DEC><<<A.nranges, 32, 0, st>>>(P, A);
    return cudaGetLastError();
}
bool SX_CAT(has_sparse_, SX_INST)() { return true; }
template <class Dec, bool kHas> struct SparseLauncher {
    static cudaError_t go(const ScanParams&, const ScanOut&, const ExactCfg&, const SparseBufs&, const SparseLaunchCfg&, cudaStream_t,
                          cudaEvent_t*, cudaStream_t, cudaEvent_t*) { return cudaErrorNotSupported; }
};
template <class Dec> struct SparseLauncher<Dec, true> {
    static cudaError_t go(const ScanParams& P, const ScanOut& O, const ExactCfg& X, const SparseBufs& B, const SparseLaunchCfg& L,
                          cudaStream_t st, cudaEvent_t* ev, cudaStream_t side, cudaEvent_t* evs) {
        return launch_sparse_impl<Dec>(P, O, X, B, L, st, ev, side, evs);
    }
};
cudaError_t SX_CAT(launch_sparse_, SX_INST)(const ScanParams& P, const ScanOut& O, const ExactCfg& X, const SparseBufs& B,
                                            const SparseLaunchCfg& L, cudaStream_t st, cudaEvent_t* ev, cudaStream_t side, cudaEvent_t* evs) {
    return SparseLauncher<SX_DEC, true>::go(P, O, X, B, L, st, ev, side, evs);
}
}  // namespace sx
