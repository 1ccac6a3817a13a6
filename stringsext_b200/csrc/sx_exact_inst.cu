// sx_exact_inst.cu -- one instantiation of sx_exact_kernel<Dec> per translation unit (-DSX_INST=n).
#include "sx_exact.cuh"
#if SX_INST == 0 || SX_INST == 1 || SX_INST == 4
#include "sx_sparse_utf8.cuh"
#define SX_HAS_SPARSE 1
#endif

namespace sx {
#if SX_INST == 0
#define SX_DEC DecXud
#define SX_NAME launch_exact_xud
#elif SX_INST == 1
#define SX_DEC DecUtf8
#define SX_NAME launch_exact_utf8
#elif SX_INST == 2
#define SX_DEC DecUtf16<false>
#define SX_NAME launch_exact_utf16le
#elif SX_INST == 3
#define SX_DEC DecUtf16<true>
#define SX_NAME launch_exact_utf16be
#elif SX_INST == 4
#define SX_DEC DecSb
#define SX_NAME launch_exact_sb
#elif SX_INST == 5
#define SX_DEC DecUtf32<false>
#define SX_NAME launch_exact_utf32le
#elif SX_INST == 6
#define SX_DEC DecUtf32<true>
#define SX_NAME launch_exact_utf32be
#else
#error "SX_INST must be 0..6"
#endif

cudaError_t SX_NAME(const ScanParams& P, const ScanOut& O, const ExactCfg& X, unsigned grid, cudaStream_t st) {
    sx_exact_kernel<SX_DEC><<<grid, kThreads, 0, st>>>(P, O, X);
    return cudaGetLastError();
}
#if defined(SX_HAS_SPARSE)
#if SX_INST == 0
#define SX_SPARSE_NAME launch_sparse_xud
#elif SX_INST == 1
#define SX_SPARSE_NAME launch_sparse_utf8
#else
#define SX_SPARSE_NAME launch_sparse_sb
#endif
cudaError_t SX_SPARSE_NAME(const ScanParams& P, const ScanOut& O, const ExactCfg& X, void* entries, void* btot, void* tables,
                           void* queue, long long NE, int num_sms, cudaStream_t st, cudaEvent_t* ev, cudaStream_t side, cudaEvent_t* evs) {
    SparseBufs B;
    B.E = static_cast<EntryState*>(entries);
    B.btot = static_cast<ulonglong2*>(btot);
    B.tables = static_cast<Utf8Tables*>(tables);
    B.queue = static_cast<uint32_t*>(queue);
    B.qcount = O.counters + 3;  // the block kernel's claim counter, unused on this path (zeroed per attempt)
    B.queue2 = B.queue + NE + 32;
    B.qcount2 = O.counters + 4;  // [5]: snapshot of [4] for sx_sp_declined_kernel
    B.NE = NE;
    return launch_sparse_impl<SX_DEC>(P, O, X, B, num_sms, st, ev, side, evs);
}
#endif
#if SX_INST == 1
size_t sparse_entry_bytes() { return sizeof(EntryState); }
size_t sparse_tables_bytes() { return sizeof(Utf8Tables); }
uint32_t sparse_threads() { return kSpThreads; }
uint32_t sparse_launches() { return kSparseLaunches; }
#endif
}  // namespace sx
