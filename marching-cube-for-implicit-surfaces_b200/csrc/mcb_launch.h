/* How the host code launches a kernel and how a kernel names its dynamic shared memory.
 * The product build (nvcc) uses the CUDA forms below.  A test harness may define both before this header is seen
 * (tests/emu/cuda_runtime.h executes the kernels' source on the host for the CPU test tier); nothing in the product
 * build does. */
#ifndef MCB_LAUNCH
#include <cstdio>
#include <cstdlib>
/* $MCB_DEBUG_SYNC=1 (debugging aid, read once): wait for every kernel right after its launch and name the one that
 * failed on stderr — an asynchronous fault is otherwise reported by whatever call synchronises next. */
inline bool mcb_debug_sync_on() {
    static const bool on = [] { const char* e = std::getenv("MCB_DEBUG_SYNC"); return e && e[0] == '1'; }();
    return on;
}
inline void mcb_debug_sync_check(const char* kernel, cudaStream_t stream) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) std::fprintf(stderr, "MCB_DEBUG_SYNC: kernel %s: %s\n", kernel, cudaGetErrorString(e));
}
#define MCB_UNPAREN(...) __VA_ARGS__
#define MCB_LAUNCH(kern, grid, block, smem, stream, ...)                                  \
    do {                                                                                  \
        MCB_UNPAREN kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);             \
        if (mcb_debug_sync_on()) mcb_debug_sync_check(#kern, (stream));                   \
    } while (0)
#define MCB_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#endif
