/* How the host code launches a kernel and how a kernel names its dynamic shared memory.
 * The product build (nvcc) uses the CUDA forms below.  A test harness may define both before this header is seen
 * (tests/emu/cuda_runtime.h executes the kernels' source on the host for the CPU test tier); nothing in the product
 * build does. */
#ifndef MCB_LAUNCH
#define MCB_UNPAREN(...) __VA_ARGS__
#define MCB_LAUNCH(kern, grid, block, smem, stream, ...) MCB_UNPAREN kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MCB_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#endif
