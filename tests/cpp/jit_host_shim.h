/* TEST INFRASTRUCTURE ONLY.  Lets g++ compile the CUDA source that mcb_jit.cpp generates, so that the generated
 * kernels can be executed on the host in the CPU test tier (tests/test_host_logic.py).  Lanes execute one after the
 * other; the only warp-wide operation, __ballot_sync, is emulated by running every warp twice: the first pass records
 * each lane's predicate per ballot call (control flow is uniform across a warp, so the calls line up), the second
 * returns the assembled masks.  Build with -ffp-contract=off. */
#include <cmath>
#include <cstdint>
#include <cstring>
#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
struct dim3s { unsigned x, y, z; };
static dim3s threadIdx, blockIdx, gridDim, blockDim;
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline void __stcs(float* p, float v) { *p = v; }
static inline int min(int a, int b) { return a < b ? a : b; }
static int g_pass = 0, g_lane = 0, g_seq = 0;
static unsigned g_masks[1024];
static inline unsigned __ballot_sync(unsigned, bool p) {
    const int s = g_seq++;
    if (g_pass == 0) { if (p) g_masks[s] |= 1u << g_lane; return 0u; }
    return g_masks[s];
}
