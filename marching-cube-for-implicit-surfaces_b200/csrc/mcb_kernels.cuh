/* sm_100a kernels of the polygoniser.  Included once by mcb_api.cu.
 *
 * Device data layout (one z-slab of cube layers [kb,ke), M cubes per axis):
 *   vertices carry a one-vertex apron on every side (needed by the central-difference normals and, in z, the slab
 *   halo):  NV = M+3 vertex columns per axis, vertex v in [-1, M+1] stored at index v+1;
 *   coords   cs[P]              cs[v+1] = coordinate of vertex v (host-accumulated exactly like marching.cpp:375-377)
 *   field    F[NZ][NV][P]       fp32, x fastest, row pitch P = NV rounded up to 32, NZ = (ke-kb)+3 planes
 *   signs    S[NZ][NV][WP]      1 bit per vertex: F > iso (strict, NaN -> 0; marching.cpp:497-505), WP = P/32
 *   valid    V[NZ][NV][WP]      1 bit per vertex: all constraints in use hold (marching.cpp:255-280); optional
 *   tables   T[axis][slot][P]   hoisted single-variable subtrees per grid coordinate
 *   records  R[capA] (u64)      i | j<<12 | k<<24 | code<<36 | table_idx<<44, in cube loop order
 *   trioff   O[capA] (u32)      index of the cube's first triangle
 *   pos,nrm  float4[3*capT]     triangle soup in the reference's emission order
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mcb_bytecode.h"
#include "mcb_tables.h"

namespace mcbk {

constexpr int kEvalThreads = 128;   /* 4 warps; each warp: 32 x-columns x kEvalRows y-rows of one z-plane */
constexpr int kEvalRows = 8;
constexpr int kClsThreads = 256;
constexpr int kClsItems = 8;        /* consecutive 32-cube words per thread in classify */
constexpr int kEmitThreads = 128;   /* = active cubes per emit chunk */
constexpr uint32_t kSpinLimit = 1u << 26;

struct Grid {
    int M;        /* cubes per axis */
    int NV;       /* M + 3 */
    int P;        /* row pitch of F (floats) */
    int WP;       /* row pitch of S/V (words) */
    int kb, ke;   /* cube layers of the slab */
    int NZ;       /* (ke-kb) + 3 vertex planes */
    float sx, sy, sz;
    float iso;
};

struct Counters {
    unsigned long long active;
    unsigned long long triangles;
    unsigned long long ambiguous;
    unsigned long long redirected;
    unsigned int tile_ticket;
    unsigned int error; /* 1 = look-back spin limit hit */
};

/* ---------------------------------------------------------------------------------------------------------------
 * K0a  fold_constants: evaluate every maximal constant subtree once (same interpreter, same arithmetic as the
 *      per-point evaluation, so the folded value is bit-identical to what the reference recomputes per call).
 * K0b  axis_tables: evaluate every hoisted single-variable subtree at each grid coordinate of its axis.
 * ------------------------------------------------------------------------------------------------------------- */
struct SlotDesc {
    int begin, len; /* program inside slot_code */
    int axis;       /* -1 constant, else 0..2 */
    int index;      /* constant-pool index / table index within the axis */
};

__global__ void fold_constants_kernel(const uint32_t* __restrict__ slot_code, const SlotDesc* __restrict__ slots,
                                      int nslots, const float* kpool, float* out /* may alias kpool */) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots || slots[s].axis >= 0) return;
    out[slots[s].index] = mcb_interp_scalar(slot_code + slots[s].begin, slots[s].len, kpool, 0.f, 0.f, 0.f,
                                            nullptr, nullptr, nullptr);
}

__global__ void axis_tables_kernel(const uint32_t* __restrict__ slot_code, const SlotDesc* __restrict__ slots,
                                   int first_axis_slot, const float* __restrict__ kpool,
                                   const float* __restrict__ cs, int NV, int P, int max_slots_per_axis,
                                   float sx, float sy, float sz, float* __restrict__ tables) {
    int s = first_axis_slot + blockIdx.y;
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P) return;
    SlotDesc d = slots[s];
    float scale = d.axis == 0 ? sx : d.axis == 1 ? sy : sz;
    float c = scale * cs[v < NV ? v : NV - 1]; /* Marching::evaluate scales first (marching.cpp:211) */
    float r = mcb_interp_scalar(slot_code + d.begin, d.len, kpool, c, c, c, nullptr, nullptr, nullptr);
    tables[((size_t)d.axis * max_slots_per_axis + d.index) * P + v] = r;
}

/* ---------------------------------------------------------------------------------------------------------------
 * K1  eval_field: the bytecode interpreter over the grid.
 *     A warp owns 32 consecutive x-columns x kEvalRows consecutive y-rows of one z-plane; a lane keeps the 4 row
 *     values of the top of the operand stack in registers and the deeper levels in shared memory
 *     ([level][row][thread], conflict-free).  Every lane executes the same instruction word, fetched from the
 *     kernel-parameter block (constant bank).  Outputs: F (coalesced 128 B per warp store) and the sign bit-plane
 *     S (one ballot per row) — the comparison against iso is fused here so that classification never has to read
 *     the 4 B/vertex field again.
 * ------------------------------------------------------------------------------------------------------------- */
__device__ __noinline__ float powf_call(float a, float b) { return mcb_powf(a, b); } /* one copy, register args */

__global__ void __launch_bounds__(kEvalThreads)
eval_field_kernel(const __grid_constant__ mcb_program prog, const Grid g, const float* __restrict__ cs,
                  const float* __restrict__ tables, int max_slots_per_axis, float* __restrict__ F,
                  uint32_t* __restrict__ S, unsigned items_per_plane) {
    extern __shared__ float stack_smem[]; /* [depth][kEvalRows][kEvalThreads] */
    const int lane = threadIdx.x & 31;
    const unsigned item = blockIdx.x * (kEvalThreads / 32) + (threadIdx.x >> 5); /* (row group, word) of plane blockIdx.y */
    if (item >= items_per_plane) return; /* whole warp exits together */
    const int w = (int)(item % (unsigned)g.WP);
    const int yq = (int)(item / (unsigned)g.WP);
    const int pz = (int)blockIdx.y;
    const int x = w * 32 + lane;
    const int y0 = yq * kEvalRows;

    const float X = g.sx * cs[x < g.NV ? x : g.NV - 1];
    float Y[kEvalRows];
#pragma unroll
    for (int e = 0; e < kEvalRows; e++) {
        int y = y0 + e < g.NV ? y0 + e : g.NV - 1;
        Y[e] = g.sy * cs[y];
    }
    const float Z = g.sz * cs[pz + (g.kb)]; /* plane pz holds vertex kb-1+pz -> cs index kb+pz */
    const float* tabx = tables + (size_t)0 * max_slots_per_axis * g.P;
    const float* taby = tables + (size_t)1 * max_slots_per_axis * g.P;
    const float* tabz = tables + (size_t)2 * max_slots_per_axis * g.P;
    const int zi = pz + g.kb;

    float t0[kEvalRows];
#pragma unroll
    for (int e = 0; e < kEvalRows; e++) t0[e] = 0.f;
    float* sp = stack_smem + threadIdx.x; /* next free level */
    constexpr int kLevel = kEvalRows * kEvalThreads;

#pragma unroll 1
    for (int pc = 0; pc < prog.n; pc++) {
        const uint32_t insn = prog.code[pc];
        const uint32_t op = MCB_INSN_OP(insn), arg = MCB_INSN_ARG(insn);
        if (op <= MCB_OP_PUSH_TZ) { /* pushes: spill the cached top, load the new one */
#pragma unroll
            for (int e = 0; e < kEvalRows; e++) sp[e * kEvalThreads] = t0[e];
            sp += kLevel;
            switch (op) {
                case MCB_OP_PUSH_X:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = X;
                    break;
                case MCB_OP_PUSH_Y:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = Y[e];
                    break;
                case MCB_OP_PUSH_Z:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = Z;
                    break;
                case MCB_OP_PUSH_K: {
                    float kv = prog.k[arg];
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = kv;
                    break;
                }
                case MCB_OP_PUSH_TX: {
                    float tv = __ldg(tabx + (size_t)arg * g.P + x);
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = tv;
                    break;
                }
                case MCB_OP_PUSH_TY: {
                    const float* ty = taby + (size_t)arg * g.P + y0;
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = __ldg(ty + e); /* tables are padded past NV+kEvalRows */
                    break;
                }
                default: { /* MCB_OP_PUSH_TZ (MCB_OP_END never appears inside a program) */
                    float tv = __ldg(tabz + (size_t)arg * g.P + zi);
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = tv;
                    break;
                }
            }
        } else if (op == MCB_OP_NEG) {
#pragma unroll
            for (int e = 0; e < kEvalRows; e++) t0[e] = -t0[e];
        } else {
            sp -= kLevel;
            float s2[kEvalRows];
#pragma unroll
            for (int e = 0; e < kEvalRows; e++) s2[e] = sp[e * kEvalThreads];
            switch (op) {
                case MCB_OP_ADD:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = s2[e] + t0[e];
                    break;
                case MCB_OP_SUB:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = s2[e] - t0[e];
                    break;
                case MCB_OP_RSUB:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = t0[e] - s2[e];
                    break;
                case MCB_OP_MUL:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = s2[e] * t0[e];
                    break;
                case MCB_OP_DIV:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = s2[e] / t0[e];
                    break;
                case MCB_OP_RDIV:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = t0[e] / s2[e];
                    break;
                case MCB_OP_POW:
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = powf_call(s2[e], t0[e]);
                    break;
                default: /* MCB_OP_RPOW */
#pragma unroll
                    for (int e = 0; e < kEvalRows; e++) t0[e] = powf_call(t0[e], s2[e]);
                    break;
            }
        }
    }

    const size_t row0 = (size_t)pz * g.NV + y0;
    float* fp = F + row0 * g.P + x;
    uint32_t* sp_bits = S + row0 * g.WP + w;
#pragma unroll
    for (int e = 0; e < kEvalRows; e++) {
        const bool in = (y0 + e < g.NV); /* uniform per warp */
        if (in) fp[(size_t)e * g.P] = t0[e]; /* x < P always: P is the padded pitch */
        const unsigned bits = __ballot_sync(0xffffffffu, in && x < g.NV && t0[e] > g.iso);
        if (in && lane == e) sp_bits[e * g.WP] = bits;
    }
}

/* K1b  constraint validity bit-plane: V &= (lhs(sx*x,sy*y,sz*z) op rhs), one launch per constraint in use. */
__global__ void __launch_bounds__(256)
eval_constraint_kernel(const __grid_constant__ mcb_program prog, const Grid g, const float* __restrict__ cs,
                       int op, float rhs, int first, uint32_t* __restrict__ V, long long total_words) {
    const int lane = threadIdx.x & 31;
    const long long widx = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (widx >= total_words) return;
    const int w = (int)(widx % g.WP);
    const long long row = widx / g.WP;
    const int y = (int)(row % g.NV);
    const int pz = (int)(row / g.NV);
    const int x = w * 32 + lane;
    bool ok = false;
    if (x < g.NV) {
        float v = mcb_interp_scalar(prog.code, prog.n, prog.k, g.sx * cs[x], g.sy * cs[y], g.sz * cs[pz + g.kb],
                                    nullptr, nullptr, nullptr);
        ok = op == 0 ? v > rhs : op == 1 ? v < rhs : op == 2 ? v >= rhs : v <= rhs;
    }
    unsigned bits = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) V[widx] = first ? bits : (V[widx] & bits);
}

/* ---------------------------------------------------------------------------------------------------------------
 * K2  classify_compact: sign bit-planes -> cube codes -> ambiguity redirect -> per-cube triangle counts ->
 *     single-pass device-wide exclusive scan (decoupled look-back) of (active cubes, triangles) -> compacted,
 *     loop-ordered active-cube records with their triangle offsets.
 *
 *     Work item = one 32-cube word of a cube row; items are numbered in the reference's loop order
 *     (z slowest, then y, then x), a tile is kClsThreads*kClsItems consecutive items, and tiles take their
 *     number from an atomic ticket so that a tile only ever waits for tiles that already started.
 * ------------------------------------------------------------------------------------------------------------- */
struct ScanState {
    uint32_t* flag;       /* 0 = nothing, 1 = aggregate published, 2 = inclusive prefix published */
    uint32_t* agg_active;
    uint32_t* agg_tris;
    unsigned long long* inc_active;
    unsigned long long* inc_tris;
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *(const volatile unsigned long long*)p;
}

struct ClsTables { /* built once per context on the host (mcb_tables.h); read through the read-only path */
    uint64_t tri[256];
    int8_t face[256];
    uint8_t ntri[256];
};

/* the four sign-word rows a cube row touches: (y, z), (y+1, z), (y, z+1), (y+1, z+1) */
struct RowPtrs {
    const uint32_t* r[4];
};
__device__ __forceinline__ RowPtrs cube_row_ptrs(const uint32_t* __restrict__ B, const Grid& g, int kz, int j) {
    /* cube (i,j,kz) corner (dx,dy,dz) = vertex index (i+1+dx, j+1+dy, plane kz+1+dz) */
    RowPtrs p;
    p.r[0] = B + ((size_t)(kz + 1) * g.NV + (j + 1)) * g.WP;
    p.r[1] = p.r[0] + g.WP;
    p.r[2] = p.r[0] + (size_t)g.NV * g.WP;
    p.r[3] = p.r[2] + g.WP;
    return p;
}
/* corner sign words of the 32 cubes of word w: bit b of c[v] = sign of corner v of cube 32w+b.
 * lo[] = words w of the four rows (carried from the previous item when possible), hi[] = words w+1. */
__device__ __forceinline__ void corner_words(const uint32_t lo[4], const uint32_t hi[4], uint32_t c[8]) {
    c[0] = __funnelshift_r(lo[0], hi[0], 1); c[1] = __funnelshift_r(lo[0], hi[0], 2);
    c[3] = __funnelshift_r(lo[1], hi[1], 1); c[2] = __funnelshift_r(lo[1], hi[1], 2);
    c[4] = __funnelshift_r(lo[2], hi[2], 1); c[5] = __funnelshift_r(lo[2], hi[2], 2);
    c[7] = __funnelshift_r(lo[3], hi[3], 1); c[6] = __funnelshift_r(lo[3], hi[3], 2);
}

__device__ __forceinline__ int code_of(const uint32_t c[8], int b) {
    int code = 0;
#pragma unroll
    for (int v = 0; v < 8; v++) code |= ((c[v] >> b) & 1u) << v;
    return code;
}

/* Face-centre test of marching.cpp:527-547; returns true when the redirect row 255-code must be used. */
__device__ __noinline__ bool ambiguity_redirects(const mcb_program& prog, const Grid& g, const float* __restrict__ cs,
                                                 int face, int i, int j, int k) {
    float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int ofs = mcb_corner_ofs(mcb_face_corner(face, q));
        mx += cs[i + 1 + (ofs & 1)];
        my += cs[j + 1 + ((ofs >> 1) & 1)];
        mz += cs[k + 1 + ((ofs >> 2) & 1)];
    }
    mx /= 4.0f; my /= 4.0f; mz /= 4.0f;
    const float mid = mcb_interp_scalar(prog.code, prog.n, prog.k, g.sx * mx, g.sy * my, g.sz * mz, nullptr, nullptr, nullptr);
    return mid > g.iso;
}

__global__ void __launch_bounds__(kClsThreads)
classify_compact_kernel(const __grid_constant__ mcb_program point_prog, const Grid g, const float* __restrict__ cs,
                        const ClsTables* __restrict__ gtb, const uint32_t* __restrict__ S,
                        const uint32_t* __restrict__ V, int WC /* words per cube row */,
                        long long total_items, ScanState st, Counters* __restrict__ ctr,
                        unsigned long long* __restrict__ rec, uint32_t* __restrict__ trioff,
                        unsigned long long cap_active) {
    __shared__ uint32_t warp_a[kClsThreads / 32], warp_t[kClsThreads / 32];
    __shared__ unsigned long long base_a_s, base_t_s;
    __shared__ uint32_t tile_s;

    if (threadIdx.x == 0) tile_s = atomicAdd(&ctr->tile_ticket, 1u);
    __syncthreads();
    const uint32_t tile = tile_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    /* this thread's kClsItems consecutive items start at (kz0, j0, w0) */
    const long long item0 = ((long long)tile * kClsThreads + threadIdx.x) * kClsItems;
    int nitems = 0, kz0 = 0, j0 = 0, w0 = 0;
    if (item0 < total_items) {
        const long long left = total_items - item0;
        nitems = left < kClsItems ? (int)left : kClsItems;
        const long long row = item0 / WC;
        w0 = (int)(item0 - row * WC);
        kz0 = (int)(row / g.M);
        j0 = (int)(row - (long long)kz0 * g.M);
    }

    /* ---- phase 1: masks and counts -------------------------------------------------------------------------- */
    uint32_t act[kClsItems], red[kClsItems];
    uint32_t n_act = 0, n_tri = 0, n_amb = 0, n_red = 0;
    {
        int kz = kz0, j = j0, w = w0;
        RowPtrs rp = cube_row_ptrs(S, g, kz, j);
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int r = 0; r < 4; r++) lo[r] = nitems ? __ldg(rp.r[r] + w) : 0u;
#pragma unroll
        for (int q = 0; q < kClsItems; q++) {
            act[q] = 0; red[q] = 0;
            if (q < nitems) {
                const bool hi_ok = (w + 1 < g.WP);
#pragma unroll
                for (int r = 0; r < 4; r++) hi[r] = hi_ok ? __ldg(rp.r[r] + w + 1) : 0u;
                uint32_t c[8];
                corner_words(lo, hi, c);
                const uint32_t any = c[0] | c[1] | c[2] | c[3] | c[4] | c[5] | c[6] | c[7];
                const uint32_t all = c[0] & c[1] & c[2] & c[3] & c[4] & c[5] & c[6] & c[7];
                const int ncubes = g.M - w * 32; /* cubes of this word inside the row */
                uint32_t m = any & ~all & (ncubes >= 32 ? 0xffffffffu : ((1u << ncubes) - 1u));
                if (V != nullptr && m) { /* constraints: all 8 corners must be valid (marching.cpp:475-477) */
                    const RowPtrs vp = cube_row_ptrs(V, g, kz, j);
                    uint32_t vlo[4], vhi[4], v[8];
#pragma unroll
                    for (int r = 0; r < 4; r++) { vlo[r] = __ldg(vp.r[r] + w); vhi[r] = hi_ok ? __ldg(vp.r[r] + w + 1) : 0u; }
                    corner_words(vlo, vhi, v);
                    m &= v[0] & v[1] & v[2] & v[3] & v[4] & v[5] & v[6] & v[7];
                }
                act[q] = m;
                n_act += __popc(m);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    int code = code_of(c, b);
                    const int face = (int)(int8_t)__ldg((const signed char*)gtb->face + code);
                    if (face >= 0) {
                        n_amb++;
                        if (ambiguity_redirects(point_prog, g, cs, face, w * 32 + b, j, kz + g.kb)) {
                            code = 255 - code;
                            red[q] |= 1u << b;
                            n_red++;
                        }
                    }
                    n_tri += __ldg(gtb->ntri + code);
                }
                /* advance to the next item in loop order */
                if (++w == WC) {
                    w = 0;
                    if (++j == g.M) { j = 0; kz++; }
                    rp = cube_row_ptrs(S, g, kz, j);
                    if (q + 1 < nitems) {
#pragma unroll
                        for (int r = 0; r < 4; r++) lo[r] = __ldg(rp.r[r]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 4; r++) lo[r] = hi[r];
                }
            }
        }
    }

    /* ---- block-wide exclusive scan of (n_act, n_tri): shuffles inside a warp, shared memory across warps ---- */
    uint32_t inc_a = n_act, inc_t = n_tri;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t ua = __shfl_up_sync(0xffffffffu, inc_a, d), ut = __shfl_up_sync(0xffffffffu, inc_t, d);
        if (lane >= d) { inc_a += ua; inc_t += ut; }
    }
    if (lane == 31) { warp_a[warp] = inc_a; warp_t[warp] = inc_t; }
    if (__any_sync(0xffffffffu, n_amb != 0)) { /* statistics: one atomic per warp, only when needed */
        uint32_t sa = n_amb, sr = n_red;
#pragma unroll
        for (int d = 16; d; d >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, d); sr += __shfl_xor_sync(0xffffffffu, sr, d); }
        if (lane == 0) { atomicAdd(&ctr->ambiguous, (unsigned long long)sa); if (sr) atomicAdd(&ctr->redirected, (unsigned long long)sr); }
    }
    __syncthreads();
    uint32_t wbase_a = 0, wbase_t = 0, tot_a = 0, tot_t = 0;
#pragma unroll
    for (int q = 0; q < kClsThreads / 32; q++) {
        if (q < warp) { wbase_a += warp_a[q]; wbase_t += warp_t[q]; }
        tot_a += warp_a[q]; tot_t += warp_t[q];
    }
    const uint32_t excl_a = wbase_a + inc_a - n_act, excl_t = wbase_t + inc_t - n_tri;

    /* ---- decoupled look-back (warp 0) ----------------------------------------------------------------------- */
    if (warp == 0) {
        unsigned long long pa = 0, pt = 0;
        if (tile > 0) {
            if (lane == 0) {
                st.agg_active[tile] = tot_a; st.agg_tris[tile] = tot_t;
                __threadfence();
                *(volatile uint32_t*)(st.flag + tile) = 1u;
            }
            long long basei = (long long)tile - 1;
            uint32_t spins = 0;
            bool done = false;
            while (!done) {
                const long long idx = basei - lane;
                uint32_t f = idx >= 0 ? ld_volatile_u32(st.flag + idx) : 2u;
                while (__any_sync(0xffffffffu, f == 0u)) {
                    if (f == 0u) f = ld_volatile_u32(st.flag + idx);
                    if (++spins > kSpinLimit) { if (lane == 0) atomicExch(&ctr->error, 1u); f = 2u; }
                }
                __threadfence();
                const uint32_t pmask = __ballot_sync(0xffffffffu, f == 2u);
                const int first = pmask ? __ffs(pmask) - 1 : 32;
                unsigned long long va = 0, vt = 0;
                if (lane < first) { va = ld_volatile_u32(st.agg_active + idx); vt = ld_volatile_u32(st.agg_tris + idx); }
                else if (lane == first && idx >= 0) { va = ld_volatile_u64(st.inc_active + idx); vt = ld_volatile_u64(st.inc_tris + idx); }
#pragma unroll
                for (int d = 16; d; d >>= 1) { va += __shfl_xor_sync(0xffffffffu, va, d); vt += __shfl_xor_sync(0xffffffffu, vt, d); }
                pa += va; pt += vt;
                done = pmask != 0u;
                basei -= 32;
            }
        }
        if (lane == 0) {
            st.inc_active[tile] = pa + tot_a; st.inc_tris[tile] = pt + tot_t;
            __threadfence();
            *(volatile uint32_t*)(st.flag + tile) = 2u;
            base_a_s = pa; base_t_s = pt;
            if ((long long)(tile + 1) * kClsThreads * kClsItems >= total_items) { /* last tile: grand totals */
                ctr->active = pa + tot_a;
                ctr->triangles = pt + tot_t;
            }
        }
    }
    if (tot_a == 0) return; /* nothing to write in this tile (uniform per block) */
    __syncthreads();

    /* ---- phase 2: write the compacted records in loop order ------------------------------------------------- */
    if (n_act == 0) return;
    unsigned long long oa = base_a_s + excl_a, ot = base_t_s + excl_t;
    int kz = kz0, j = j0, w = w0;
#pragma unroll
    for (int q = 0; q < kClsItems; q++) {
        uint32_t m = act[q];
        if (m) {
            const RowPtrs rp = cube_row_ptrs(S, g, kz, j);
            const bool hi_ok = (w + 1 < g.WP);
            uint32_t lo[4], hi[4], c[8];
#pragma unroll
            for (int r = 0; r < 4; r++) { lo[r] = __ldg(rp.r[r] + w); hi[r] = hi_ok ? __ldg(rp.r[r] + w + 1) : 0u; }
            corner_words(lo, hi, c);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int code = code_of(c, b);
                const int tidx = (red[q] >> b) & 1u ? 255 - code : code;
                if (oa < cap_active) {
                    rec[oa] = (unsigned long long)(w * 32 + b) | ((unsigned long long)j << 12) |
                              ((unsigned long long)(kz + g.kb) << 24) | ((unsigned long long)code << 36) |
                              ((unsigned long long)tidx << 44);
                    trioff[oa] = (uint32_t)ot;
                }
                oa++;
                ot += __ldg(gtb->ntri + tidx);
            }
        }
        if (++w == WC) { w = 0; if (++j == g.M) { j = 0; kz++; } }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * K3  emit: one thread interpolates the <=12 edge vertices (and normals) of one active cube into shared memory,
 *     then the block writes the chunk's contiguous output range with one float4 per thread per store.
 * ------------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ float interp_ref(float xs, float xe, float vs, float ve, float iso) {
    /* Marching::interp, marching.cpp:437-446 */
    float v = ((iso - vs) / (ve - vs)) * (xe - xs);
    if (isinf(v) || isnan(v)) return (float)((double)xs + 0.5 * (double)(xe - xs));
    return xs + v;
}

constexpr int kEdgeStride = 37; /* 12 edges x 3 floats, +1 to spread banks */

template <bool NORMALS>
__global__ void __launch_bounds__(kEmitThreads)
emit_kernel(const Grid g, const float* __restrict__ cs, const float* __restrict__ F,
            const unsigned long long* __restrict__ rec, const uint32_t* __restrict__ trioff,
            const Counters* __restrict__ ctr, unsigned long long cap_active, unsigned long long cap_tris,
            float4* __restrict__ pos, float4* __restrict__ nrm) {
    __shared__ float epos[kEmitThreads * kEdgeStride];
    __shared__ float enrm[NORMALS ? kEmitThreads * kEdgeStride : 1];
    __shared__ uint32_t off_s[kEmitThreads + 1];
    __shared__ uint64_t triw_s[kEmitThreads];

    unsigned long long A = ctr->active, T = ctr->triangles;
    if (A > cap_active) A = cap_active; /* the host re-runs with larger buffers when counts exceed capacity */
    const unsigned long long nchunks = (A + kEmitThreads - 1) / kEmitThreads;

    for (unsigned long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        __syncthreads();
        const unsigned long long c0 = chunk * kEmitThreads;
        const int n = (int)((A - c0) < (unsigned long long)kEmitThreads ? (A - c0) : kEmitThreads);
        const int t = threadIdx.x;
        if (t < n) {
            const unsigned long long r = rec[c0 + t];
            const int i = (int)(r & 0xFFF), j = (int)((r >> 12) & 0xFFF), k = (int)((r >> 24) & 0xFFF);
            const int code = (int)((r >> 36) & 0xFF), tidx = (int)((r >> 44) & 0xFF);
            off_s[t] = trioff[c0 + t];
            triw_s[t] = mcb_tri_word(tidx);
            const int vx = i + 1, vy = j + 1, vz = k + 1;   /* indices into cs */
            const int pz = k - g.kb + 1;                     /* local plane of corner dz=0 */
            const size_t rowp = (size_t)g.P, planep = (size_t)g.NV * g.P;
            const float* f0 = F + (size_t)pz * planep + (size_t)vy * rowp + vx;
            float val[8], gx[8], gy[8], gz[8];
#pragma unroll
            for (int v = 0; v < 8; v++) {
                const int o = mcb_corner_ofs(v);
                const int dx = o & 1, dy = (o >> 1) & 1, dz = (o >> 2) & 1;
                const float* p = f0 + dz * planep + dy * rowp + dx;
                val[v] = __ldg(p);
                if (NORMALS) { /* central differences at the grid vertex, DESIGN.md §normals */
                    gx[v] = (__ldg(p + 1) - __ldg(p - 1)) / (cs[vx + dx + 1] - cs[vx + dx - 1]);
                    gy[v] = (__ldg(p + rowp) - __ldg(p - rowp)) / (cs[vy + dy + 1] - cs[vy + dy - 1]);
                    gz[v] = (__ldg(p + planep) - __ldg(p - planep)) / (cs[vz + dz + 1] - cs[vz + dz - 1]);
                }
            }
            const float cx[2] = {cs[vx], cs[vx + 1]}, cy[2] = {cs[vy], cs[vy + 1]}, cz[2] = {cs[vz], cs[vz + 1]};
            float* ep = epos + t * kEdgeStride;
            float* en = enrm + (NORMALS ? t * kEdgeStride : 0);
#pragma unroll
            for (int e = 0; e < 12; e++) {
                const int a = mcb_edge_a(e), b = mcb_edge_b(e);
                if ((((code >> a) ^ (code >> b)) & 1) == 0) continue; /* marching.cpp:563-566 */
                const int oa = mcb_corner_ofs(a), ob = mcb_corner_ofs(b);
                const float f1 = val[a], f2 = val[b];
                ep[3 * e + 0] = interp_ref(cx[oa & 1], cx[ob & 1], f1, f2, g.iso);
                ep[3 * e + 1] = interp_ref(cy[(oa >> 1) & 1], cy[(ob >> 1) & 1], f1, f2, g.iso);
                ep[3 * e + 2] = interp_ref(cz[(oa >> 2) & 1], cz[(ob >> 2) & 1], f1, f2, g.iso);
                if (NORMALS) {
                    float tt = (g.iso - f1) / (f2 - f1);
                    if (isinf(tt) || isnan(tt)) tt = 0.5f;
                    float nx = gx[a] + tt * (gx[b] - gx[a]);
                    float ny = gy[a] + tt * (gy[b] - gy[a]);
                    float nz = gz[a] + tt * (gz[b] - gz[a]);
                    const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
                    en[3 * e + 0] = nx * inv; en[3 * e + 1] = ny * inv; en[3 * e + 2] = nz * inv;
                }
            }
        }
        __syncthreads();
        if (t == 0) /* end of the chunk's output range; a capacity-truncated run is repeated by the host anyway */
            off_s[n] = (c0 + n < A) ? trioff[c0 + n] : (A == ctr->active ? (uint32_t)T : off_s[n - 1]);
        __syncthreads();
        const unsigned long long v_begin = 3ull * off_s[0], v_end = 3ull * off_s[n];
        for (unsigned long long ov = v_begin + t; ov < v_end; ov += kEmitThreads) {
            const uint32_t tri = (uint32_t)(ov / 3);
            const int corner = (int)(ov - 3ull * tri);
            int lo = 0, hi = n - 1; /* last local cube whose first triangle is <= tri */
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (off_s[mid] <= tri) lo = mid; else hi = mid - 1;
            }
            const int lt = (int)(tri - off_s[lo]);
            const int e = (int)((triw_s[lo] >> (4 * (3 * lt + corner))) & 0xF);
            if (tri < cap_tris) {
                const float* ep = epos + lo * kEdgeStride + 3 * e;
                pos[ov] = make_float4(ep[0], ep[1], ep[2], 1.0f);
                if (NORMALS) {
                    const float* en = enrm + lo * kEdgeStride + 3 * e;
                    nrm[ov] = make_float4(en[0], en[1], en[2], 0.0f);
                }
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Parity hooks (not on the hot path).
 * ------------------------------------------------------------------------------------------------------------- */
__global__ void dense_codes_kernel(const Grid g, const uint32_t* __restrict__ S, const uint32_t* __restrict__ V,
                                   uint8_t* __restrict__ code_out, uint8_t* __restrict__ tidx_out, long long ncubes) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncubes) return;
    const int i = (int)(idx % g.M);
    const long long r = idx / g.M;
    const int j = (int)(r % g.M), kz = (int)(r / g.M);
    int code = 0;
    bool ok = true;
#pragma unroll
    for (int v = 0; v < 8; v++) {
        const int o = mcb_corner_ofs(v);
        const int x = i + 1 + (o & 1), y = j + 1 + ((o >> 1) & 1), z = kz + 1 + ((o >> 2) & 1);
        const size_t widx = ((size_t)z * g.NV + y) * g.WP + (x >> 5);
        code |= (int)((S[widx] >> (x & 31)) & 1u) << v;
        if (V) ok = ok && ((V[widx] >> (x & 31)) & 1u);
    }
    if (!ok) code = 0;
    if (code_out) code_out[idx] = (uint8_t)code;
    if (tidx_out) tidx_out[idx] = (uint8_t)code; /* redirects are patched in by scatter_tidx_kernel */
}

__global__ void scatter_tidx_kernel(const Grid g, const unsigned long long* __restrict__ rec, unsigned long long A,
                                    uint8_t* __restrict__ tidx_out) {
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= A) return;
    const unsigned long long r = rec[q];
    const long long i = (long long)(r & 0xFFF), j = (long long)((r >> 12) & 0xFFF), k = (long long)((r >> 24) & 0xFFF);
    tidx_out[((k - g.kb) * g.M + j) * g.M + i] = (uint8_t)((r >> 44) & 0xFF);
}

__global__ void eval_points_kernel(const __grid_constant__ mcb_program prog, const float* __restrict__ xyz,
                                   float* __restrict__ out, long long n, float sx, float sy, float sz) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    out[q] = mcb_interp_scalar(prog.code, prog.n, prog.k, sx * xyz[3 * q], sy * xyz[3 * q + 1], sz * xyz[3 * q + 2],
                               nullptr, nullptr, nullptr);
}

} /* namespace mcbk */
