/* TEST INFRASTRUCTURE ONLY — runs the product's bytecode on the HOST so that the lowering (mcb_lower.cpp) can be
 * checked against the compiled reference in the CPU-only test tier (there is no GPU in the build container).
 * The product never links or calls this; on the GPU the same programs are interpreted by mcb_kernels.cuh.
 * Built into oracle/libmcoracle_host.so by oracle/Makefile with -ffp-contract=off.
 */
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../marching-cube-for-implicit-surfaces_b200/csrc/mcb_lower.h"
#include "../marching-cube-for-implicit-surfaces_b200/csrc/mcb_interval.h"
#include "../include/mcb.h"

namespace {
/* folds constant slots with the host build of the same interpreter */
void fold(mcb::Compiled& c) {
    for (const mcb::Slot& s : c.slots)
        if (s.axis < 0)
            c.kpool[s.kindex] = mcb_interp_scalar(c.slot_code.data() + s.code_begin, s.code_len, c.kpool.data(), 0, 0, 0,
                                                  nullptr, nullptr, nullptr);
}
}  // namespace

extern "C" {

/* which: 0 = point program at n arbitrary points (xyz = 3n floats) */
int mcoh_eval_points(const char* eq, const float* xyz, float* out, long n) {
    mcb::Compiled c;
    int rc = mcb::compile(eq, c, nullptr);
    if (rc != MCB_OK) return rc;
    fold(c);
    for (long i = 0; i < n; i++)
        out[i] = mcb_interp_scalar(c.point_code.data(), (int)c.point_code.size(), c.kpool.data(), xyz[3 * i], xyz[3 * i + 1],
                                   xyz[3 * i + 2], nullptr, nullptr, nullptr);
    return MCB_OK;
}

/* grid program over the tensor grid cx[nx] x cy[ny] x cz[nz] (already scaled coordinates), x fastest */
int mcoh_eval_grid(const char* eq, const float* cx, int nx, const float* cy, int ny, const float* cz, int nz, float* out) {
    mcb::Compiled c;
    int rc = mcb::compile(eq, c, nullptr);
    if (rc != MCB_OK) return rc;
    fold(c);
    const float* ax[3] = {cx, cy, cz};
    int an[3] = {nx, ny, nz};
    std::vector<std::vector<float>> tab[3];
    for (int a = 0; a < 3; a++) tab[a].assign(c.n_axis_slots[a], std::vector<float>(an[a]));
    for (const mcb::Slot& s : c.slots) {
        if (s.axis < 0) continue;
        for (int i = 0; i < an[s.axis]; i++) {
            float v = ax[s.axis][i];
            tab[s.axis][s.kindex][i] = mcb_interp_scalar(c.slot_code.data() + s.code_begin, s.code_len, c.kpool.data(), v, v, v,
                                                         nullptr, nullptr, nullptr);
        }
    }
    std::vector<float> tx(c.n_axis_slots[0] + 1), ty(c.n_axis_slots[1] + 1), tz(c.n_axis_slots[2] + 1);
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                for (int s = 0; s < c.n_axis_slots[0]; s++) tx[s] = tab[0][s][i];
                for (int s = 0; s < c.n_axis_slots[1]; s++) ty[s] = tab[1][s][j];
                for (int s = 0; s < c.n_axis_slots[2]; s++) tz[s] = tab[2][s][k];
                /* the fused accumulator form is what the grid kernel runs; the postfix form must agree with it */
                const float fused = mcb_interp_fused_scalar(c.grid_fused.data(), (int)c.grid_fused.size(), c.kpool.data(),
                                                            cx[i], cy[j], cz[k], tx.data(), ty.data(), tz.data());
                const float postfix = mcb_interp_scalar(c.grid_code.data(), (int)c.grid_code.size(), c.kpool.data(),
                                                        cx[i], cy[j], cz[k], tx.data(), ty.data(), tz.data());
                if (std::memcmp(&fused, &postfix, 4) != 0 && !(fused != fused && postfix != postfix)) return -100;
                out[((size_t)k * ny + j) * nx + i] = fused;
            }
    return MCB_OK;
}

/* What the product's K0 kernels leave on the device for a grid program: the folded constant pool (MCB_MAX_K floats) and
 * the per-axis tables of the hoisted subtrees in the product's layout tables[(axis * spa + slot) * P + index], evaluated at
 * the (already scaled) coordinates ax[axis][0..P).  Returns spa (slots per axis) or a negative status. */
int mcoh_tables(const char* eq, const float* cx, const float* cy, const float* cz, int P, float* tables, int tables_cap, float* kpool) {
    mcb::Compiled c;
    int rc = mcb::compile(eq, c, nullptr);
    if (rc != MCB_OK) return rc;
    fold(c);
    const int spa = std::max(1, std::max(c.n_axis_slots[0], std::max(c.n_axis_slots[1], c.n_axis_slots[2])));
    if ((long)3 * spa * P > tables_cap) return MCB_E_CAPACITY;
    const float* ax[3] = {cx, cy, cz};
    for (const mcb::Slot& s : c.slots) {
        if (s.axis < 0) continue;
        for (int i = 0; i < P; i++) {
            const float v = ax[s.axis][i];
            tables[((size_t)s.axis * spa + s.kindex) * P + i] =
                mcb_interp_scalar(c.slot_code.data() + s.code_begin, s.code_len, c.kpool.data(), v, v, v, nullptr, nullptr, nullptr);
        }
    }
    for (size_t i = 0; i < c.kpool.size() && i < MCB_MAX_K; i++) kpool[i] = c.kpool[i];
    return spa;
}

int mcoh_depths(const char* eq, int* point_depth, int* grid_depth, int* n_point, int* n_grid, int* n_slots) {
    mcb::Compiled c;
    int rc = mcb::compile(eq, c, nullptr);
    if (rc != MCB_OK) return rc;
    *point_depth = c.point_depth; *grid_depth = c.grid_depth;
    *n_point = (int)c.point_code.size(); *n_grid = (int)c.grid_code.size(); *n_slots = (int)c.slots.size();
    return MCB_OK;
}

/* Brute-force check of the interval proof (csrc/mcb_interval.h) on the tensor grid cx[n] x cy[n] x cz[n] (already scaled
 * coordinates) cut into boxes of bxs x bys x bzs vertices: per box the class the interval evaluation proves (0 unknown,
 * 1 all signs 0, 2 all signs 1) against the sign of f at EVERY vertex, and the enclosure lo <= f <= hi.
 * stats[0] boxes, [1] proven sign 0, [2] proven sign 1, [3] boxes that really are uniform, [4] WRONG proofs (must be 0),
 * [5] enclosure violations (must be 0). */
int mcoh_interval_check(const char* eq, const float* cx, const float* cy, const float* cz, int n, int bxs, int bys, int bzs,
                        float iso, long long* stats) {
    mcb::Compiled c;
    int rc = mcb::compile(eq, c, nullptr);
    if (rc != MCB_OK) return rc;
    fold(c);
    const float* ax[3] = {cx, cy, cz};
    const int bs[3] = {bxs, bys, bzs};
    const int spa = std::max(1, std::max(c.n_axis_slots[0], std::max(c.n_axis_slots[1], c.n_axis_slots[2])));
    std::vector<std::vector<float>> tab[3];
    for (int a = 0; a < 3; a++) tab[a].assign(spa, std::vector<float>(n, 0.f));
    for (const mcb::Slot& s : c.slots) {
        if (s.axis < 0) continue;
        for (int i = 0; i < n; i++) {
            const float v = ax[s.axis][i];
            tab[s.axis][s.kindex][i] = mcb_interp_scalar(c.slot_code.data() + s.code_begin, s.code_len, c.kpool.data(), v, v, v, nullptr, nullptr, nullptr);
        }
    }
    int nbk[3], nb = 1;
    for (int a = 0; a < 3; a++) { nbk[a] = (n + bs[a] - 1) / bs[a]; nb = std::max(nb, nbk[a]); }
    std::vector<mcb_ival> B((size_t)3 * spa * nb);
    for (int a = 0; a < 3; a++)
        for (int s = 0; s < c.n_axis_slots[a]; s++)
            for (int b = 0; b < nbk[a]; b++) {
                float lo = INFINITY, hi = -INFINITY;
                bool ok = true;
                for (int i = b * bs[a]; i < std::min(n, (b + 1) * bs[a]); i++) {
                    const float v = tab[a][s][i];
                    if (!mcb_iv_finite(v)) ok = false;
                    lo = std::min(lo, v); hi = std::max(hi, v);
                }
                mcb_ival iv; iv.lo = ok ? lo : -INFINITY; iv.hi = ok ? hi : INFINITY;
                B[((size_t)a * spa + s) * nb + b] = iv;
            }
    for (int q = 0; q < 6; q++) stats[q] = 0;
    std::vector<float> tx(spa), ty(spa), tz(spa);
    for (int bz = 0; bz < nbk[2]; bz++)
        for (int by = 0; by < nbk[1]; by++)
            for (int bx = 0; bx < nbk[0]; bx++) {
                mcb_ival iv;
                const int cls = mcb_interval_class(c.grid_fused.data(), (int)c.grid_fused.size(), c.kpool.data(), B.data(), spa, nb, bx, by, bz, iso, &iv);
                stats[0]++; if (cls == 1) stats[1]++; if (cls == 2) stats[2]++;
                int n1 = 0, nall = 0;
                for (int k = bz * bzs; k < std::min(n, (bz + 1) * bzs); k++)
                    for (int j = by * bys; j < std::min(n, (by + 1) * bys); j++)
                        for (int i = bx * bxs; i < std::min(n, (bx + 1) * bxs); i++) {
                            for (int s = 0; s < spa; s++) { tx[s] = tab[0][s][i]; ty[s] = tab[1][s][j]; tz[s] = tab[2][s][k]; }
                            const float f = mcb_interp_fused_scalar(c.grid_fused.data(), (int)c.grid_fused.size(), c.kpool.data(), cx[i], cy[j], cz[k], tx.data(), ty.data(), tz.data());
                            nall++; n1 += f > iso ? 1 : 0;
                            if (cls != 0 && !(f >= iv.lo && f <= iv.hi)) stats[5]++;
                        }
                if (n1 == 0 || n1 == nall) stats[3]++;
                if ((cls == 1 && n1 != 0) || (cls == 2 && n1 != nall)) stats[4]++;
            }
    return MCB_OK;
}
}
