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
}
