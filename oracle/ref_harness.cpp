/* TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
 *
 * oracle/_ref/libmcref.so: a C-ABI window onto the UNMODIFIED reference implementation
 * (/root/reference/Source/evaluator.cpp, marching.cpp, normal.h), compiled where the sources
 * lie by oracle/Makefile.  Nothing from the reference is copied into this repository: this
 * file only *calls* the reference classes (including their private members, through the
 * `#define private public` trick described in SURVEY.md §4) and reads their results back.
 *
 * Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs, as the checker / the timed CPU baseline.  The product (libmcb200.so) does not know it exists.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <iostream>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#define private public
#include "marching.h" /* reference header: class Marching, Evaluator, Poly_Data, Step_Data */
#undef private

#include "glm/glm.hpp" /* reference's vendored GLM 0.9.5.3 (Dependencies/glm) */
#include "normal.h"     /* reference: CalculateNormal (normal.h:3-42) */

/* Tables defined (non-static) by marching.cpp through its #include "marching_lookup.h". */
extern int tri_table[][16];
extern int ambiguity_check_and_redirect[][5];
extern int cube_edge_vertex_table[][2];

struct mcref {
    Evaluator eval;
    Marching march;
    std::vector<glm::vec3> normals;
    mcref() { march.set_evaluator(&eval); }
};

/* The reference's grid loop (marching.cpp:372-377), run on one axis: returns the loop-variable
 * values.  c.size() == M (cubes per axis); the far corner of the last cube is c[M-1]+step. */
static std::vector<float> ref_axis(float step) {
    std::vector<float> c;
    float lower_bound = -1.0;
    float upper_bound = 1.0 + 0.5 * step;
    for (float v = lower_bound; v <= upper_bound; v += step) c.push_back(v);
    return c;
}

static int row_matches(const int* row, const std::vector<int>& edges) {
    size_t n = 0;
    while (n < 16 && row[n] != -1) n++;
    if (n != edges.size()) return 0;
    for (size_t i = 0; i < n; i++)
        if (row[i] != edges[i]) return 0;
    return 1;
}

extern "C" {

mcref* mcref_create(void) { return new mcref(); }
void mcref_destroy(mcref* h) { delete h; }

/* Evaluator::tokenize accept/reject on a scratch evaluator (evaluator.cpp:139-237). */
int mcref_parse_ok(const char* eq) {
    Evaluator e;
    return e.tokenize(std::string(eq)) ? 1 : 0;
}

int mcref_set_equation(mcref* h, const char* eq) { return h->eval.set_equation(std::string(eq)) ? 1 : 0; }

float mcref_evaluate(mcref* h, float x, float y, float z) { return h->eval.evaluate(x, y, z); }

void mcref_eval_points(mcref* h, const float* xyz, float* out, long n) {
    for (long i = 0; i < n; i++) out[i] = h->eval.evaluate(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

/* Marching::evaluate (marching.cpp:209-224): applies the per-axis scaling first. */
void mcref_march_eval_points(mcref* h, const float* xyz, float* out, long n) {
    for (long i = 0; i < n; i++) out[i] = h->march.evaluate(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

int mcref_set_step(mcref* h, float s) { return h->march.set_grid_step_size(s) ? 1 : 0; }
/* D4: steps below 0.001 (2048^3) are rejected by the setter; write the field directly. */
void mcref_force_step(mcref* h, float s) { h->march.grid_step_size = s; }
float mcref_get_step(mcref* h) { return h->march.get_grid_size(); }
void mcref_set_scale(mcref* h, float sx, float sy, float sz) {
    h->march.set_scaling_x(sx); h->march.set_scaling_y(sy); h->march.set_scaling_z(sz);
}
void mcref_set_iso(mcref* h, float c) { h->march.set_surface_constant(c); }

/* D5: Marching::set_constraint has no return on success (marching.cpp:173-200) and traps under g++,
 * so the constraint is installed field by field.  op: 0 '>', 1 '<', 2 '>=', 3 '<=' (Comp_Op order). */
int mcref_set_constraint(mcref* h, int i, const char* lhs, int op, float rhs, int in_use) {
    if (i < 0 || i > 2 || op < 0 || op > 3) return 0;
    Constraint& c = h->march.constraints[i];
    if (!c.eval.set_equation(std::string(lhs))) return 0;
    c.valid = true; c.lhs = lhs; c.op = (Comp_Op)op; c.rhs = rhs; c.in_use = in_use != 0;
    return 1;
}

/* Repeating-surface mode (marching.cpp:156-170, 481-494) through the reference's own setters. */
int mcref_set_repeat(mcref* h, int on, float distance) {
    if (on && !h->march.set_surface_repeat_step_distance(distance)) return 0;
    return h->march.repeating_surface_mode(on != 0) ? 1 : 0;
}
int mcref_recalculate(mcref* h) { return h->march.recalculate() ? 1 : 0; }
/* Step-by-step mode (marching.cpp:386-428): recalculate() once per cube until it reports that it has finished, then
 * back to the full-grid mode.  Returns the number of calls that returned true. */
long mcref_step_all(mcref* h, long max_calls) {
    h->march.step_by_step_mode(true);
    long n = 0;
    while (n < max_calls && h->march.recalculate()) n++;
    h->march.step_by_step_mode(false);
    return n;
}
/* Seed mode (marching.cpp:42-137, 310-331): Marching::set_seed + seed_mode(true) + recalculate(), then back to the
 * full-grid mode.  Returns 0 when set_seed rejects the point. */
int mcref_seed_recalculate(mcref* h, float sx, float sy, float sz) {
    if (!h->march.set_seed(sx, sy, sz)) return 0;
    h->march.seed_mode(true);
    const bool ok = h->march.recalculate();
    h->march.seed_mode(false);
    return ok ? 1 : 0;
}
long mcref_num_vertices(mcref* h) { return (long)h->march.poly_data.vertex_list.size() / 3; }
long mcref_num_triangles(mcref* h) { return (long)h->march.poly_data.tri_list.size() / 3; }
void mcref_copy_mesh(mcref* h, float* verts, unsigned* tris) {
    const Poly_Data* p = h->march.get_poly_data();
    if (verts) std::memcpy(verts, p->vertex_list.data(), p->vertex_list.size() * sizeof(float));
    if (tris) std::memcpy(tris, p->tri_list.data(), p->tri_list.size() * sizeof(unsigned));
}
/* normal.h:3-42 on the current Poly_Data; out = 3 floats per welded vertex. */
void mcref_normals(mcref* h, float* out) {
    std::vector<glm::vec3> n = CalculateNormal(h->march.get_poly_data());
    for (size_t i = 0; i < n.size(); i++) { out[3 * i] = n[i].x; out[3 * i + 1] = n[i].y; out[3 * i + 2] = n[i].z; }
}

/* Loop-variable values of the reference grid loop; returns M. coords gets M+1 floats when cap allows
 * (the extra one is the far corner of the last cube, c[M-1]+step as calculate_step computes it). */
int mcref_grid_coords(mcref* h, float* coords, int cap) {
    float step = h->march.grid_step_size;
    std::vector<float> c = ref_axis(step);
    int M = (int)c.size();
    if (coords && cap >= M + 1) {
        for (int i = 0; i < M; i++) coords[i] = c[i];
        coords[M] = c[M - 1] + step;
    }
    return M;
}

/* Per-cube sweep over cube layers k in [k0,k1) in the reference's loop order, through the PRIVATE
 * Marching::calculate_step (marching.cpp:456-595) and, when weld!=0, add_step_to_poly_data (:599-623).
 * Per cube (index within the sweep, x fastest): raw cube_code, effective tri_table row, triangle count.
 * soup (optional): 9 floats per triangle, in emission order.  Returns the number of triangles; if that
 * exceeds soup_cap_tris the soup is truncated (the count is still exact).
 * A cube skipped by a constraint reports code = table_idx = 0. */
long mcref_sweep(mcref* h, int k0, int k1, uint8_t* code, uint8_t* tidx, uint8_t* ntri, float* corner_vals,
                 float* soup, long soup_cap_tris, int weld, long* n_active, long* n_ambiguous, long* n_redirected) {
    Marching& m = h->march;
    float step = m.grid_step_size;
    std::vector<float> c = ref_axis(step);
    int M = (int)c.size();
    if (k0 < 0) k0 = 0;
    if (k1 > M) k1 = M;
    if (weld) m.reset_all_data();
    bool any_constraint = false;
    for (size_t i = 0; i < m.constraints.size(); i++) any_constraint |= (m.constraints[i].valid && m.constraints[i].in_use);
    Step_Data* sd = &m.poly_data.step_data;
    long T = 0, A = 0, AMB = 0, RED = 0, idx = 0;
    std::vector<int> edges;
    for (int k = k0; k < k1; k++)
        for (int j = 0; j < M; j++)
            for (int i = 0; i < M; i++, idx++) {
                m.calculate_step(c[i], c[j], c[k]);
                if (weld) m.add_step_to_poly_data();
                int cc = 0, ti = 0, nt = 0;
                bool skipped = false;
                if (any_constraint) {
                    for (int v = 0; v < 8 && !skipped; v++)
                        skipped = !m.check_constraints(sd->corner_coords[3 * v], sd->corner_coords[3 * v + 1], sd->corner_coords[3 * v + 2]);
                }
                if (!skipped) {
                    float iso = m.is_repeating_surface ? sd->surf_constant : m.surface_constant;
                    for (int v = 0; v < 8; v++)
                        if (sd->corner_values[v] > iso) cc |= (1 << v);
                    nt = (int)sd->tri_vlist.size() / 3;
                    ti = cc;
                    if (cc != 0 && cc != 255) {
                        A++;
                        edges.clear();
                        for (size_t t = 0; t < sd->tri_vlist.size(); t++) edges.push_back(sd->edge_list[sd->tri_vlist[t]]);
                        int alt = ambiguity_check_and_redirect[cc][0];
                        if (alt >= 0) AMB++;
                        if (row_matches(tri_table[cc], edges)) ti = cc;
                        else if (alt >= 0 && row_matches(tri_table[alt], edges)) { ti = alt; RED++; }
                        else ti = -1; /* cannot happen; surfaces as a parity failure */
                    }
                    if (corner_vals)
                        for (int v = 0; v < 8; v++) corner_vals[8 * idx + v] = sd->corner_values[v];
                }
                if (code) code[idx] = (uint8_t)cc;
                if (tidx) tidx[idx] = (uint8_t)ti;
                if (ntri) ntri[idx] = (uint8_t)nt;
                if (soup)
                    for (int t = 0; t < nt; t++) {
                        if (T + t >= soup_cap_tris) break;
                        for (int v = 0; v < 3; v++) {
                            int li = sd->tri_vlist[3 * t + v];
                            for (int a = 0; a < 3; a++) soup[9 * (T + t) + 3 * v + a] = sd->intersect_coord[3 * li + a];
                        }
                    }
                T += nt;
            }
    if (n_active) *n_active = A;
    if (n_ambiguous) *n_ambiguous = AMB;
    if (n_redirected) *n_redirected = RED;
    return T;
}

/* the same for cube rows [row0, row1) (row = k * M + j): lets a test spread one large layer over several threads, each with
 * its own Evaluator + Marching pair */
long mcref_sweep_rows(mcref* h, long row0, long row1, uint8_t* code, uint8_t* tidx, uint8_t* ntri, float* corner_vals,
                 float* soup, long soup_cap_tris, int weld, long* n_active, long* n_ambiguous, long* n_redirected) {
    Marching& m = h->march;
    float step = m.grid_step_size;
    std::vector<float> c = ref_axis(step);
    int M = (int)c.size();
    if (row0 < 0) row0 = 0;
    if (row1 > (long)M * M) row1 = (long)M * M;
    if (weld) m.reset_all_data();
    bool any_constraint = false;
    for (size_t i = 0; i < m.constraints.size(); i++) any_constraint |= (m.constraints[i].valid && m.constraints[i].in_use);
    Step_Data* sd = &m.poly_data.step_data;
    long T = 0, A = 0, AMB = 0, RED = 0, idx = 0;
    std::vector<int> edges;
    for (long row = row0; row < row1; row++)
        for (int i = 0, j = (int)(row % M), k = (int)(row / M); i < M; i++, idx++) {
                m.calculate_step(c[i], c[j], c[k]);
                if (weld) m.add_step_to_poly_data();
                int cc = 0, ti = 0, nt = 0;
                bool skipped = false;
                if (any_constraint) {
                    for (int v = 0; v < 8 && !skipped; v++)
                        skipped = !m.check_constraints(sd->corner_coords[3 * v], sd->corner_coords[3 * v + 1], sd->corner_coords[3 * v + 2]);
                }
                if (!skipped) {
                    float iso = m.is_repeating_surface ? sd->surf_constant : m.surface_constant;
                    for (int v = 0; v < 8; v++)
                        if (sd->corner_values[v] > iso) cc |= (1 << v);
                    nt = (int)sd->tri_vlist.size() / 3;
                    ti = cc;
                    if (cc != 0 && cc != 255) {
                        A++;
                        edges.clear();
                        for (size_t t = 0; t < sd->tri_vlist.size(); t++) edges.push_back(sd->edge_list[sd->tri_vlist[t]]);
                        int alt = ambiguity_check_and_redirect[cc][0];
                        if (alt >= 0) AMB++;
                        if (row_matches(tri_table[cc], edges)) ti = cc;
                        else if (alt >= 0 && row_matches(tri_table[alt], edges)) { ti = alt; RED++; }
                        else ti = -1; /* cannot happen; surfaces as a parity failure */
                    }
                    if (corner_vals)
                        for (int v = 0; v < 8; v++) corner_vals[8 * idx + v] = sd->corner_values[v];
                }
                if (code) code[idx] = (uint8_t)cc;
                if (tidx) tidx[idx] = (uint8_t)ti;
                if (ntri) ntri[idx] = (uint8_t)nt;
                if (soup)
                    for (int t = 0; t < nt; t++) {
                        if (T + t >= soup_cap_tris) break;
                        for (int v = 0; v < 3; v++) {
                            int li = sd->tri_vlist[3 * t + v];
                            for (int a = 0; a < 3; a++) soup[9 * (T + t) + 3 * v + a] = sd->intersect_coord[3 * li + a];
                        }
                    }
                T += nt;
            }
    if (n_active) *n_active = A;
    if (n_ambiguous) *n_ambiguous = AMB;
    if (n_redirected) *n_redirected = RED;
    return T;
}

/* CPU baseline: nthreads independent Evaluator+Marching pairs (they share no state), each running the reference's
 * calculate_step + add_step_to_poly_data over its own contiguous share of `nrows` cube rows starting at row0
 * (row = k*M + j, M cubes each, loop order).  Returns wall seconds; *cubes / *tris = totals.  nthreads==1 with
 * row0=0, nrows=M*M is exactly the work of Marching::recalculate()'s full-grid branch (marching.cpp:368-384). */
double mcref_timed_rows_mt(const char* eq, float step, float sx, float sy, float sz, float iso, long row0, long nrows,
                           int nthreads, long* cubes, long* tris) {
    std::vector<float> c = ref_axis(step);
    long M = (long)c.size();
    if (row0 < 0) row0 = 0;
    if (row0 > M * M) row0 = M * M;
    if (row0 + nrows > M * M) nrows = M * M - row0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nrows) nthreads = nrows > 0 ? (int)nrows : 1;
    std::vector<long> t_tris(nthreads, 0);
    std::vector<mcref*> hs(nthreads);
    for (int t = 0; t < nthreads; t++) {
        hs[t] = new mcref();
        hs[t]->eval.set_equation(std::string(eq));
        hs[t]->march.grid_step_size = step;
        hs[t]->march.set_scaling_x(sx); hs[t]->march.set_scaling_y(sy); hs[t]->march.set_scaling_z(sz);
        hs[t]->march.set_surface_constant(iso);
    }
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        long a = row0 + nrows * t / nthreads, b = row0 + nrows * (t + 1) / nthreads;
        th.emplace_back([&, t, a, b]() {
            Marching& m = hs[t]->march;
            m.reset_all_data();
            for (long r = a; r < b; r++) {
                const float z0 = c[r / M], y0 = c[r % M];
                for (long i = 0; i < M; i++) {
                    m.calculate_step(c[i], y0, z0);
                    m.add_step_to_poly_data();
                }
            }
            t_tris[t] = (long)m.poly_data.tri_list.size() / 3;
        });
    }
    for (auto& x : th) x.join();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long T = 0;
    for (int t = 0; t < nthreads; t++) { T += t_tris[t]; delete hs[t]; }
    if (cubes) *cubes = nrows * M;
    if (tris) *tris = T;
    return sec;
}

/* Timed unmodified Marching::recalculate(); returns wall seconds. */
double mcref_timed_recalculate(mcref* h) {
    auto t0 = std::chrono::steady_clock::now();
    h->march.recalculate();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* Read-only views of the reference tables, for tests that pin the product's packed tables. */
void mcref_tables(int* tri /*256*16*/, int* amb /*256*5*/, int* edge /*12*2*/) {
    if (tri) for (int i = 0; i < 256; i++) for (int j = 0; j < 16; j++) tri[16 * i + j] = tri_table[i][j];
    if (amb) for (int i = 0; i < 256; i++) for (int j = 0; j < 5; j++) amb[5 * i + j] = ambiguity_check_and_redirect[i][j];
    if (edge) for (int i = 0; i < 12; i++) for (int j = 0; j < 2; j++) edge[2 * i + j] = cube_edge_vertex_table[i][j];
}

} /* extern "C" */
