// Exercises the drop-in headers the way the reference's main.cpp:9-21 and drawer.cpp:785-831 do:
// Evaluator + Marching on the stack, set_evaluator, set_grid_step_size, set_scaling_*, recalculate, get_poly_data.
// Prints one line per case: name, welded vertices, triangles, FNV-1a-64 of vertex_list and tri_list bytes.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "marching.h"

static uint64_t fnv(const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    /* standard offset basis; SURVEY.md Appendix B was made with 1469598103934665603 (a digit short), selectable here */
    static const char* basis = std::getenv("FNV_BASIS");
    uint64_t h = basis ? std::strtoull(basis, nullptr, 0) : 0xCBF29CE484222325ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001B3ull; }
    return h;
}

int main(int argc, char** argv) {
    // usage: dropin_main name equation step sx sy sz iso [name equation ...]
    Evaluator evaluator;
    Marching march_maker;
    march_maker.set_evaluator(&evaluator);
    const Poly_Data* pData = march_maker.get_poly_data();   // cached once, like drawer.cpp:785-793
    for (int a = 1; a + 6 < argc; a += 7) {
        if (!evaluator.set_equation(argv[a + 1])) { std::printf("%s PARSE_ERROR\n", argv[a]); continue; }
        if (!march_maker.set_grid_step_size((float)atof(argv[a + 2]))) { std::printf("%s STEP_REJECTED\n", argv[a]); continue; }
        march_maker.set_scaling_x((float)atof(argv[a + 3]));
        march_maker.set_scaling_y((float)atof(argv[a + 4]));
        march_maker.set_scaling_z((float)atof(argv[a + 5]));
        march_maker.set_surface_constant((float)atof(argv[a + 6]));
        const bool stepping = std::strncmp(argv[a], "step_", 5) == 0; /* step-by-step mode, one cube per call */
        long calls = 0;
        if (stepping) {
            march_maker.step_by_step_mode(true);
            while (calls < 10000000 && march_maker.recalculate()) calls++;
            march_maker.step_by_step_mode(false);
            std::printf("%s_calls %ld\n", argv[a], calls);
        } else if (!march_maker.recalculate()) { std::printf("%s RECALCULATE_FAILED\n", argv[a]); continue; }
        std::printf("%s %zu %zu %016llx %016llx\n", argv[a], pData->vertex_list.size() / 3, pData->tri_list.size() / 3,
                    (unsigned long long)fnv(pData->vertex_list.data(), pData->vertex_list.size() * 4),
                    (unsigned long long)fnv(pData->tri_list.data(), pData->tri_list.size() * 4));
    }
    Evaluator e2;
    std::printf("evaluate %.9g\n", (double)e2.evaluate(3.f, 2.f, 5.f));   // "x+y" -> 5
    try { Evaluator bad("(x(y)"); std::printf("ctor NOTHROW\n"); } catch (std::exception&) { std::printf("ctor THROW\n"); }
    std::printf("step_rejected %d %d\n", (int)march_maker.set_grid_step_size(0.0005f), (int)march_maker.set_grid_step_size(0.6f));
    return 0;
}
