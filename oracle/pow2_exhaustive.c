/* TEST INFRASTRUCTURE ONLY.  Exhaustive proof of the x^2 fast path in csrc/mcb_pow.h:
 * for every one of the 2^32 float bit patterns, mcb_powf(x, 2) (fast path + fallback) must equal the full
 * glibc-restated algorithm mcb_powf_full(x, 2) bit for bit; it also reports how often the fast path is taken and
 * the largest rounding-residual fraction at which the correctly rounded square differs from powf.
 *   build + run:  make -C oracle pow2     (about 25 s on 8 cores)
 *   sampled run:  oracle/pow2_exhaustive 64     (every 64th input; used by tests/test_pow.py)
 *   vs libm:      oracle/pow2_exhaustive 1 20000000   (pseudo-random (x,y) pairs against this machine's powf)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../marching-cube-for-implicit-surfaces_b200/csrc/mcb_pow.h"

typedef struct { uint64_t lo, hi, stride, bad, fast, n, libm_bad; } Job;

static void* run(void* a) {
    Job* j = (Job*)a;
    volatile float two = 2.0f; /* keeps the compiler from folding powf(x, 2) into x*x */
    for (uint64_t u = j->lo; u < j->hi; u += j->stride) {
        uint32_t b = (uint32_t)u;
        float x, r;
        memcpy(&x, &b, 4);
        const float full = mcb_powf_full(x, 2.0f), fastp = mcb_powf(x, 2.0f);
        j->n++;
        j->fast += mcb_pow2_try(x, &r) ? 1 : 0;
        if (!(full != full && fastp != fastp) && memcmp(&full, &fastp, 4)) j->bad++;
        if ((u & 1023) == 0) { /* and the restatement itself against this machine's libm */
            const float lm = powf(x, two);
            if (!(full != full && lm != lm) && memcmp(&full, &lm, 4)) j->libm_bad++;
        }
    }
    return 0;
}

/* mcb_powf against this machine's libm on pseudo-random (x, y): raw bit patterns, and pairs drawn from the ranges the
 * equation language actually produces (|x| in [2^-20, 2^20], small integer, half-integer and arbitrary exponents) */
static uint64_t rng(uint64_t* s) { *s = *s * 6364136223846793005ull + 1442695040888963407ull; return *s >> 16; }
static int random_pairs(uint64_t n) {
    uint64_t st = 0x9E3779B97F4A7C15ull, bad = 0;
    for (uint64_t q = 0; q < n; q++) {
        float x, y;
        uint32_t bx = (uint32_t)rng(&st), by = (uint32_t)rng(&st);
        const unsigned kind = (unsigned)(rng(&st) & 7u);
        if (kind >= 2) { /* realistic: exponent of x in [-20, 20] */
            bx = (bx & 0x807fffffu) | ((107u + (uint32_t)(rng(&st) % 41u)) << 23);
            memcpy(&x, &bx, 4);
            const int r = (int)(rng(&st) % 33u) - 16;
            y = kind == 2 ? (float)r : kind == 3 ? (float)r * 0.5f : kind == 4 ? 2.0f : (float)r + (float)(by & 0xffff) / 65536.0f;
        } else { memcpy(&x, &bx, 4); memcpy(&y, &by, 4); }
        volatile float vy = y;
        const float a = mcb_powf(x, y), b = powf(x, vy);
        if (!(a != a && b != b) && memcmp(&a, &b, 4)) {
            if (bad < 5) printf("  x=%a y=%a mcb=%a libm=%a\n", x, y, a, b);
            bad++;
        }
    }
    printf("random pairs %llu mismatches_vs_libm %llu\n", (unsigned long long)n, (unsigned long long)bad);
    return bad != 0;
}

int main(int argc, char** argv) {
    enum { NT = 16 };
    if (argc > 2) return random_pairs(strtoull(argv[2], 0, 0));
    const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 0) : 1;
    pthread_t th[NT];
    Job jobs[NT];
    memset(jobs, 0, sizeof jobs);
    for (int t = 0; t < NT; t++) {
        jobs[t].lo = (1ull << 32) * t / NT; jobs[t].hi = (1ull << 32) * (t + 1) / NT; jobs[t].stride = stride;
        pthread_create(&th[t], 0, run, &jobs[t]);
    }
    uint64_t bad = 0, fast = 0, n = 0, lb = 0;
    for (int t = 0; t < NT; t++) { pthread_join(th[t], 0); bad += jobs[t].bad; fast += jobs[t].fast; n += jobs[t].n; lb += jobs[t].libm_bad; }
    printf("inputs %llu fast_path %llu (%.4f %%) mismatches_fast_vs_full %llu mismatches_full_vs_libm_sampled %llu\n",
           (unsigned long long)n, (unsigned long long)fast, 100.0 * (double)fast / (double)n, (unsigned long long)bad, (unsigned long long)lb);
    return bad || lb ? 1 : 0;
}
