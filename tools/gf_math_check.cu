// gf_math_check — host-side check that gyroflow_b200/csrc/gf_math.cuh returns the very bits this box's
// libm returns (the functions Rust's std calls on Linux).  Build (no GPU needed):
//   nvcc -O2 -Xcompiler -ffp-contract=off,-mfma -o /tmp/gf_math_check tools/gf_math_check.cu
// or g++ -x c++ -O2 -ffp-contract=off -mfma.   Usage: gf_math_check [full|quick]
#include <cstdio>
#include <cstring>
#include <cmath>
#include <thread>
#include <vector>
#include "../gyroflow_b200/csrc/gf_math.cuh"

typedef float (*fn1)(float);
struct Res { unsigned long long bad = 0; uint32_t first = 0; };

static void sweep(fn1 ref, fn1 mine, uint64_t lo, uint64_t hi, uint64_t step, Res* r) {
    for (uint64_t u = lo; u < hi; u += step) {
        float x = gf::u2f((uint32_t)u);
        float a = ref(x), b = mine(x);
        if (gf::f2u(a) != gf::f2u(b) && !(a != a && b != b)) { if (!r->bad) r->first = (uint32_t)u; r->bad++; }
    }
}
static unsigned long long run(const char* name, fn1 ref, fn1 mine, uint64_t step) {
    const int T = (int)std::thread::hardware_concurrency() > 0 ? (int)std::thread::hardware_concurrency() : 4;
    std::vector<std::thread> th; std::vector<Res> res(T);
    const uint64_t N = 1ull << 32;
    for (int i = 0; i < T; ++i) th.emplace_back(sweep, ref, mine, N / T * i, i == T - 1 ? N : N / T * (i + 1), step, &res[i]);
    unsigned long long bad = 0; uint32_t first = 0;
    for (int i = 0; i < T; ++i) { th[i].join(); if (res[i].bad && !bad) first = res[i].first; bad += res[i].bad; }
    printf("%-6s step %llu: %llu mismatches", name, (unsigned long long)step, bad);
    if (bad) { float x = gf::u2f(first); printf("  (first x=%a [%08x] libm=%a gf=%a)", x, first, ref(x), mine(x)); }
    printf("\n");
    return bad;
}
static float rs_round_ref(float x) { return roundf(x); }
int main(int argc, char** argv) {
    const bool full = argc > 1 && !strcmp(argv[1], "full");
    const uint64_t step = full ? 1 : 257;   // 257 is odd: the quick sweep still walks every exponent/sign
    unsigned long long bad = 0;
    bad += run("atanf", atanf, gf::gf_atanf, step);
    bad += run("sinf",  sinf,  gf::gf_sinf,  step);
    bad += run("cosf",  cosf,  gf::gf_cosf,  step);
    bad += run("tanf",  tanf,  gf::gf_tanf,  step);
    bad += run("round", rs_round_ref, gf::rs_round, step);
    printf("TOTAL %llu\n", bad);
    return bad ? 1 : 0;
}
