// Host check of the identity k_diff_apply (csrc/kmeans_tc.cu: quant_sum) relies on: for |x| <= 1,
// rne(x * 2^40) == (h << 20) + l with h, l obtained from two fp32 magic-number additions.  Run by tests/test_quant_identity.py.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <cstring>
// host check of the magic-add quantisation against the plain conversion
static long long quant40(float v) { return llrint((double)v * 1099511627776.0); }
int main() {
    const float MAGIC = 12582912.f; const uint32_t MB = 0x4B400000u;
    long long bad = 0; srand(1);
    for (long long it = 0; it < 20000000LL; ++it) {
        uint32_t bits = ((uint32_t)rand() << 16) ^ (uint32_t)rand() ^ ((uint32_t)rand() << 31);
        float x; memcpy(&x, &bits, 4);
        if (!(fabsf(x) <= 1.f)) continue;
        float th = fmaf(x, 1048576.f, MAGIC);
        float r = fmaf(th - MAGIC, -9.5367431640625e-07f, x);
        float tl = fmaf(r, 1099511627776.f, MAGIC);
        uint32_t a, b; memcpy(&a, &th, 4); memcpy(&b, &tl, 4);
        long long q = ((long long)(int)(a - MB) << 20) + (long long)(int)(b - MB);
        if (q != quant40(x)) { if (bad < 5) printf("x=%a q=%lld ref=%lld\n", x, q, quant40(x)); ++bad; }
    }
    // edge values
    float edges[] = {1.f, -1.f, 0.f, -0.f, 0x1p-20f, 0x1p-21f, 0x1.8p-21f, 0x1p-41f, 0x1.8p-41f, -0x1p-41f, 0x1p-40f, 0x1.fffffep-1f, 0x1p-126f, 0x1p-149f, 0x1.000002p-21f, 0x1.fffffep-22f, 0x1.8p-40f, 0x1.4p-39f};
    for (float x : edges) {
        float th = fmaf(x, 1048576.f, MAGIC);
        float r = fmaf(th - MAGIC, -9.5367431640625e-07f, x);
        float tl = fmaf(r, 1099511627776.f, MAGIC);
        uint32_t a, b; memcpy(&a, &th, 4); memcpy(&b, &tl, 4);
        long long q = ((long long)(int)(a - MB) << 20) + (long long)(int)(b - MB);
        if (q != quant40(x)) { printf("EDGE x=%a q=%lld ref=%lld\n", x, q, quant40(x)); ++bad; }
    }
    printf("mismatches: %lld\n", bad);
    return bad != 0;
}
