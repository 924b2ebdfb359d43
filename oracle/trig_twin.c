/*
 * oracle/trig_twin.c -- C twin of pyracecarsimulator_b200/csrc/glibc_trig.cuh.
 * TEST INFRASTRUCTURE ONLY.  The oracle's marcher calls the host libm cosf/sinf exactly like
 * the reference does; this file exists so that tests/test_trig_twin.py can show, on the CPU,
 * that the algorithm the CUDA kernels evaluate (glibc >= 2.28 sinf/cosf: sincosf.h, s_sinf.c,
 * s_cosf.c, FMA build) returns the host libm's bits.  Same statement order as the .cuh.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_EXPORT __attribute__((visibility("default")))

static const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
static const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                    C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
static const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
static const double PI63 = 0x1.921FB54442D18p-62;
static const uint32_t INV_PIO4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

static double sin_poly(double x, double x2)
{
    double x3 = x * x2, s1 = fma(x2, S3, S2), x7 = x3 * x2, s = fma(x3, S1, x);
    return fma(x7, s1, s);
}

static double cos_poly(double x2)
{
    double x4 = x2 * x2, c2 = fma(x2, C4, C3), c1 = fma(x2, C1, C0), x6 = x4 * x2, c = fma(x4, C2, c1);
    return fma(x6, c2, c);
}

static double reduce_large(uint32_t xi, int *np)
{
    const uint32_t *arr = &INV_PIO4[(xi >> 26) & 15];
    int shift = (xi >> 23) & 7;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    uint64_t res0 = (uint32_t)(xi * arr[0]);
    uint64_t res1 = (uint64_t)xi * arr[4], res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    uint64_t n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    *np = (int)n;
    return (double)(int64_t)res0 * PI63;
}

ORC_EXPORT void orc_twin_sincosf(float y, float *sp, float *cp)
{
    uint32_t yi;
    memcpy(&yi, &y, 4);
    const uint32_t top = (yi >> 20) & 0x7ff;
    double x = (double)y;
    if (top < 0x3f4u) {
        if (top < 0x398u) { *sp = y; *cp = 1.0f; return; }
        const double x2 = x * x;
        *sp = (float)sin_poly(x, x2);
        *cp = (float)cos_poly(x2);
        return;
    }
    int n, nq;
    if (top < 0x42fu) {
        const double r = x * HPI_INV;
        n = ((int32_t)r + 0x800000) >> 24;
        x = fma(-(double)n, HPI, x);
        nq = n;
    } else if (top < 0x7f8u) {
        x = reduce_large(yi, &n);
        nq = n + (int)(yi >> 31);
    } else {
        *sp = *cp = y - y;
        return;
    }
    const double xs = ((nq + 1) & 2) ? -x : x;
    const double x2 = x * x;
    const double s = sin_poly(xs, x2);
    double c = cos_poly(x2);
    if (nq & 2) c = -c;
    if (n & 1) { *sp = (float)c; *cp = (float)s; }
    else { *sp = (float)s; *cp = (float)c; }
}

/* Compare against the host libm over `count` bit patterns u = start, start+stride, ... (both  */
/* signs); returns the number of arguments where either sinf or cosf differs in any bit.       */
ORC_EXPORT uint64_t orc_twin_mismatches(uint32_t start, uint32_t stride, uint64_t count)
{
    uint64_t bad = 0;
    uint32_t u = start;
    for (uint64_t i = 0; i < count && u < 0x7f800000u; ++i, u += stride) {
        for (uint32_t sg = 0; sg < 2; ++sg) {
            uint32_t b = u | (sg << 31);
            float y, s, c, ls, lc;
            memcpy(&y, &b, 4);
            orc_twin_sincosf(y, &s, &c);
            ls = sinf(y);
            lc = cosf(y);
            if (memcmp(&s, &ls, 4) != 0 || memcmp(&c, &lc, 4) != 0) ++bad;
        }
    }
    return bad;
}
