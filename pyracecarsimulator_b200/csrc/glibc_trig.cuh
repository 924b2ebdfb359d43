// glibc_trig.cuh -- sinf/cosf that return the same bits as the host libm the reference calls.
//
// RayMarching::calc_range evaluates cosf(theta)/sinf(theta) per ray on the host.
// CUDA's own sincosf differs from glibc's in the last ulp for a few per cent of arguments; one
// flipped ulp can make a grazing ray visit a different cell, so for bit parity the device
// evaluates glibc's algorithm itself: glibc >= 2.28 sinf/cosf (sysdeps/ieee754/flt-32/s_sinf.c,
// s_cosf.c, sincosf.h -- the ARM optimized-routines implementation): argument widened to
// double, quadrant reduction with one multiply + one fused multiply-subtract (|x| < 120) or the
// 192-bit 4/pi table (larger), degree-7/8 polynomials in double, one rounding to float.  The
// x86-64 build selects its FMA variant on every FMA-capable CPU, which is what the explicit
// fma() below reproduce.  oracle/trig_twin.c is the same code in C and
// tests/test_trig_twin.py checks it against the host libm bit for bit.
#pragma once
#include <stdint.h>

namespace rl {

namespace trig {
// sincosf.h: __sincosf_table[0]; table[1] is the same with c0..c4 negated
constexpr double HPI_INV = 0x1.45F306DC9C883p+23;  // 2/pi * 2^24
constexpr double HPI = 0x1.921FB54442D18p0;        // pi/2
constexpr double C0 = 0x1p0;
constexpr double C1 = -0x1.ffffffd0c621cp-2;
constexpr double C2 = 0x1.55553e1068f19p-5;
constexpr double C3 = -0x1.6c087e89a359dp-10;
constexpr double C4 = 0x1.99343027bf8c3p-16;
constexpr double S1 = -0x1.555545995a603p-3;
constexpr double S2 = 0x1.1107605230bc4p-7;
constexpr double S3 = -0x1.994eb3774cf24p-13;
constexpr double PI63 = 0x1.921FB54442D18p-62;     // pi / 2^64
}  // namespace trig

// 4/pi to 192 bits, 8 new bits per entry (sincosf_data.c: __inv_pio4)
static __device__ __constant__ uint32_t g_inv_pio4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

// reduce_large: |y| >= 120 (finite).  Returns the reduced argument, quadrant in n.
static __device__ __noinline__ double trig_reduce_large(uint32_t xi, int &n_out)
{
    const uint32_t *arr = &g_inv_pio4[(xi >> 26) & 15];
    const int shift = (xi >> 23) & 7;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    uint64_t res0 = (uint64_t)(uint32_t)(xi * arr[0]);
    const uint64_t res1 = (uint64_t)xi * arr[4];
    const uint64_t res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const uint64_t n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    n_out = (int)n;
    return __dmul_rn((double)(int64_t)res0, trig::PI63);
}

// The coefficients in constant memory: as immediates every double costs two MOVs before its FMA; as
// constant-bank operands (or one LDC.64) they cost nothing (or one instruction).
static __device__ __constant__ double g_trig_k[9] = {
    trig::HPI_INV, trig::HPI, trig::C1, trig::C2, trig::C3, trig::C4, trig::S1, trig::S2, trig::S3};

// sin and cos of y with glibc's bits.  *sp = sinf(y), *cp = cosf(y).  The arithmetic is glibc's, operation
// for operation (oracle/trig_twin.c is the literal C twin); the control flow around it is trimmed because
// this sits on the setup path of every ray (~50 instead of ~70 instructions, 84.8 -> 83.7 us on config 2):
//  * glibc's |y| < pi/4 shortcut is the general path with n = 0 (the reduction leaves x untouched:
//    -0 * pi/2 + x == x), so it needs no branch of its own; the |y| < 2^-12 shortcut stays (it alone
//    returns -0 for -0);
//  * sign[nq & 3] and the negated-cosine table are applied by flipping the sign bit of the double's high word;
//  * the sin/cos swap for odd quadrants is done on the two floats.
__device__ __forceinline__ void glibc_sincosf(float y, float *sp, float *cp)
{
    const uint32_t yi = __float_as_uint(y);
    const uint32_t top = (yi >> 20) & 0x7ff;  // abstop12
    double x = (double)y;
    int n, nq;
    if (top - 0x398u < 0x42fu - 0x398u) {     // 2^-12 <= |y| < 120
        const double r = __dmul_rn(x, g_trig_k[0]);
        n = (__double2int_rz(r) + 0x800000) >> 24;
        x = fma(-(double)n, g_trig_k[1], x);
        nq = n;
    } else if (top < 0x398u) {                // |y| < 2^-12
        *sp = y;
        *cp = 1.0f;
        return;
    } else if (top < 0x7f8u) {                // finite: reduce |y|, fold the sign into nq only
        x = trig_reduce_large(yi, n);
        nq = n + (int)(yi >> 31);
    } else {                                  // inf / nan -> nan
        *sp = *cp = y - y;
        return;
    }
    const double x2 = __dmul_rn(x, x);
    // sign[nq & 3] = {1, -1, -1, 1} multiplies the sine's argument; quadrants 2, 3 negate the cosine
    const double xs = __hiloint2double(__double2hiint(x) ^ (((nq + 1) & 2) << 30), __double2loint(x));
    // sincosf_poly (sincosf.h), operation for operation
    const double x3 = __dmul_rn(xs, x2);
    const double s1 = fma(x2, g_trig_k[8], g_trig_k[7]);
    const double x7 = __dmul_rn(x3, x2);
    const double s0 = fma(x3, g_trig_k[6], xs);
    const double s = fma(x7, s1, s0);
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = fma(x2, g_trig_k[5], g_trig_k[4]);
    const double c1 = fma(x2, g_trig_k[2], trig::C0);
    const double x6 = __dmul_rn(x4, x2);
    const double cc = fma(x4, g_trig_k[3], c1);
    const double c0 = fma(x6, c2, cc);
    const double c = __hiloint2double(__double2hiint(c0) ^ ((nq & 2) << 30), __double2loint(c0));
    const float sf = (float)s, cf = (float)c;
    const bool swap = n & 1;
    *sp = swap ? cf : sf;
    *cp = swap ? sf : cf;
}

}  // namespace rl
