/*
 * oracle/philox_oracle.c -- CPU restatement of the rollout action schedule the library draws on the
 * device (csrc/car.cu rollout_actions_kernel).  TEST INFRASTRUCTURE ONLY (see the header of
 * rangelib_oracle.c).
 *
 * What it follows: scripts/mcts.py:216-222 draws, on every 10th rollout step,
 *     rand_steer = np.random.uniform(-max_steer_ang, max_steer_ang)
 *     rand_speed = np.random.uniform(0, max_speed)
 * from numpy's global (unseeded) Mersenne Twister, so there is no reference stream to reproduce;
 * SURVEY.md 8d fixes a counter-based generator instead (Philox, seed 42) so that checker and GPU see
 * identical actions.  The generator is Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy
 * as 1, 2, 3", SC'11; the same rounds and constants as cuRAND's curand_Philox4x32_10), PINNED by the
 * known-answer vectors in tests/test_philox_oracle.py (zero block; all-ones block; the pi-digits
 * block, cross-checked here against the CUDA toolkit's curand_philox4x32_x.h compiled for the host).
 *
 * Block layout (part of the spec): counter = (action, car_lo32, car_hi32, stream_id),
 * key = (seed_lo32, seed_hi32); words 0,1 -> steer, words 2,3 -> speed (steer is drawn first in the
 * reference); u = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 (numpy's 53-bit double from two 32-bit words);
 * value = lo + (hi - lo) * u with separately rounded product and sum (numpy's uniform()).
 */
#include <stdint.h>

#define ORC_EXPORT __attribute__((visibility("default")))

ORC_EXPORT void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double unit_double(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

/* actions (n_cars, n_actions, 2) = (speed, steer) for global cars car_offset .. car_offset+n_cars-1 */
ORC_EXPORT void orc_rollout_actions(double *actions, int64_t n_cars, int32_t n_actions, uint64_t seed,
                                    uint32_t stream_id, int64_t car_offset, double speed_lo,
                                    double speed_hi, double steer_lo, double steer_hi)
{
    const uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    for (int64_t c = 0; c < n_cars; ++c) {
        const uint64_t gc = (uint64_t)(c + car_offset);
        for (int32_t a = 0; a < n_actions; ++a) {
            const uint32_t ctr[4] = {(uint32_t)a, (uint32_t)(gc & 0xffffffffu), (uint32_t)(gc >> 32), stream_id};
            uint32_t w[4];
            orc_philox4x32_10(ctr, key, w);
            double *o = actions + 2 * (c * n_actions + a);
            const double sr = steer_hi - steer_lo, vr = speed_hi - speed_lo;
            const double ps = sr * unit_double(w[0], w[1]), pv = vr * unit_double(w[2], w[3]);
            o[1] = steer_lo + ps;
            o[0] = speed_lo + pv;
        }
    }
}
