/*
 * oracle/followgap_oracle.c -- CPU restatement of the reference's follow-the-gap action generator
 * (followgap/followgap.hpp:5-130; called on every MCTS expansion, scripts/mcts.py:262-267).
 * TEST INFRASTRUCTURE ONLY.  PINNED: oracle/Makefile also compiles the unmodified header into
 * oracle/_ref/libfollowgap_ref.so and tests/test_followgap_oracle.py compares the two bit for bit.
 *
 * One deliberate difference: the reference reads lidar[best_point] with best_point == size when the
 * widest gap is the single last beam (followgap.hpp:121-125, out of bounds); here and in the CUDA
 * kernel best_point is clamped to size-1.
 */
#include <math.h>
#include <stdlib.h>

#define ORC_EXPORT __attribute__((visibility("default")))

ORC_EXPORT float orc_followgap_eval(const float *lidar, int size, float max_distance, float max_angle,
                                    float angle_inc)
{
    float *v = (float *)malloc(sizeof(float) * size);
    for (int i = 0; i < size; ++i) v[i] = lidar[i];
    /* preprocessLidar: clip all but the last 10 beams (:18-28) */
    for (int i = 0; i < size - 10; ++i)
        if (v[i] > max_distance) v[i] = max_distance;
    /* closest non-zero return (:111-118) */
    int min_point = 0;
    for (int i = 0; i < size; ++i)
        if (v[i] != 0 && v[i] < v[min_point]) min_point = i;
    /* safetyBubble(v, min_point, 5) (:66-79) */
    v[min_point] = 0.0f;
    for (int i = -5; i < 5; ++i)
        if (min_point + i > 0 && min_point + i < size - 1) v[min_point + i] = 0.0f;
    /* findMaxGap: first longest run of beams farther than 1.75 m (:30-64) */
    int max_start = 0, max_size = 0, c = 0;
    while (c < size) {
        int start = c, len = 0;
        while (c < size && v[c] > 1.75) { ++len; ++c; }
        if (len > max_size) { max_start = start; max_size = len; }
        ++c;
    }
    int best = (max_start + (max_start + max_size + 1)) / 2;   /* findBestPoint (:99-102) */
    if (best > size - 1) best = size - 1;
    /* getSteerAng(lidar[best], size, best) (:81-97): the two branches are the same expression */
    float angle;
    if (best > size / 2) angle = (float)(-angle_inc * ((size / 2.0) - best));
    else angle = (float)(angle_inc * (best - (size / 2.0)));
    angle = 2 * (angle / lidar[best]);
    free(v);
    float lo = fmaxf(angle, -max_angle);
    if (angle != angle) lo = angle;             /* std::max(NaN, x) keeps NaN */
    float hi = (max_angle < lo) ? max_angle : lo;
    return hi;
}
