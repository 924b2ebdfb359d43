// probe_chain.cu -- what one march step costs a warp that is alone on its SM (the drain of a launch).
// One lane marches a ray that creeps at the 1 px minimum step through a field of ones for 300 steps, with the
// product's march_ray (csrc/march.cuh): first pass = cold L1 (the look-ahead touches of the tail mode have to do
// the work), second pass of the same ray = every sample an L1 hit (the floor of the dependent chain
// FFMA -> F2I -> IMAD -> IMAD.WIDE -> LDG -> FADD -> FSETP -> BRA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I../pyracecarsimulator_b200/csrc -o probe_chain probe_chain.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "march.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

namespace rl {
void set_error(const std::string &) {}
int32_t fail(int32_t code, const std::string &) { return code; }
}

__global__ void chain_kernel(rl::MarchParams P, float x0, float y0, float dx, float dy, int lanes, long long *cycles, unsigned *steps, float *out)
{
    if ((int)threadIdx.x >= lanes) return;
    // lanes > 1: neighbouring beams, a fraction of a degree apart
    const float a = 1e-3f * threadIdx.x;
    const float ddx = dx * cosf(a) - dy * sinf(a), ddy = dx * sinf(a) + dy * cosf(a);
    for (int pass = 0; pass < 3; ++pass) {
        uint32_t n = 0;
        const rl::FirstSample f0 = rl::first_sample(P, x0, y0);
        const long long t0 = clock64();
        const float r = rl::march_ray<true, true, true>(P, x0, y0, ddx, ddy, n, f0);
        const long long t1 = clock64();
        if (threadIdx.x == 0) { cycles[pass] = t1 - t0; steps[pass] = n; out[pass] = r; }
    }
}

// The same chain with every load forced to L2 (ld.global.cg): what a step costs when the sample misses L1.
__global__ void chain_l2_kernel(rl::MarchParams P, float x0, float y0, float dx, float dy, long long *cycles, unsigned *steps)
{
    float t = 1.0f;
    unsigned n = 0;
    const long long t0 = clock64();
    for (;;) {
        const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
        float s;
        asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(s) : "l"(P.dist + (px * P.stride + py)));
        ++n;
        t = __fadd_rn(t, s);
        if (!(t < P.max_range)) break;
    }
    const long long t1 = clock64();
    cycles[0] = t1 - t0;
    steps[0] = n;
}

int main()
{
    const int rows = 1024, cols = 1024, pad = 320;
    const int stride = cols + 2 * pad, prow = rows + 2 * pad;
    std::vector<float> h((size_t)prow * stride, nanf(""));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) h[(size_t)(r + pad) * stride + c + pad] = 1.0f;
    float *d;
    CK(cudaMalloc(&d, h.size() * sizeof(float)));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    rl::MarchParams P{};
    P.dist = d + (size_t)pad * stride + pad;
    P.rows = rows; P.cols = cols; P.stride = stride; P.pad = pad;
    P.frows = rows; P.fcols = cols; P.max_range = 300.0f;
    {   // offset / stride by multiply-high (march.cu: make_fast_div)
        int sh = 1;
        while ((1u << sh) < (unsigned)stride) ++sh;
        P.stride_magic = (uint32_t)((((uint64_t)1 << (31 + sh)) + stride - 1) / stride);
        P.stride_shift = (uint32_t)(sh - 1);
    }
    long long *cyc; unsigned *st; float *out;
    CK(cudaMallocManaged(&cyc, 4 * sizeof(long long)));
    CK(cudaMallocManaged(&st, 4 * sizeof(unsigned)));
    CK(cudaMallocManaged(&out, 4 * sizeof(float)));
    const struct { const char *name; float deg; } dirs[] = {{"along a row (fast axis)", 90.f}, {"across rows (slow axis)", 0.f}, {"diagonal", 45.f}, {"10 degrees off the slow axis", 10.f}};
    for (auto &dd : dirs) {
        const float a = dd.deg * 3.14159265f / 180.f;
        for (int lanes : {1, 32}) {
            chain_kernel<<<1, 32>>>(P, 512.3f, 512.7f, cosf(a), sinf(a), lanes, cyc, st, out);
            CK(cudaDeviceSynchronize());
            printf("{\"probe\": \"chain\", \"direction\": \"%s\", \"lanes\": %d, \"steps\": %u, \"cycles_per_step_cold_l1\": %.1f, \"cycles_per_step_second_pass\": %.1f, \"cycles_per_step_third_pass\": %.1f}\n",
                   dd.name, lanes, st[0], (double)cyc[0] / st[0], (double)cyc[1] / st[1], (double)cyc[2] / st[2]);
        }
        chain_l2_kernel<<<1, 1>>>(P, 512.3f, 512.7f, cosf(a), sinf(a), cyc, st);
        CK(cudaDeviceSynchronize());
        printf("{\"probe\": \"chain\", \"direction\": \"%s\", \"loads\": \"ld.global.cg (every sample from L2)\", \"steps\": %u, \"cycles_per_step\": %.1f}\n", dd.name, st[0], (double)cyc[0] / st[0]);
    }
    return 0;
}
