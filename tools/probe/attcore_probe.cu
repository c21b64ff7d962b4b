// Probe: the tcgen05 + TMA attention core (csrc/att_core_tc.cuh) against a CPU computation, scores first (K-major SW128
// descriptors + tensor-map boxes), then the output (MN-major SW128 B operand).  usage: attcore_probe [Tc] [B*H... as B, H=4]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../rtfs_net_b200/csrc/att_core_tc.cuh"
using namespace rtfs;

static float tf32h(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}
int main(int argc, char** argv) {
    const int Tc = argc > 1 ? atoi(argv[1]) : 125, B = argc > 2 ? atoi(argv[2]) : 1, H = 4, BH = B * H;
    const int KP = Tc > 128 ? 256 : 128, QT = (Tc + 127) / 128;
    std::vector<float> q((size_t)BH * Tc * 256), k(q.size()), v((size_t)BH * Tc * 1024), o((size_t)B * Tc * 64 * 64), s((size_t)BH * QT * 128 * KP);
    srand(1);
    auto rnd = [] { return tf32h((float)rand() / RAND_MAX * 2.f - 1.f); };
    for (auto& x : q) x = rnd();
    for (auto& x : k) x = rnd();
    for (auto& x : v) x = rnd();
    float *dq, *dk, *dv, *dout, *ds;
    cudaMalloc(&dq, q.size() * 4); cudaMalloc(&dk, k.size() * 4); cudaMalloc(&dv, v.size() * 4); cudaMalloc(&dout, o.size() * 4); cudaMalloc(&ds, s.size() * 4);
    cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dk, k.data(), k.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dv, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, o.size() * 4);
    cudaMemset(ds, 0, s.size() * 4);
    cudaError_t e = launch_attn_core_tc(dq, dk, dv, dout, B, H, Tc, 0, ds);
    printf("launch: %s\n", cudaGetErrorString(e));
    e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(s.data(), ds, s.size() * 4, cudaMemcpyDeviceToHost);
#ifdef ATC_PROBE_RAW
    for (int i = 0; i < 4; ++i) {
        printf("row %d: sum %g raw O:", i, s[(size_t)i * KP]);
        for (int j = 0; j < 12; ++j) printf(" %g", s[(size_t)i * KP + 1 + j]);
        printf("\n");
    }
#endif
    // scores
    double es = 0, ns = 0, eo = 0, no = 0;
    int shown = 0;
    std::vector<double> p(Tc);
    for (int bh = 0; bh < BH; ++bh)
        for (int i = 0; i < Tc; ++i) {
            double mx = -1e30;
            for (int j = 0; j < Tc; ++j) {
                double acc = 0;
                for (int f = 0; f < 256; ++f) acc += (double)q[((size_t)bh * Tc + i) * 256 + f] * k[((size_t)bh * Tc + j) * 256 + f];
                const double got = s[((size_t)bh * QT * 128 + i) * KP + j];
                es += (got - acc) * (got - acc);
                ns += acc * acc;
                if (fabs(got - acc) > 1e-2 * (1 + fabs(acc)) && shown < 12) { printf("S[bh %d][%d][%d] got %g want %g\n", bh, i, j, got, acc); ++shown; }
                p[j] = acc / 16.0;
                mx = fmax(mx, p[j]);
            }
            double sum = 0;
            for (int j = 0; j < Tc; ++j) { p[j] = exp(p[j] - mx); sum += p[j]; }
            const int b = bh / H, h = bh % H;
            for (int n = 0; n < 1024; ++n) {
                double acc = 0;
                for (int j = 0; j < Tc; ++j) acc += p[j] * v[((size_t)bh * Tc + j) * 1024 + n];
                acc /= sum;
                const double got = o[(((size_t)b * Tc + i) * 64 + (n >> 4)) * 64 + h * 16 + (n & 15)];
                eo += (got - acc) * (got - acc);
                no += acc * acc;
                if (fabs(got - acc) > 1e-2 * (1 + fabs(acc)) && shown < 24) { printf("O[bh %d][%d][%d] got %g want %g\n", bh, i, n, got, acc); ++shown; }
            }
        }
    printf("Tc %d BH %d: scores rel_l2 %.3e   output rel_l2 %.3e\n", Tc, BH, sqrt(es / ns), sqrt(eo / no));
    // timing
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch_attn_core_tc(dq, dk, dv, dout, B, H, Tc, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) launch_attn_core_tc(dq, dk, dv, dout, B, H, Tc, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("time per launch: %.1f us (%s)\n", ms * 50.f, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
