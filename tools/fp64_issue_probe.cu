// FP64 issue-rate probe (not part of the library): how close to the DFMA peak can realistic
// instruction streams get on one B200 SM?  Variants:
//   A  8 chains/thread, multiplier and addend loop-invariant registers shared by all chains
//   B  8 chains/thread, three distinct register operands per DFMA (b_i, c_i per chain)
//   C  4 chains/thread, three distinct register operands
//   D  as C plus one integer instruction per 3 DFMA (the exp/sqrt bit manipulation ratio)
//   E  4 chains/thread, Horner-like: a_i = fma(a_i, r_i, K_k) with K_k loop-invariant registers
// Each for 16, 12 and 8 resident warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_issue_probe.cu -o /tmp/issue && /tmp/issue
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void probe(int iters, double* sink, const double* in) {
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; ++i) {
        a[i] = in[i] + threadIdx.x * 1e-9;
        b[i] = in[8 + i];
        c[i] = in[16 + i];
    }
    const double m = in[24], k0 = in[25], k1 = in[26], k2 = in[27], k3 = in[28];
    int t = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (V == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, k0);
            } else if (V == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b[i], c[i]);
            } else if (V == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = fma(a[i], b[i], c[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = fma(a[i], b[i], c[i]);
            } else if (V == 3) {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = fma(a[i], b[i], c[i]);
                t = t * 3 + 1;
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = fma(a[i], b[i], c[i]);
                t ^= t >> 3;
                t += 7;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = fma(a[i], b[i], k0);
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = fma(a[i], b[i], k1);
            }
        }
        if (V == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = fma(a[i], b[i], k2) + k3;
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456 || t == 123456789) sink[0] = s;
}

template <int V>
void run(const char* name, int threads, int blocks_per_sm, double* sink, double* in) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<V><<<sms * blocks_per_sm, threads>>>(iters / 10, sink, in);
    cudaEventRecord(e0);
    probe<V><<<sms * blocks_per_sm, threads>>>(iters, sink, in);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (V == 4 ? 64.0 + 4.0 : 64.0) * iters * double(threads) * blocks_per_sm * sms;
    const double extra = (V == 4) ? 4.0 * iters * double(threads) * blocks_per_sm * sms : 0.0;  // the DADDs
    const double rate = (dfma + extra) / (ms * 1e-3);                  // FP64 instructions (lanes) per second
    const double peak = sms * 64.0 * 1.965e9;                          // lanes per second at 1.965 GHz
    printf("%-44s warps/SM %2d: %.3f of the nominal FP64 issue rate (%.1f TFLOP/s as FMA)\n", name,
           threads * blocks_per_sm / 32, rate / peak, 2 * rate / 1e12);
}

int main() {
    double *sink, *in;
    cudaMalloc(&sink, 8);
    cudaMalloc(&in, 32 * 8);
    double h[32];
    for (int i = 0; i < 32; ++i) h[i] = 0.999999 + 1e-7 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int cfg[3][2] = {{256, 2}, {128, 3}, {128, 2}};
    for (auto& c : cfg) {
        run<0>("A 8 chains, shared invariant operands", c[0], c[1], sink, in);
        run<1>("B 8 chains, 3 distinct register operands", c[0], c[1], sink, in);
        run<2>("C 4 chains, 3 distinct register operands", c[0], c[1], sink, in);
        run<3>("D 4 chains + 1 integer op per 3 DFMA", c[0], c[1], sink, in);
        run<4>("E 4 chains, Horner-like (invariant addend)", c[0], c[1], sink, in);
    }
    return 0;
}
