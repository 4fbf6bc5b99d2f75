// FP64 tensor-core (DMMA) throughput on the box, register operands only -- the yardstick SURVEY.md section 8d
// (unit U2) asks for before choosing between DFMA and mma.sync for the gLISA Hessian contraction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dmma_probe.cu -o /tmp/dmma && /tmp/dmma
// Each warp keeps NACC independent accumulator fragments and issues mma.sync back to back; reported: TFLOP/s
// (2 flop per multiply-add) for m8n8k4 and m16n8k16, next to a plain DFMA loop with the same accumulators.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) dmma_884(int iters, const double* in, double* out) {
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_16816(int iters, const double* in, double* out) {
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = in[(threadIdx.x + i) & 63];
    for (int i = 0; i < 4; ++i) b[i] = in[(threadIdx.x + 7 * i) & 63];
    double c[NACC][4];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
                  "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(int iters, const double* in, double* out) {
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c[NACC];
    for (int i = 0; i < NACC; ++i) c[i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, b, c[i]);
        a += 1e-9;
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

template <typename K>
void run(const char* name, K kernel, double flop_per_warp_iter, int warps_per_sm, const double* in, double* out) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000, threads = 256, blocks = sms * warps_per_sm / 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kernel<<<blocks, threads>>>(iters / 10, in, out);
    cudaEventRecord(e0);
    kernel<<<blocks, threads>>>(iters, in, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = flop_per_warp_iter * iters * double(blocks) * (threads / 32);
    printf("%-34s %2d warps/SM: %7.2f TFLOP/s  (%s)\n", name, warps_per_sm, flop / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double h[64], *in, *out;
    for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-3 * i;
    cudaMalloc(&in, sizeof(h));
    cudaMalloc(&out, 8);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {8, 16, 32}) {
        run("DFMA, 16 accumulators/thread", dfma_loop<16>, 16.0 * 32 * 2, w, in, out);
        run("DMMA m8n8k4, 8 accumulators/warp", dmma_884<8>, 8.0 * 8 * 8 * 4 * 2, w, in, out);
        run("DMMA m8n8k4, 16 accumulators/warp", dmma_884<16>, 16.0 * 8 * 8 * 4 * 2, w, in, out);
        run("DMMA m16n8k16, 4 accumulators/warp", dmma_16816<4>, 4.0 * 16 * 8 * 16 * 2, w, in, out);
        run("DMMA m16n8k16, 8 accumulators/warp", dmma_16816<8>, 8.0 * 16 * 8 * 16 * 2, w, in, out);
    }
    return 0;
}
