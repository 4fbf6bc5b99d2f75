// Which part of the per-pair instruction stream keeps the promolecule kernel at ~80 % of the FP64
// issue rate?  Builds the stream up step by step on NP independent points per thread (registers only,
// no memory traffic) and reports FP64 instructions per SM-cycle relative to the nominal 2.0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I horton_part_b200/csrc -I include tools/fp64_mix_probe.cu -o /tmp/mix && /tmp/mix
#include <cstdio>
#include <cuda_runtime.h>

#include "hp_math.cuh"
using namespace hp;

// STAGE 1: Horner only (11 DFMA)   2: + range reduction & exponent splice (16)   3: + sqrt_fast (21)
//       4: + distance (27)         5: + A*e and the two accumulating adds (30) = the kernel's fast path
template <int STAGE, int NP, int MINB = 1>
__global__ void __launch_bounds__(128 * (MINB > 1 ? 1 : 2), MINB) mix(int iters, const double* in, double* sink) {
    double x[NP], y[NP], z[NP], pro[NP];
    for (int j = 0; j < NP; ++j) {
        x[j] = in[j] + threadIdx.x * 1e-3;
        y[j] = in[8 + j];
        z[j] = in[16 + j];
        pro[j] = 0.0;
    }
    double ax = in[24], ay = in[25], az = in[26], A = in[27], na = -in[28], off = in[29];
    ExpConsts ec;
    ec.load();
    Exp2Consts ec2;
    ec2.load();
    const double nb = na * ec2.log2e;
    for (int it = 0; it < iters; ++it) {
        double f[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            double v;
            if (STAGE >= 4) {
                const double dx = x[j] - ax, dy = y[j] - ay, dz = z[j] - az;
                v = fma(dz, dz, fma(dy, dy, dx * dx));
            } else {
                v = x[j] + ax;  // one DADD; keeps the stage from being hoisted out of the loop
            }
            if (STAGE == 3 || STAGE >= 4) v = sqrt_fast(v);
            if (STAGE == 6) {  // sqrt surrogate without MUFU / integer ops: 5 dependent FP64 ops
                double g = v * 0.7;
                g = fma(fma(-g, g, v), 0.3, g);
                v = fma(fma(-g, g, v), 0.3, g);
            }
            if (STAGE == 8) {  // stage 5 with the base-2 exponential (15 instead of 16 FP64 ops)
                f[j] = exp2_neg_poly_regs(nb * v, ec2);
            } else if (STAGE >= 2 && STAGE != 7) {
                f[j] = exp_neg_poly_regs(na * v, ec);
            } else {
                if (STAGE == 7) v = v * 0.01;
                double g = fma(c_exp_regs[9], v, c_exp_regs[8]);
#pragma unroll
                for (int i = 7; i >= 0; --i) g = fma(g, v, c_exp_regs[i]);
                g = fma(g, v, 1.0);
                f[j] = fma(g, v, 1.0);
            }
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            if (STAGE >= 5) pro[j] = (pro[j] + A * f[j]) + off;
            else pro[j] += f[j];
        }
        ax += 1e-7;  // next "atom"
    }
    double s = 0;
    for (int j = 0; j < NP; ++j) s += pro[j];
    if (s == 123.456) sink[0] = s;
}

template <int STAGE, int NP, int MINB = 1>
void run(int threads, int bps, const double* in, double* sink) {
    // FP64 instructions per point and iteration (counted from the source; +1 accumulate)
    const int fp64_per_point[9] = {0, 13, 18, 23, 28, 31, 23, 18, 30};
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 40000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mix<STAGE, NP, MINB>, threads, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, mix<STAGE, NP, MINB>);
    if (occ < bps) printf("(occupancy %d blocks/SM < %d requested, %d regs) ", occ, bps, fa.numRegs);
    mix<STAGE, NP, MINB><<<sms * bps, threads>>>(iters / 10, in, sink);
    cudaEventRecord(e0);
    mix<STAGE, NP, MINB><<<sms * bps, threads>>>(iters, in, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = double(fp64_per_point[STAGE]) * NP * iters * (threads / 32) * bps;  // per SM
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("stage %d, %d points/thread, %2d warps/SM (%d chains/SMSP, %d regs): %.3f FP64 instr/cycle/SM = %.1f %% of 2.0\n", STAGE, NP,
           threads * bps / 32, NP * threads * bps / 128, fa.numRegs, warp_instr / cycles, 50.0 * warp_instr / cycles);
}

int main() {
    double h[32], *in, *sink;
    for (int i = 0; i < 32; ++i) h[i] = 0.37 + 0.011 * i;
    h[27] = 1.3; h[28] = 1.9; h[29] = 1e-100;
    cudaMalloc(&in, sizeof(h));
    cudaMalloc(&sink, 8);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    // stage 6 = stage 3 with the MUFU-seeded sqrt replaced by 5 plain FP64 ops; stage 7 = sqrt_fast + Horner only
    run<1, 8>(128, 2, in, sink); run<2, 8>(128, 2, in, sink); run<3, 8>(128, 2, in, sink);
    run<6, 8>(128, 2, in, sink); run<7, 8>(128, 2, in, sink); run<4, 8>(128, 2, in, sink); run<5, 8>(128, 2, in, sink);
    run<1, 4>(128, 4, in, sink); run<2, 4>(128, 4, in, sink); run<3, 4>(128, 4, in, sink);
    run<6, 4>(128, 4, in, sink); run<7, 4>(128, 4, in, sink); run<4, 4>(128, 4, in, sink); run<5, 4>(128, 4, in, sink);
    run<5, 4>(128, 3, in, sink); run<5, 8>(256, 1, in, sink);
    run<8, 8>(128, 2, in, sink); run<8, 4>(128, 4, in, sink); run<5, 8>(128, 1, in, sink); run<5, 8>(64, 2, in, sink);
    // more independent chains in flight per scheduler (the FP64 dependent-issue latency is ~36 cycles)
    run<5, 8, 3>(128, 3, in, sink); run<5, 6, 4>(128, 4, in, sink); run<5, 5, 4>(128, 4, in, sink); run<5, 4, 6>(128, 6, in, sink);
    run<8, 8, 3>(128, 3, in, sink); run<8, 6, 4>(128, 4, in, sink); run<5, 3, 8>(128, 8, in, sink); run<1, 8, 4>(128, 4, in, sink);
    return 0;
}
