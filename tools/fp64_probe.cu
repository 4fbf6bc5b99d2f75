// Scratch probe (not part of the library): accuracy of MUFU.RSQ64H and of candidate no-call sqrt
// sequences against IEEE sqrt, on a log-uniform sample of d2.  Run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_probe.cu -o /tmp/probe && /tmp/probe
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

__device__ double sqrt_a(double d2) {  // library-like: cubic + Heron (8 ops)
    double y = seed(d2);
    const double e = fma(d2, -(y * y), 1.0);
    const double p = fma(e, 0.375, 0.5);
    y = fma(p, y * e, y);
    const double g = d2 * y;
    const double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
    return fma(fma(-g, g, d2), h, g);
}
__device__ double sqrt_b(double d2) {  // Goldschmidt x2 (6 ops)
    const double y = seed(d2);
    double g = d2 * y;
    double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    return fma(g, r, g);
}
__device__ double sqrt_c(double d2) {  // Goldschmidt + Heron (6 ops)
    const double y = seed(d2);
    double g = d2 * y;
    double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, d2), h, g);
}

__global__ void probe(int n, double* out) {
    // out[0..3]: max rel err of seed, a, b, c (in units of 2^-53)
    double m0 = 0, ma = 0, mb = 0, mc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double u = (i + 0.5) / n;
        const double d2 = exp(-40.0 + 110.0 * u) * (1.0 + 0.37 * u);
        const double s = sqrt(d2);
        const double inv = 1.0 / s;
        m0 = fmax(m0, fabs(seed(d2) * s - 1.0));
        ma = fmax(ma, fabs(sqrt_a(d2) - s) * inv);
        mb = fmax(mb, fabs(sqrt_b(d2) - s) * inv);
        mc = fmax(mc, fabs(sqrt_c(d2) - s) * inv);
    }
    // crude reduction through atomics on bit patterns (all values positive)
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(m0));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(ma));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(mb));
    atomicMax((unsigned long long*)&out[3], (unsigned long long)__double_as_longlong(mc));
}

int main() {
    double* d; cudaMalloc(&d, 4 * sizeof(double)); cudaMemset(d, 0, 4 * sizeof(double));
    probe<<<592, 256>>>(1 << 26, d);
    double h[4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double ulp = ldexp(1.0, -53);
    printf("seed rel err = %.3e (2^%.1f)\n", h[0], log2(h[0]));
    printf("sqrt_a (cubic+Heron, 8 ops)      max err = %.3f ulp\n", h[1] / ulp / 2);
    printf("sqrt_b (Goldschmidt x2, 6 ops)    max err = %.3f ulp\n", h[2] / ulp / 2);
    printf("sqrt_c (Goldschmidt+Heron, 6 ops) max err = %.3f ulp\n", h[3] / ulp / 2);
    return 0;
}
