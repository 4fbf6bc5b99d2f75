// Accuracy probe for horton_part_b200/csrc/hp_math.cuh (not part of the library).  Run on the B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I horton_part_b200/csrc tools/fp64_probe.cu -o /tmp/fp64probe && /tmp/fp64probe
// sqrt: max deviation from IEEE sqrt on 6.7e7 log-uniform samples of d2.
// exp:  max deviation from a long-double (64-bit mantissa) host reference on 2^20 samples in [-708, 0].
#include <cmath>
#include <cstdio>
#include <vector>

#include "hp_math.cuh"

using namespace hp;

__global__ void probe_sqrt(int n, double* out) {
    double m0 = 0, m1 = 0, m2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double u = (i + 0.5) / n;
        const double d2 = exp(-40.0 + 110.0 * u) * (1.0 + 0.37 * u);
        const double s = sqrt(d2);
        m0 = fmax(m0, fabs(rsqrt_seed(d2) * s - 1.0));
        m1 = fmax(m1, fabs(sqrt_nocall(d2) - s) / s);
        m2 = fmax(m2, fabs(sqrt_fast(d2) - s) / s);
    }
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(m2));
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(m0));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(m1));
}

// e_tab slot: the unguarded variant exp_neg_poly<false> (valid for -700 < x <= 0)
__global__ void probe_exp(int n, const double* x, double* e_poly, double* e_tab, double* e_lib) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    e_poly[i] = exp_neg_poly(x[i]);
    e_tab[i] = x[i] > -700.0 ? exp_neg_poly<false>(x[i]) : exp_neg_poly(x[i]);
    e_lib[i] = exp(x[i]);
}

int main() {
    double* d;
    cudaMalloc(&d, 3 * sizeof(double));
    cudaMemset(d, 0, 3 * sizeof(double));
    probe_sqrt<<<592, 256>>>(1 << 26, d);
    double h[3];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("rsqrt seed rel err = %.3e (2^%.1f)\n", h[0], log2(h[0]));
    printf("sqrt_nocall max rel deviation from IEEE sqrt = %.3e (%.3f ulp)\n", h[1], h[1] / ldexp(1.0, -52));
    printf("sqrt_fast   max rel deviation from IEEE sqrt = %.3e (%.3f ulp)\n", h[2], h[2] / ldexp(1.0, -52));

    const int n = 1 << 20;
    std::vector<double> x(n), a(n), b(n), c(n);
    for (int i = 0; i < n; ++i) {
        const double u = (i + 0.5) / n;
        x[i] = (i % 3 == 0) ? -708.0 * u : ((i % 3 == 1) ? -40.0 * u * u : -u * 1e-3);
    }
    x[0] = 0.0; x[1] = -0.0; x[2] = -707.999; x[3] = -708.0; x[4] = -1e6; x[5] = -1e300;
    double *dx, *da, *db, *dc;
    cudaMalloc(&dx, n * 8); cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dc, n * 8);
    cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice);
    probe_exp<<<n / 256, 256>>>(n, dx, da, db, dc);
    cudaMemcpy(a.data(), da, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), db, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(c.data(), dc, n * 8, cudaMemcpyDeviceToHost);
    double ea = 0, eb = 0, ec = 0;
    for (int i = 6; i < n; ++i) {
        const long double ref = expl((long double)x[i]);
        const double ulp = (double)ref * ldexp(1.0, -52);
        ea = fmax(ea, fabs((double)((long double)a[i] - ref)) / ulp);
        eb = fmax(eb, fabs((double)((long double)b[i] - ref)) / ulp);
        ec = fmax(ec, fabs((double)((long double)c[i] - ref)) / ulp);
    }
    printf("exp max error vs long double: poly %.3f ulp, unguarded poly %.3f ulp, CUDA exp() %.3f ulp\n", ea, eb, ec);
    printf("edge cases (poly | table | lib): exp(0)=%.17g|%.17g|%.17g exp(-0)=%g|%g exp(-707.999)=%g|%g|%g exp(-708)=%g|%g|%g exp(-1e6)=%g|%g exp(-1e300)=%g|%g\n",
           a[0], b[0], c[0], a[1], b[1], a[2], b[2], c[2], a[3], b[3], c[3], a[4], b[4], a[5], b[5]);
    return 0;
}
