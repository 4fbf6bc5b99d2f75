// Row L of SURVEY.md section 8a: cut-off local grid index (bit-exact, sorted ascending), plus the
// AoS -> SoA split of the grid points.
//
// The inclusion test reproduces cKDTree.query_ball_point(center, radius, p=2) exactly:
//     ((dx*dx + dy*dy) + dz*dz) <= radius*radius
// evaluated with individually rounded FP64 operations (__dmul_rn/__dadd_rn are never contracted
// into FMAs).  Selection is a three-kernel stream compaction (per-block counts -> exclusive scan ->
// ordered scatter), so the output order is the grid order and the result is deterministic.
#include "hp_common.cuh"

namespace hp {

constexpr int kIdxThreads = 256;
constexpr int kIdxPerThread = 8;
constexpr int kIdxSpan = kIdxThreads * kIdxPerThread;

__device__ __forceinline__ double dist2_unfused(const double* __restrict__ pts, int64_t p, double cx,
                                                double cy, double cz) {
    const double dx = __dsub_rn(pts[3 * p + 0], cx);
    const double dy = __dsub_rn(pts[3 * p + 1], cy);
    const double dz = __dsub_rn(pts[3 * p + 2], cz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__global__ void __launch_bounds__(kIdxThreads)
index_count_kernel(const double* __restrict__ pts, int64_t npts, double cx, double cy, double cz,
                   double rad2, int64_t* __restrict__ block_counts) {
    __shared__ int s_cnt[kIdxThreads / 32];
    const int64_t base = int64_t(blockIdx.x) * kIdxSpan + int64_t(threadIdx.x) * kIdxPerThread;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kIdxPerThread; ++j) {
        const int64_t p = base + j;
        if (p < npts) cnt += dist2_unfused(pts, p, cx, cy, cz) <= rad2;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int i = 0; i < kIdxThreads / 32; ++i) total += s_cnt[i];
        block_counts[blockIdx.x] = total;
    }
}

// single-block exclusive scan over the per-block counts (in place), total -> out_count
__global__ void __launch_bounds__(1024)
index_scan_kernel(int64_t* __restrict__ block_counts, int64_t nblocks, int64_t* __restrict__ out_count) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t start = 0; start < nblocks; start += blockDim.x) {
        const int64_t i = start + threadIdx.x;
        const int64_t v = i < nblocks ? block_counts[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t n = __shfl_up_sync(0xffffffffu, incl, off);
            if ((threadIdx.x & 31) >= off) incl += n;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t ws = s_warp[threadIdx.x];
            int64_t wincl = ws;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t n = __shfl_up_sync(0xffffffffu, wincl, off);
                if (threadIdx.x >= off) wincl += n;
            }
            s_warp[threadIdx.x] = wincl - ws;  // exclusive prefix of warp sums
        }
        __syncthreads();
        const int64_t excl = s_carry + s_warp[threadIdx.x >> 5] + incl - v;
        if (i < nblocks) block_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = s_carry;
}

__global__ void __launch_bounds__(kIdxThreads)
index_scatter_kernel(const double* __restrict__ pts, int64_t npts, double cx, double cy, double cz,
                     double rad2, int64_t begin, int64_t end, const int64_t* __restrict__ block_excl,
                     int64_t* __restrict__ out_idx, uint8_t* __restrict__ out_overlap,
                     double* __restrict__ out_dist) {
    __shared__ int s_cnt[kIdxThreads / 32];
    const int64_t base = int64_t(blockIdx.x) * kIdxSpan + int64_t(threadIdx.x) * kIdxPerThread;
    double d2[kIdxPerThread];
    unsigned mask = 0;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kIdxPerThread; ++j) {
        const int64_t p = base + j;
        d2[j] = 0.0;
        if (p < npts) {
            d2[j] = dist2_unfused(pts, p, cx, cy, cz);
            if (d2[j] <= rad2) {
                mask |= 1u << j;
                ++cnt;
            }
        }
    }
    // exclusive prefix of cnt over the block, in thread order
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, off);
        if ((threadIdx.x & 31) >= off) incl += n;
    }
    if ((threadIdx.x & 31) == 31) s_cnt[threadIdx.x >> 5] = incl;
    __syncthreads();
    int warp_excl = 0;
    for (int i = 0; i < (threadIdx.x >> 5); ++i) warp_excl += s_cnt[i];
    int64_t pos = block_excl[blockIdx.x] + warp_excl + incl - cnt;
#pragma unroll
    for (int j = 0; j < kIdxPerThread; ++j) {
        if (mask & (1u << j)) {
            const int64_t p = base + j;
            out_idx[pos] = p;
            if (out_overlap) out_overlap[pos] = (begin <= p) && (p < end);
            // np.linalg.norm(local.points - center, axis=1): sqrt of the same unfused sum
            if (out_dist) out_dist[pos] = __dsqrt_rn(d2[j]);
            ++pos;
        }
    }
}

__global__ void __launch_bounds__(256)
split_points_kernel(const double* __restrict__ pts, int64_t npts, double* __restrict__ px,
                    double* __restrict__ py, double* __restrict__ pz) {
    __shared__ double tile[3 * 256];
    const int64_t base = int64_t(blockIdx.x) * 256;
    const int64_t n = min(int64_t(256), npts - base);
    for (int i = threadIdx.x; i < 3 * n; i += 256) tile[i] = pts[3 * base + i];
    __syncthreads();
    if (threadIdx.x < n) {
        px[base + threadIdx.x] = tile[3 * threadIdx.x + 0];
        py[base + threadIdx.x] = tile[3 * threadIdx.x + 1];
        pz[base + threadIdx.x] = tile[3 * threadIdx.x + 2];
    }
}

// Neighbour counts for the sharding work estimate (core/device.py estimate_dense_work): for atom a
// and threshold c,  counts[a][c] = #{ b : |R_a - R_b|^2 <= radii2[kind_a][kind_b][c] }.
// One block per atom, atoms b strided over the threads; nrad <= kMaxCountRadii.
constexpr int kMaxCountRadii = 128;

__global__ void __launch_bounds__(128)
neighbor_counts_kernel(int natom, const double* __restrict__ xyz, const int* __restrict__ kind, int nkind,
                       int nrad, const double* __restrict__ radii2, double* __restrict__ counts) {
    __shared__ int s_cnt[kMaxCountRadii];
    const int a = blockIdx.x;
    for (int c = threadIdx.x; c < nrad; c += blockDim.x) s_cnt[c] = 0;
    __syncthreads();
    const double ax = xyz[3 * a], ay = xyz[3 * a + 1], az = xyz[3 * a + 2];
    const double* row = radii2 + size_t(kind[a]) * nkind * nrad;
    for (int b = threadIdx.x; b < natom; b += blockDim.x) {
        const double dx = xyz[3 * b] - ax, dy = xyz[3 * b + 1] - ay, dz = xyz[3 * b + 2] - az;
        const double d2 = dx * dx + dy * dy + dz * dz;
        const double* r2 = row + size_t(kind[b]) * nrad;
        for (int c = 0; c < nrad; ++c)
            if (d2 <= r2[c]) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nrad; c += blockDim.x) counts[size_t(a) * nrad + c] = double(s_cnt[c]);
}

}  // namespace hp

using namespace hp;

extern "C" size_t hp_local_index_scratch_bytes(int64_t npts) {
    const int64_t nblocks = (npts + kIdxSpan - 1) / kIdxSpan;
    return sizeof(int64_t) * size_t(nblocks > 0 ? nblocks : 1);
}

extern "C" int hp_build_local_index(const double* points_xyz, int64_t npts,
                                    const double* center_host, double radius, int64_t begin,
                                    int64_t end, int64_t* out_indices, uint8_t* out_overlap,
                                    double* out_dist, int64_t* out_count, void* scratch,
                                    size_t scratch_bytes, void* stream) {
    HP_REQUIRE(npts >= 0 && center_host && out_count, "bad arguments");
    cudaStream_t st = as_stream(stream);
    if (npts == 0) return check_cuda(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st), "memset");
    HP_REQUIRE(points_xyz && out_indices && scratch, "null buffer");
    HP_REQUIRE(scratch_bytes >= hp_local_index_scratch_bytes(npts), "scratch too small");
    HP_REQUIRE(radius >= 0.0, "negative radius");
    const int64_t nblocks = (npts + kIdxSpan - 1) / kIdxSpan;
    HP_REQUIRE(nblocks < (int64_t(1) << 31), "grid too large for one launch");
    // cKDTree compares squared distances against radius*radius (rounded product)
    const double rad2 = radius * radius;
    const double cx = center_host[0], cy = center_host[1], cz = center_host[2];
    int64_t* counts = static_cast<int64_t*>(scratch);
    index_count_kernel<<<int(nblocks), kIdxThreads, 0, st>>>(points_xyz, npts, cx, cy, cz, rad2, counts);
    HP_LAUNCH_CHECK("index_count_kernel");
    index_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, out_count);
    HP_LAUNCH_CHECK("index_scan_kernel");
    index_scatter_kernel<<<int(nblocks), kIdxThreads, 0, st>>>(points_xyz, npts, cx, cy, cz, rad2, begin,
                                                               end, counts, out_indices, out_overlap,
                                                               out_dist);
    HP_LAUNCH_CHECK("index_scatter_kernel");
    return HP_OK;
}

extern "C" int hp_split_points(const double* points_xyz, int64_t npts, double* px, double* py,
                               double* pz, void* stream) {
    HP_REQUIRE(npts >= 0, "bad size");
    if (npts == 0) return HP_OK;
    HP_REQUIRE(points_xyz && px && py && pz, "null buffer");
    const int64_t blocks = (npts + 255) / 256;
    HP_REQUIRE(blocks < (int64_t(1) << 31), "grid too large for one launch");
    split_points_kernel<<<int(blocks), 256, 0, as_stream(stream)>>>(points_xyz, npts, px, py, pz);
    HP_LAUNCH_CHECK("split_points_kernel");
    return HP_OK;
}

extern "C" int hp_neighbor_counts(int32_t natom, const double* atom_xyz, const int32_t* kind, int32_t nkind,
                                  int32_t nrad, const double* radii2, double* counts, void* stream) {
    HP_REQUIRE(natom >= 0 && nkind > 0 && nrad > 0, "bad sizes");
    HP_REQUIRE(nrad <= kMaxCountRadii, "too many radii (max 128)");
    if (natom == 0) return HP_OK;
    HP_REQUIRE(atom_xyz && kind && radii2 && counts, "null buffer");
    neighbor_counts_kernel<<<natom, 128, 0, as_stream(stream)>>>(natom, atom_xyz, kind, nkind, nrad, radii2, counts);
    HP_LAUNCH_CHECK("neighbor_counts_kernel");
    return HP_OK;
}
