// Host <-> device transfer helpers for the slab upload (NumPy arrays are pageable memory).
//
// cudaMemcpy from pageable memory is staged by the driver through a small pinned buffer on one
// thread (measured 4-9 GB/s on the B200 box for the 2.8 GB slab of config 5).  hp_host_to_device
// pipelines the same thing at memcpy speed: the source is cut into chunks, `nthreads` host threads
// copy chunk i into one half of a caller-provided pinned staging buffer while the DMA engine moves
// chunk i-1 from the other half.  A source that is already page-locked goes out in one async copy.
#include <cstring>
#include <thread>
#include <vector>

#include "hp_common.cuh"

namespace hp {

static void parallel_copy(char* dst, const char* src, size_t bytes, int nthreads) {
    if (nthreads <= 1 || bytes < (size_t(1) << 22)) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const size_t align = 4096;
    size_t per = (bytes / size_t(nthreads) + align - 1) / align * align;
    std::vector<std::thread> pool;
    pool.reserve(size_t(nthreads));
    for (int t = 0; t < nthreads; ++t) {
        const size_t lo = size_t(t) * per;
        if (lo >= bytes) break;
        const size_t n = (lo + per <= bytes) ? per : (bytes - lo);
        pool.emplace_back([=] { std::memcpy(dst + lo, src + lo, n); });
    }
    for (auto& th : pool) th.join();
}

static bool is_page_locked(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();  // clear the sticky "invalid value" older drivers report for pageable memory
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

}  // namespace hp

using namespace hp;

extern "C" int hp_host_is_pinned(const void* host_ptr) { return host_ptr && is_page_locked(host_ptr) ? 1 : 0; }

extern "C" int hp_host_to_device(void* dst_dev, const void* src_host, size_t bytes, void* staging,
                                 size_t staging_bytes, int32_t nthreads, void* stream) {
    HP_REQUIRE(dst_dev && src_host, "null pointer");
    if (bytes == 0) return HP_OK;
    cudaStream_t st = as_stream(stream);
    if (is_page_locked(src_host))
        return check_cuda(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync");
    const size_t half = staging ? staging_bytes / 2 / 4096 * 4096 : 0;
    if (half < (size_t(1) << 20) || !is_page_locked(staging)) {
        // no usable staging buffer: the driver's own pageable path (synchronous)
        int rc = check_cuda(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync");
        if (rc) return rc;
        return check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    }
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (auto& e : ev) {
        int rc = check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        if (rc) {
            if (ev[0]) cudaEventDestroy(ev[0]);  // the second creation failed: do not leak the first
            return rc;
        }
    }
    bool used[2] = {false, false};
    int rc = HP_OK, i = 0;
    const char* src = static_cast<const char*>(src_host);
    char* dst = static_cast<char*>(dst_dev);
    for (size_t off = 0; off < bytes && rc == HP_OK; off += half, ++i) {
        const int b = i & 1;
        const size_t n = (off + half <= bytes) ? half : (bytes - off);
        char* stage = static_cast<char*>(staging) + size_t(b) * half;
        if (used[b]) rc = check_cuda(cudaEventSynchronize(ev[b]), "cudaEventSynchronize");
        if (rc) break;
        parallel_copy(stage, src + off, n, nthreads);
        rc = check_cuda(cudaMemcpyAsync(dst + off, stage, n, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync");
        if (rc) break;
        rc = check_cuda(cudaEventRecord(ev[b], st), "cudaEventRecord");
        used[b] = true;
    }
    // the staging buffer belongs to the caller again on return
    for (int b = 0; b < 2; ++b) {
        if (used[b]) {
            int r2 = check_cuda(cudaEventSynchronize(ev[b]), "cudaEventSynchronize");
            if (rc == HP_OK) rc = r2;
        }
        cudaEventDestroy(ev[b]);
    }
    return rc;
}
