// Device-resident outer loop: the whole `while change >= threshold and counter < maxiter` of
// do_partitioning (core/iterstock.py:171-188) as ONE CUDA-graph launch.
//
// Small systems are launch- and sync-bound: one stockholder iteration of H2O is seven kernels of a few
// microseconds each, and the host loop pays a D2H copy plus a stream synchronisation per iteration just
// to read `change`.  Here the iteration's launches are captured once into the body of a conditional
// WHILE node (CUDA 12.4+); the last kernel of the body (`loop_commit_kernel`) appends the iteration's
// state vector, change and entropy to a device-side history, advances the iteration counter and sets
// the node's condition with the reference's own stopping rule, so the GPU runs iterations back to back
// with no host involvement and the host reads the whole history with one copy at the end.
// `niter`, every history entry and the final weights are identical to the host-driven loop: the same
// kernels run in the same order on the same buffers.
#include "hp_common.cuh"

namespace hp {

struct Loop {
    cudaGraph_t graph = nullptr;
    cudaGraph_t body = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphConditionalHandle handle = 0;
    bool capturing = false;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// stamps[2 * row + slot] = now;  row = *counter + row_shift (the iteration being executed is *counter)
__global__ void loop_stamp_kernel(const int* __restrict__ counter, int row_shift, int slot,
                                  unsigned long long* __restrict__ stamps) {
    stamps[2 * (*counter + row_shift) + slot] = global_timer_ns();
}

// history row c = [state vector (nvec) | change | entropy];  continue iff not (change < threshold) and
// c + 1 < maxiter (core/iterstock.py:187 `if change < self._threshold or counter >= self._maxiter: break`)
__global__ void __launch_bounds__(256)
loop_commit_kernel(cudaGraphConditionalHandle handle, int nvec, const double* __restrict__ vec,
                   const double* __restrict__ out2, double threshold, int maxiter,
                   double* __restrict__ history, int* __restrict__ counter,
                   unsigned long long* __restrict__ stamps) {
    const int c = *counter;
    double* row = history + size_t(c) * (nvec + 2);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) row[i] = vec[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        const double change = out2[0];
        row[nvec] = change;
        row[nvec + 1] = out2[1];
        if (stamps) stamps[2 * c + 1] = global_timer_ns();
        *counter = c + 1;
        const bool stop = (change < threshold) || (c + 1 >= maxiter);
        cudaGraphSetConditional(handle, stop ? 0u : 1u);
    }
}

}  // namespace hp

using namespace hp;

extern "C" int hp_loop_begin(void* stream, void** loop_out) {
    HP_REQUIRE(loop_out, "null output");
    Loop* L = new Loop();
    int rc = check_cuda(cudaGraphCreate(&L->graph, 0), "cudaGraphCreate");
    if (rc == HP_OK)
        rc = check_cuda(cudaGraphConditionalHandleCreate(&L->handle, L->graph, 1, cudaGraphCondAssignDefault),
                        "cudaGraphConditionalHandleCreate");
    cudaGraphNode_t node = nullptr;
    if (rc == HP_OK) {
        cudaGraphNodeParams p = {};
        p.type = cudaGraphNodeTypeConditional;
        p.conditional.handle = L->handle;
        p.conditional.type = cudaGraphCondTypeWhile;
        p.conditional.size = 1;
        rc = check_cuda(cudaGraphAddNode(&node, L->graph, nullptr, 0, &p), "cudaGraphAddNode(conditional while)");
        if (rc == HP_OK) L->body = p.conditional.phGraph_out[0];
    }
    if (rc == HP_OK)
        rc = check_cuda(cudaStreamBeginCaptureToGraph(as_stream(stream), L->body, nullptr, nullptr, 0,
                                                      cudaStreamCaptureModeRelaxed),
                        "cudaStreamBeginCaptureToGraph");
    if (rc != HP_OK) {
        if (L->graph) cudaGraphDestroy(L->graph);
        delete L;
        return rc;
    }
    L->capturing = true;
    *loop_out = L;
    return HP_OK;
}

extern "C" int hp_loop_stamp(void* loop, const int32_t* counter, int32_t row_shift, int32_t slot,
                             uint64_t* stamps, void* stream) {
    HP_REQUIRE(counter && stamps && (slot == 0 || slot == 1), "bad arguments");
    (void)loop;
    loop_stamp_kernel<<<1, 1, 0, as_stream(stream)>>>(counter, row_shift, slot,
                                                      reinterpret_cast<unsigned long long*>(stamps));
    HP_LAUNCH_CHECK("loop_stamp_kernel");
    return HP_OK;
}

extern "C" int hp_loop_commit(void* loop, int32_t nvec, const double* state_vec, const double* out2,
                              double threshold, int32_t maxiter, double* history, int32_t* counter,
                              uint64_t* stamps, void* stream) {
    Loop* L = static_cast<Loop*>(loop);
    HP_REQUIRE(L && L->capturing, "hp_loop_commit outside hp_loop_begin / hp_loop_end");
    HP_REQUIRE(nvec > 0 && state_vec && out2 && history && counter && maxiter > 0, "bad arguments");
    loop_commit_kernel<<<1, 256, 0, as_stream(stream)>>>(L->handle, nvec, state_vec, out2, threshold, maxiter,
                                                         history, counter,
                                                         reinterpret_cast<unsigned long long*>(stamps));
    HP_LAUNCH_CHECK("loop_commit_kernel");
    return HP_OK;
}

extern "C" int hp_loop_end(void* loop, void* stream) {
    Loop* L = static_cast<Loop*>(loop);
    HP_REQUIRE(L && L->capturing, "no capture in progress");
    cudaGraph_t captured = nullptr;
    L->capturing = false;
    int rc = check_cuda(cudaStreamEndCapture(as_stream(stream), &captured), "cudaStreamEndCapture");
    if (rc != HP_OK) return rc;
    return check_cuda(cudaGraphInstantiate(&L->exec, L->graph, 0), "cudaGraphInstantiate");
}

extern "C" int hp_loop_launch(void* loop, void* stream) {
    Loop* L = static_cast<Loop*>(loop);
    HP_REQUIRE(L && L->exec, "loop graph not instantiated");
    return check_cuda(cudaGraphLaunch(L->exec, as_stream(stream)), "cudaGraphLaunch");
}

extern "C" int hp_loop_destroy(void* loop, void* stream) {
    Loop* L = static_cast<Loop*>(loop);
    if (!L) return HP_OK;
    if (L->capturing) {  // abandon an unfinished capture so that the stream is usable again
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(as_stream(stream), &g);
        cudaGetLastError();
    }
    if (L->exec) cudaGraphExecDestroy(L->exec);
    if (L->graph) cudaGraphDestroy(L->graph);
    delete L;
    return HP_OK;
}
