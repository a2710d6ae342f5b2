// Peer-memory plumbing for the fused Ulysses exchange (include/qie.h "qie_peers"): IPC-exportable allocations, mapping of
// the other ranks' buffers, and the stream-ordered all-ranks barrier that separates the QKV -> ATTN -> POST phases.
// The data itself is moved by the GEMM / attention epilogues (peer stores over NVLink), not here.
#include <string.h>

#include "common.cuh"

namespace qie {

__device__ int g_barrier_timeouts = 0;

struct PeerFlags {
    unsigned* p[8];
};

// One thread per peer: publish my arrival in the peer's flag array, then wait for the peer's arrival in mine.
// flags[r][s] = last epoch at which rank s arrived at rank r.  Epochs only grow, so a fast rank that is already one
// barrier ahead still satisfies the `>=` test of a slow one.
__global__ void peer_barrier_kernel(PeerFlags f, int rank, int size, unsigned epoch) {
    const int t = threadIdx.x;
    if (t >= size) return;
    __threadfence_system();   // the peer stores of the kernels before this one are ordered before the flag
    unsigned* theirs = f.p[t] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const unsigned* mine = f.p[rank] + t;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) {   // 2 s: a rank is missing; fail loudly instead of hanging the GPU
            atomicAdd(&g_barrier_timeouts, 1);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

}  // namespace qie

using namespace qie;

extern "C" int qie_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle_out64) {
    QIE_REQUIRE(bytes > 0 && dev_ptr, QIE_EINVAL, "qie_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    QIE_CUDA_OK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess && handle_out64) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e == cudaSuccess) memcpy(handle_out64, &h, 64);
    }
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "qie_peer_alloc");
    }
    *dev_ptr = p;
    return QIE_OK;
}

extern "C" int qie_peer_free(void* dev_ptr) {
    if (dev_ptr) QIE_CUDA_OK(cudaFree(dev_ptr));
    return QIE_OK;
}

extern "C" int qie_peer_open(const unsigned char* handle64, void** dev_ptr) {
    QIE_REQUIRE(handle64 && dev_ptr, QIE_EINVAL, "qie_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    QIE_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QIE_OK;
}

extern "C" int qie_peer_close(void* dev_ptr) {
    if (dev_ptr) QIE_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
    return QIE_OK;
}

extern "C" int qie_peer_barrier(void* const* flags_host, int rank, int size, unsigned epoch, void* stream) {
    QIE_REQUIRE(flags_host && size >= 1 && size <= 8 && rank >= 0 && rank < size && epoch > 0, QIE_EINVAL,
                "qie_peer_barrier: bad argument");
    PeerFlags f{};
    for (int i = 0; i < size; ++i) {
        QIE_REQUIRE(flags_host[i], QIE_EINVAL, "qie_peer_barrier: flags of rank %d are null", i);
        f.p[i] = (unsigned*)flags_host[i];
    }
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, size, epoch);
    QIE_LAUNCH_OK("peer_barrier_kernel");
    return QIE_OK;
}

extern "C" int qie_peer_barrier_timeouts(void) {
    int n = 0;
    if (cudaMemcpyFromSymbol(&n, g_barrier_timeouts, sizeof(int)) != cudaSuccess) return -1;
    return n;
}
