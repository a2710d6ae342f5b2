// Peer-memory plumbing for the fused Ulysses exchange (include/qie.h "qie_peers"): IPC-exportable allocations, mapping of
// the other ranks' buffers, and the stream-ordered all-ranks barrier that separates the QKV -> ATTN -> POST phases.
// The data itself is moved by the GEMM / attention epilogues (peer stores over NVLink), not here.
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace qie {

// Sticky failure flag of the peer barriers: one word of mapped, pinned host memory per process, so that the host can read it
// without synchronising any stream.  A barrier that gives up (a rank is missing or late by > 2 s) raises it; every later
// qie_forward* call of this process then returns QIE_ECUDA instead of computing on half-written buffers.
static volatile int* g_sticky_host = nullptr;
static int* g_sticky_dev = nullptr;
static std::mutex g_sticky_mu;

static int sticky_init() {
    std::lock_guard<std::mutex> lk(g_sticky_mu);
    if (g_sticky_host) return QIE_OK;
    int* p = nullptr;
    QIE_CUDA_OK(cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(p, 0, 64);
    QIE_CUDA_OK(cudaHostGetDevicePointer(&g_sticky_dev, p, 0));
    g_sticky_host = p;
    return QIE_OK;
}

int peer_sticky_error(bool clear) {
    if (!g_sticky_host || *g_sticky_host == 0) return QIE_OK;
    const int n = *g_sticky_host;
    if (clear) *g_sticky_host = 0;
    (void)n;
    set_error("a peer barrier of the sequence-parallel group timed out: a rank is missing or more than 2 s late; everything "
              "computed after it is invalid");
    return QIE_ECUDA;
}

struct PeerFlags {
    unsigned* p[8];
};

// One thread per peer: publish my arrival in the peer's flag array, then wait for the peer's arrival in mine.
// flags[r][s] = last epoch at which rank s arrived at rank r; flags[r][8] = rank r's own barrier count (device-side, so the
// kernel can be replayed from a CUDA graph: every rank executes the same sequence of barriers, the counts stay in step).
// Epochs only grow, so a fast rank that is already one barrier ahead still satisfies the `>=` test of a slow one.
__global__ void peer_barrier_kernel(PeerFlags f, int rank, int size, int* sticky) {
    __shared__ unsigned epoch_s;
    const int t = threadIdx.x;
    // programmatic dependent launch: this one-warp kernel is resident before its predecessor has drained and lets its successor
    // run its prologue; the flag is only published once the predecessor has completed and flushed its (peer) stores
    griddep_launch_dependents();
    griddep_wait();
    if (t == 0) {
        unsigned* ctr = f.p[rank] + 8;
        epoch_s = *ctr + 1;
        *ctr = epoch_s;
    }
    __syncthreads();
    if (t >= size) return;
    const unsigned epoch = epoch_s;
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
            *reinterpret_cast<volatile int*>(sticky) = 1;   // plain store: host-mapped memory has no portable atomics
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

// my [B][img_rows][C] velocity rows (compact [B][img_pad][C] source) -> rows [img_offset, img_offset + img_rows) of every rank's
// [B][img_total][C] velocity buffer: 16-byte peer stores, one (row, 16 B) element per thread per peer
__global__ void peer_bcast_rows_kernel(const uint4* __restrict__ src, void* const* __restrict__ vel, int size, int batch,
                                       int img_rows, int img_pad, int img_total, int img_offset, int c16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)batch * img_rows * c16) return;
    const int c = (int)(i % c16);
    const long long r = i / c16;
    const int row = (int)(r % img_rows), b = (int)(r / img_rows);
    const uint4 v = src[((long long)b * img_pad + row) * c16 + c];
    for (int g = 0; g < size; ++g) {
        uint4* dst = reinterpret_cast<uint4*>(__ldg(reinterpret_cast<const unsigned long long*>(vel) + g));
        dst[((long long)b * img_total + img_offset + row) * c16 + c] = v;
    }
}

int peer_bcast_rows(const void* src, void* const* peer_vel_dev, const qie_peers* pr, int img_rows, int img_offset, int C,
                    cudaStream_t st) {
    QIE_REQUIRE(src && peer_vel_dev && pr && C % 8 == 0, QIE_EINVAL, "peer_bcast_rows: bad argument");
    const long long total = (long long)pr->batch * img_rows * (C / 8);
    peer_bcast_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint4*)src, peer_vel_dev, pr->size, pr->batch,
                                                                            img_rows, pr->img_pad, pr->img_total, img_offset, C / 8);
    QIE_LAUNCH_OK("peer_bcast_rows_kernel");
    return QIE_OK;
}

// my span [off16, off16 + len16) (in 16-byte units) of every batch row of a table that all ranks hold at the same layout ->
// the same span of every OTHER rank's copy (the modulation table: each rank computes 1/P of its rows)
__global__ void peer_bcast_span_kernel(const uint4* __restrict__ src, void* const* __restrict__ tab, int size, int rank, int batch,
                                       long long bstride16, long long off16, long long len16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * len16) return;
    const long long idx = (i / len16) * bstride16 + off16 + i % len16;
    const uint4 v = src[idx];
    for (int g = 0; g < size; ++g)
        if (g != rank) reinterpret_cast<uint4*>(__ldg(reinterpret_cast<const unsigned long long*>(tab) + g))[idx] = v;
}

int peer_bcast_span(const void* mine, void* const* peer_tab_dev, const qie_peers* pr, long long bstride_bytes, long long off_bytes,
                    long long len_bytes, cudaStream_t st) {
    QIE_REQUIRE(mine && peer_tab_dev && pr && bstride_bytes % 16 == 0 && off_bytes % 16 == 0 && len_bytes % 16 == 0, QIE_EINVAL,
                "peer_bcast_span: bad argument");
    const long long total = pr->batch * (len_bytes / 16);
    if (total == 0) return QIE_OK;
    peer_bcast_span_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint4*)mine, peer_tab_dev, pr->size, pr->rank,
                                                                            pr->batch, bstride_bytes / 16, off_bytes / 16, len_bytes / 16);
    QIE_LAUNCH_OK("peer_bcast_span_kernel");
    return QIE_OK;
}

int peer_barrier_launch(const qie_peers* pr, cudaStream_t st) {
    int rc = sticky_init();
    if (rc) return rc;
    PeerFlags f{};
    for (int i = 0; i < pr->size; ++i) f.p[i] = (unsigned*)pr->flags[i];
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = launch_attrs(attr, 1);
    QIE_CUDA_OK(cudaLaunchKernelEx(&cfg, peer_barrier_kernel, f, pr->rank, pr->size, g_sticky_dev));
    QIE_LAUNCH_OK("peer_barrier_kernel");
    return QIE_OK;
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

// stream-ordered device-to-device copy out of / into a qie_peer_alloc buffer (tests, debugging: the buffers are raw cudaMalloc
// allocations without a torch tensor around them)
extern "C" int qie_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
    QIE_REQUIRE(dst && src, QIE_EINVAL, "qie_peer_copy: null pointer");
    QIE_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
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

// number of peer-barrier timeouts of this process so far (sticky; read without synchronising; see peer_sticky_error)
extern "C" int qie_peer_barrier_timeouts(void) { return g_sticky_host ? *g_sticky_host : 0; }
