// All-reduce of the fixed-point deposition grids over the GPUs of one node through peer memory (NVLink / NVSwitch), fused with the fold of
// the per-CTA grid copies: what replaces `k_fold_acc` + ncclAllReduce on the hot path of a residual evaluation.
//
// The grids are 64-bit integers (ipic_common.cuh), so the sum is the same whatever the order: every rank adds the ranks' contributions
// itself and all ranks end up with identical bits -- no broadcast, no reduction tree to pin.
//
//   k_fold_publish   each rank: copy 0 + copies 1.. of its grids -> its window buffer [parity][count] (copies 1.. zeroed), then one
//                    system-scope flag store per peer: flags[my rank] = sequence number of this reduction.
//   k_pull_reduce    each rank: waits until the flags of all ranks have reached the sequence number, then reads every rank's window
//                    buffer (its own included) over NVLink, slice by slice across many CTAs, and writes the total into copy 0 of its own
//                    grids, where the field kernel collects it.
//
// The window is double-buffered by the parity of the sequence number, so one flag per rank is the only synchronisation: a rank rewrites
// buffer p two reductions later, after its own pull of the reduction in between, which waited for flags that the peers only raise after
// they have finished reading buffer p (stream order on the peer).  All ranks run the same sequence of reductions (the nonlinear solve is
// replicated and takes identical decisions on identical bits), so sequence numbers agree without being exchanged.
//
// A rank that stops raising its flag (crashed peer) must not hang the others: the wait gives up after IPIC_PEER_TIMEOUT_CYCLES, sets an
// error word the host checks at its next synchronisation, and lets the kernel finish.
#include "ipic_common.cuh"
#include <cstring>

namespace ipic {

#ifndef IPIC_PEER_TIMEOUT_CYCLES
#define IPIC_PEER_TIMEOUT_CYCLES 6000000000ll      // ~3 s at 1.965 GHz
#endif

constexpr int PEER_MAX = 16;

struct PeerArgs {
    unsigned long long *acc; size_t count; int reps; size_t rep_stride;
    unsigned long long *mine;                  // this rank's window buffer of the current parity
    const unsigned long long *theirs[PEER_MAX];   // every rank's window buffer of the current parity (own entry = mine)
    unsigned long long *flags_at[PEER_MAX];    // flags array inside every rank's window (remote stores)
    const unsigned long long *my_flags;        // flags array of this rank (peers store into it)
    unsigned *ticket;                          // CTA completion counter of k_fold_publish (self-resetting)
    int *error;                                // device error word: 1 = a peer did not show up
    unsigned long long seq;
    int nranks, rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) k_fold_publish(const PeerArgs A) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.count) {
        unsigned long long q = A.acc[i];
        for (int r = 1; r < A.reps; ++r) { unsigned long long *p = A.acc + (size_t)r * A.rep_stride + i; q += *p; *p = 0ull; }
        A.mine[i] = q;
    }
    // the last CTA to finish raises this rank's flag on every rank
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(A.ticket, 1u);
        last = (t == gridDim.x - 1);
        if (last) *A.ticket = 0u;
    }
    __syncthreads();
    if (last && threadIdx.x < A.nranks) {
        __threadfence_system();
        st_release_sys(A.flags_at[threadIdx.x] + A.rank, A.seq);
    }
}

__global__ void __launch_bounds__(256) k_pull_reduce(const PeerArgs A) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < A.nranks) {
        const long long t0 = clock64();
        while (ld_acquire_sys(A.my_flags + threadIdx.x) < A.seq) {
            if (clock64() - t0 > IPIC_PEER_TIMEOUT_CYCLES) { ok = 0; *A.error = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (!ok) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.count; i += stride) {
        unsigned long long q = 0ull;
#pragma unroll 8
        for (int r = 0; r < A.nranks; ++r) q += ld_relaxed_sys(A.theirs[r] + i);
        A.acc[i] = q;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
// window layout: [2][count_max] uint64 buffers, then PEER_MAX flags, then ticket + error words
static size_t window_words(const Ctx &c) { return 2 * (size_t)NSP_MAX * 2 * ((size_t)std::max(c.Nx, c.Nh) + 2) + PEER_MAX + 8; }

int peer_window_create(Ctx &c) {
    if (c.peer_window) return IPIC_OK;
    const size_t bytes = window_words(c) * sizeof(unsigned long long);
    IPIC_CUDA(cudaMalloc(&c.peer_window, bytes));
    IPIC_CUDA(cudaMemset(c.peer_window, 0, bytes));
    IPIC_CUDA(cudaDeviceSynchronize());
    return IPIC_OK;
}

int peer_export(Ctx &c, void *handle64) {
    int rc = peer_window_create(c); if (rc) return rc;
    cudaIpcMemHandle_t hnd;
    IPIC_CUDA(cudaIpcGetMemHandle(&hnd, c.peer_window));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle64, &hnd, 64);
    return IPIC_OK;
}

int peer_import(Ctx &c, int nranks, int rank, const void *handles) {
    if (nranks < 1 || nranks > PEER_MAX || rank < 0 || rank >= nranks) { set_error("peer import: 1..16 ranks"); return IPIC_ERR_ARG; }
    int rc = peer_window_create(c); if (rc) return rc;
    c.peer_ptr.assign(nranks, nullptr);
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { c.peer_ptr[r] = c.peer_window; continue; }
        cudaIpcMemHandle_t hnd;
        std::memcpy(&hnd, static_cast<const char *>(handles) + (size_t)64 * r, 64);
        void *p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error(std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
            c.peer_ptr.clear();
            return IPIC_ERR_CUDA;
        }
        c.peer_ptr[r] = p;
    }
    c.nranks = nranks; c.rank = rank;
    c.peer_ready = true;
    return IPIC_OK;
}

void peer_free(Ctx &c) {
    for (int r = 0; r < (int)c.peer_ptr.size(); ++r)
        if (r != c.rank && c.peer_ptr[r]) cudaIpcCloseMemHandle(c.peer_ptr[r]);
    c.peer_ptr.clear();
    if (c.peer_window) cudaFree(c.peer_window);
    c.peer_window = nullptr; c.peer_ready = false;
}

// all-reduce (sum) of the first `count` words of the deposition grids (all CTA copies folded) over the ranks; result in copy 0
int peer_allreduce_acc(Ctx &c, size_t count) {
    const size_t half = (size_t)NSP_MAX * 2 * ((size_t)std::max(c.Nx, c.Nh) + 2);
    if (count > half) { set_error("peer all-reduce: grid larger than the window"); return IPIC_ERR_ARG; }
    unsigned long long *W = static_cast<unsigned long long *>(c.peer_window);
    PeerArgs A{};
    A.acc = c.d_acc; A.count = count; A.reps = std::max(1, c.acc_reps); A.rep_stride = c.acc_rep_stride;
    A.seq = ++c.peer_seq;
    const size_t par = (size_t)(A.seq & 1ull) * half;
    A.mine = W + par;
    for (int r = 0; r < c.nranks; ++r) {
        unsigned long long *Wr = static_cast<unsigned long long *>(c.peer_ptr[r]);
        A.theirs[r] = Wr + par;
        A.flags_at[r] = Wr + 2 * half;
    }
    A.my_flags = W + 2 * half;
    A.ticket = reinterpret_cast<unsigned *>(W + 2 * half + PEER_MAX);
    A.error = reinterpret_cast<int *>(W + 2 * half + PEER_MAX + 1);
    A.nranks = c.nranks; A.rank = c.rank;
    const unsigned fold_blocks = (unsigned)((count + 255) / 256);
    c.launches += 2;
    k_fold_publish<<<fold_blocks, 256, 0, c.stream>>>(A);
    k_pull_reduce<<<std::min(fold_blocks, 64u), 256, 0, c.stream>>>(A);
    IPIC_CUDA(cudaGetLastError());
    return IPIC_OK;
}

// the device error word of the window (a peer that never raised its flag): read where the host synchronises anyway
int peer_check(Ctx &c) {
    if (!c.peer_ready) return IPIC_OK;
    const size_t half = (size_t)NSP_MAX * 2 * ((size_t)std::max(c.Nx, c.Nh) + 2);
    int err = 0;
    IPIC_CUDA(cudaMemcpyAsync(&err, static_cast<unsigned long long *>(c.peer_window) + 2 * half + PEER_MAX + 1, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    if (err) { set_error("peer all-reduce: a rank did not raise its flag in time"); return IPIC_ERR_NCCL; }
    return IPIC_OK;
}

}  // namespace ipic
