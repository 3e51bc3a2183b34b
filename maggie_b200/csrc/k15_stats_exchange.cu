// K15: SyncBatchNorm-equivalent statistics exchange INSIDE the BatchNorm kernel chain, over peer memory (NVLink / NVSwitch)
// instead of one NCCL all-reduce per layer and direction (engine/train.py:160-161, 71 BatchNorms x 2).
//
//   forward : conv epilogue (MG_CONV_STAT_COPIES statistic copies) -> [K15: reduce copies, push to peers, wait, sum] -> mg_bn_finalize
//   backward: mg_bn_bwd_reduce ([2][C] sums)      -> [K15: push to peers, wait, sum]                -> mg_bn_bwd_apply
//
// Every rank owns one exchange WINDOW (device memory exported with CUDA IPC and mapped by all peers):
//     recv  float    [SLOTS][MAX_RANKS][PAYLOAD]      payload of rank r for exchange number s lands in recv[s % SLOTS][r]
//     flags uint32   [SLOTS][MAX_RANKS][MAX_BLOCKS]   = s + 1 once that payload (the part of block b) is complete
//     seq   uint32   [2]                              local exchange counter + block arrival counter
// Push protocol: a block computes its 32 channels, STORES them into every rank's window (remote stores are posted, one
// NVLink traversal), fences, then releases the flag in every window; it then spins on its LOCAL flags and sums the LOCAL
// copies in rank order - every rank adds the same numbers in the same order, so the statistics are bit-identical
// everywhere.  Exchanges are strictly ordered on each rank's stream and a rank can be at most one exchange ahead of a peer
// (it needs the peer's payload to finish), so a ring of SLOTS = 4 never overwrites unread data; sequence numbers only grow,
// so nothing is ever reset.  The counter lives on the device: the kernel is CUDA-graph capturable.
#include "common.cuh"

#include <cstdlib>

#include <cstring>

namespace {

constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;
constexpr int SLOTS = 4, MAX_RANKS = MG_XCHG_MAX_RANKS, MAX_BLOCKS = 16, MAX_C = 32 * MAX_BLOCKS;
constexpr int PAYLOAD = 2 * MAX_C + 32;                       // floats per (slot, rank): [sum C][sumsq C] ... [count]
constexpr size_t RECV_FLOATS = (size_t)SLOTS * MAX_RANKS * PAYLOAD;
constexpr size_t FLAG_WORDS = (size_t)SLOTS * MAX_RANKS * MAX_BLOCKS;
constexpr size_t WINDOW_BYTES = RECV_FLOATS * 4 + FLAG_WORDS * 4 + 256;

struct XArgs {
    float* win[MAX_RANKS];
    int rank, world;
    const float* in;
    int n_copies, C;
    float count;            // < 0: the payload carries no element count
    float* out;             // [2][C] (+ [1] global count)
    unsigned long long timeout_ns;   // 0: wait for the peers without limit (like a collective)
};

__device__ __forceinline__ float* recv_of(float* win, int slot, int r) { return win + ((size_t)slot * MAX_RANKS + r) * PAYLOAD; }
__device__ __forceinline__ uint32_t* flag_of(float* win, int slot, int r, int b) {
    return reinterpret_cast<uint32_t*>(win + RECV_FLOATS) + ((size_t)slot * MAX_RANKS + r) * MAX_BLOCKS + b;
}
__device__ __forceinline__ uint32_t* seq_of(float* win) { return reinterpret_cast<uint32_t*>(win + RECV_FLOATS) + FLAG_WORDS; }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// grid = ceil(C / 32) blocks of 8 copy-groups x 32 channels (the shape of bn_finalize_kernel)
__global__ void __launch_bounds__(256)
stats_exchange_kernel(const XArgs a) {
    mg::pdl_prologue();
    __shared__ float s_s[8][32], s_q[8][32];
    __shared__ uint32_t s_seq;
    const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl, C = a.C;
    float* mine = a.world > 1 ? a.win[a.rank] : nullptr;
    if (threadIdx.x == 0 && mine) s_seq = *reinterpret_cast<volatile uint32_t*>(seq_of(mine));
    float s = 0.f, q = 0.f;
    if (c < C) {
        for (int k = grp; k < a.n_copies; k += 8) {
            s += __ldg(a.in + (size_t)k * 2 * C + c);
            q += __ldg(a.in + (size_t)k * 2 * C + C + c);
        }
    }
    s_s[grp][cl] = s, s_q[grp][cl] = q;
    __syncthreads();
    const bool owner = grp == 0 && c < C;
    if (grp == 0) {
        s = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_s[k][cl], q += s_q[k][cl];
    }
    if (!mine) {   // one rank: the reduced copy is the result
        if (owner) a.out[c] = s, a.out[C + c] = q;
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.count >= 0.f) a.out[2 * C] = a.count;
        return;
    }
    const uint32_t seq = s_seq, want = seq + 1u;
    const int slot = (int)(seq % SLOTS);
    // every block has read the counter before it arrives; the last arrival advances it for the next exchange
    if (threadIdx.x == 0) {
        uint32_t* sq = seq_of(mine);
        if (atomicAdd(sq + 1, 1u) == gridDim.x - 1) {
            sq[1] = 0u;
            __threadfence();
            sq[0] = want;
        }
    }
    // ---- push my part into every window (my own included)
    if (owner || (blockIdx.x == 0 && threadIdx.x == 0)) {
        for (int p = 0; p < a.world; ++p) {
            float* dst = recv_of(a.win[p], slot, a.rank);
            if (owner) dst[c] = s, dst[MAX_C + c] = q;
            if (blockIdx.x == 0 && threadIdx.x == 0) dst[2 * MAX_C] = a.count;
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x < a.world) st_release_sys(flag_of(a.win[threadIdx.x], slot, a.rank, blockIdx.x), want);
    // ---- wait for every rank's part of this block
    if (threadIdx.x < a.world) {
        const uint32_t* f = flag_of(mine, slot, threadIdx.x, blockIdx.x);
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(f) - want) < 0) {
            if (a.timeout_ns && globaltimer_ns() - t0 > a.timeout_ns) {
                printf("maggie_b200: statistics exchange %u timed out waiting for rank %d (block %d)\n", seq, (int)threadIdx.x,
                       (int)blockIdx.x);
                __trap();
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    // ---- sum in rank order (identical on every rank); the window is read past L1 (slots are reused)
    if (owner) {
        float S = 0.f, Q = 0.f;
        for (int p = 0; p < a.world; ++p) {
            const float* src = recv_of(mine, slot, p);
            S += __ldcg(src + c), Q += __ldcg(src + MAX_C + c);
        }
        a.out[c] = S, a.out[C + c] = Q;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.count >= 0.f) {
        float n = 0.f;
        for (int p = 0; p < a.world; ++p) n += __ldcg(recv_of(mine, slot, p) + 2 * MAX_C);
        a.out[2 * C] = n;
    }
}

}  // namespace

extern "C" size_t mg_xchg_window_bytes(void) { return WINDOW_BYTES; }

extern "C" int mg_xchg_window_create(void** window, void* ipc_handle) {
    MG_REQUIRE(window && ipc_handle, "mg_xchg_window_create: null pointer");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, WINDOW_BYTES);
    if (e == cudaSuccess) e = cudaMemset(p, 0, WINDOW_BYTES);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        (void)cudaGetLastError();
        mg::set_error("mg_xchg_window_create: %s", cudaGetErrorString(e));
        return MG_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == MG_XCHG_HANDLE_BYTES, "IPC handle size");
    std::memcpy(ipc_handle, &h, sizeof(h));
    *window = p;
    return MG_OK;
}

extern "C" int mg_xchg_window_open(const void* ipc_handle, void** mapped) {
    MG_REQUIRE(ipc_handle && mapped, "mg_xchg_window_open: null pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, sizeof(h));
    const cudaError_t e = cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        mg::set_error("mg_xchg_window_open: %s", cudaGetErrorString(e));
        return MG_ERR_CUDA;
    }
    return MG_OK;
}

extern "C" int mg_xchg_window_close(void* mapped) {
    if (mapped && cudaIpcCloseMemHandle(mapped) != cudaSuccess) (void)cudaGetLastError();
    return MG_OK;
}

extern "C" int mg_xchg_window_destroy(void* window) {
    if (window && cudaFree(window) != cudaSuccess) (void)cudaGetLastError();
    return MG_OK;
}

extern "C" int mg_stats_exchange(const mg_xchg_desc* x, const float* in, int n_copies, int C, float count, float* out,
                                 void* stream) {
    MG_REQUIRE(in && out && C > 0 && n_copies > 0, "mg_stats_exchange: null pointer");
    MG_REQUIRE(C <= MAX_C, "mg_stats_exchange: C = %d exceeds %d", C, MAX_C);
    XArgs a;
    std::memset(&a, 0, sizeof(a));
    a.world = 1;
    if (x && x->world > 1) {
        MG_REQUIRE(x->world <= MAX_RANKS && x->rank >= 0 && x->rank < x->world, "mg_stats_exchange: bad rank / world (%d / %d)",
                   x->rank, x->world);
        for (int r = 0; r < x->world; ++r) {
            MG_REQUIRE(x->window[r], "mg_stats_exchange: window of rank %d is not mapped", r);
            a.win[r] = static_cast<float*>(x->window[r]);
        }
        a.rank = x->rank, a.world = x->world;
    }
    a.in = in, a.n_copies = n_copies, a.C = C, a.count = count, a.out = out;
    // A rank may legitimately arrive late (rank-0-only validation between steps, a slow loader worker, a checkpoint save):
    // like NCCL's SyncBatchNorm collective the exchange WAITS.  The limit is a watchdog of NCCL order (torch's process
    // groups abort after 600 s), not a pacing requirement: MAGGIE_B200_XCHG_TIMEOUT_S seconds, 0 = never give up.
    static const unsigned long long timeout_ns = [] {
        const char* e = std::getenv("MAGGIE_B200_XCHG_TIMEOUT_S");
        const double sec = e ? std::atof(e) : 600.0;
        return sec <= 0.0 ? 0ull : (unsigned long long)(sec * 1e9);
    }();
    a.timeout_ns = timeout_ns;
    MG_LAUNCH(stats_exchange_kernel, mg::ceil_div(C, 32), 256, 0, stream, a);
    MG_CHECK_LAUNCH("mg_stats_exchange");
    return MG_OK;
}
