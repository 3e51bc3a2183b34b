// K9b: persistent variant of the rulebook convolution (K9) for the 32-channel layers on the large site lists (OS1 /
// OS2: 100 k .. 350 k rows).  The one-tile-per-CTA kernel spends most of a CTA's ~13 us lifetime on things that do not
// scale with the work - staging the weights, TMEM allocation, barrier set-up, one dependent index -> row gather latency
// chain, a tcgen05.commit round trip, the epilogue of four warps - and 2700 CTAs pay them one after the other.  Here:
//   * one CTA per SM walks over the 128-row tiles; the weight pack is staged once per CTA;
//   * four producer warps issue ALL neighbour gathers of a tile (T taps x 128 rows x 64 B) as zero-filling 16-byte
//     cp.async into one of two tile buffers (the next tile is gathered while the current one is multiplied / stored);
//   * one thread issues the T x 2 tcgen05.mma of a tile and ONE commit per tile (accumulators double buffered in TMEM);
//   * two epilogue warpgroups take alternate tiles (bias, ReLU, BatchNorm1d partial sums kept per warp over all tiles
//     and flushed once, fp16 rows at a column offset).
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

using namespace mg::ptx;

constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;
constexpr int PRODUCERS = 128;
constexpr int EPI_GROUPS = 2;
constexpr int EPI_WARPS = 4 * EPI_GROUPS;
constexpr int THREADS = PRODUCERS + 32 + 32 * EPI_WARPS;   // warps 0..3 gather, warp 4 MMA, warps 5..12 epilogue
constexpr int ROWB = 64;                                   // 32 fp16 channels per source row

struct PArgs {
    const __half* src; int src_stride;
    const int32_t* table; int T, No, Cout;             // T = VIRTUAL taps = table taps x (Cin / 32); Cout multiple of 16, <= 64
    int Tt, cl;                                         // table taps, log2(Cin / 32): virtual tap vt = (tap vt >> cl, channels (vt & mask) * 32 ..)
    const __half* w;                                    // [Cout][T*32]
    const float* bias;
    __half* out; int out_stride, c_off;
    float* stats;
    float* map; const int32_t* coords; int mapH, mapW, Cout_real;   // head mode: fp32 logit map, column 0 only
    int pre_act, n_tiles, b_tile, acc_cols, tmem_cols;
};

__device__ __forceinline__ uint32_t swz64(int r, int j) { return (uint32_t)(j ^ ((r >> 1) & 3)); }
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(THREADS, 1)
sparse_conv_persistent_kernel(const PArgs a) {
    mg::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_tile = 128 * ROWB, a_buf = a.T * a_tile;
    uint8_t* sB = smem;                                   // [T] weight sub-tiles: Cout rows of 64 B, swizzled
    uint8_t* sA = sB + (size_t)a.T * a.b_tile;            // [2][T] gathered row tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * (size_t)a_buf);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* s_stage = reinterpret_cast<float*>(tmem_slot + 4);   // [EPI_WARPS][16][36]
    float* s_part = s_stage + EPI_WARPS * 16 * 36;              // [EPI_WARPS][2][Cout]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t a_full = bar0, a_empty = bar0 + 16, t_full = bar0 + 32, t_empty = bar0 + 48;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + 8 * i, PRODUCERS);
            mbar_init(a_empty + 8 * i, 1);
            mbar_init(t_full + 8 * i, 1);
            mbar_init(t_empty + 8 * i, 4);
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    for (int i = tid; i < EPI_WARPS * 2 * a.Cout; i += THREADS) s_part[i] = 0.f;
    {
        // resident weights: sub-tile t = W[:, t*32 : t*32+32] as Cout rows of 64 bytes, swizzled (asynchronous copies)
        const int Ktot = a.T * 32, chunks = a.T * a.Cout * 4;
        for (int i = tid; i < chunks; i += THREADS) {
            const int j = i & 3, n = (i >> 2) % a.Cout, t = i / (4 * a.Cout);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB + (size_t)t * a.b_tile) + n * ROWB + swz64(n, j) * 16),
                         "l"(a.w + (size_t)n * Ktot + t * 32 + j * 8) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===== gather producers: thread = (16-byte chunk `sub` of a row, rows rr, rr + 32, rr + 64, rr + 96) =====
        const int sub = tid & 3, rr = tid >> 2;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1, ph = (it >> 1) & 1;
            const int tile0 = tile * 128;
            int idx[9][4];
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int p = tile0 + rr + i * 32;
                    idx[t][i] = -1;
                    if (t < a.T && p < a.No) idx[t][i] = a.table ? __ldg(a.table + (size_t)p * a.Tt + (t >> a.cl)) : p;
                }
            mbar_wait(a_empty + 8 * buf, ph ^ 1);
            const uint32_t base = smem_u32(sA + (size_t)buf * a_buf);
            const int cmask = (1 << a.cl) - 1;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                if (t < a.T) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = rr + i * 32;
                        const __half* srcp = idx[t][i] >= 0 ? a.src + (size_t)idx[t][i] * a.src_stride + (t & cmask) * 32 + sub * 8 : a.src;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + t * a_tile + r * ROWB + swz64(r, sub) * 16),
                                     "l"(srcp), "r"(idx[t][i] >= 0 ? 16u : 0u) : "memory");
                    }
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(a_full + 8 * buf);
        }
    } else if (warp == 4) {
        // ===== MMA issuer: T taps x 2 K steps per tile, one commit per tile.  Whole warp with warp-uniform operands (no
        // elect / R2UR waterfall per tcgen05.mma), one elected lane issues =====
        const uint32_t tmem_u = uniform_u32(tmem_base);
        const uint32_t idesc = instr_desc_f16(128, a.Cout, 0, 0);
        const uint32_t lay = swizzle_layout(ROWB), sbo = 8 * ROWB;
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        int buf = 0, ph = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            mbar_wait(t_empty + 8 * buf, ph ^ 1);
            mbar_wait(a_full + 8 * buf, ph);
            tc_fence_after();
            const uint32_t abase = sA_u + buf * a_buf, d_tmem = tmem_u + buf * a.acc_cols;
            if (elect_one()) {
                for (int t = 0; t < a.T; ++t) {
                    const uint64_t da = smem_desc(abase + t * a_tile, 0, sbo, lay), db = smem_desc(sB_u + t * a.b_tile, 0, sbo, lay);
                    mma_f16(d_tmem, da, db, idesc, t != 0);
                    mma_f16(d_tmem, da + 2, db + 2, idesc, 1u);
                }
                mma_commit(t_full + 8 * buf);
                mma_commit(a_empty + 8 * buf);
            }
            if (buf) ph ^= 1;
            buf ^= 1;
        }
    } else {
        // ===== epilogue: warps 5..12; warpgroup g takes the tiles with (it & 1) == g =====
        const int q = warp & 3, ew = warp - 5, grp = ew >> 2;
        const uint32_t stg = smem_u32(s_stage + ew * 16 * 36);
        float* part = s_part + ew * 2 * a.Cout;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != grp) continue;
            const int ph = (it >> 1) & 1;
            const int row = tile * 128 + q * 32 + lane;
            const bool valid = row < a.No;
            mbar_wait(t_full + 8 * grp, ph);
            tc_fence_after();
            for (int c0 = 0; c0 < a.Cout; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + grp * a.acc_cols + c0, r);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = __uint_as_float(r[i]);
                    if (a.bias && c0 + i < a.Cout_real) v[i] += __ldg(a.bias + c0 + i);
                    if (a.pre_act == 1) v[i] = fmaxf(v[i], 0.f);
                }
                if (a.map) {
                    // the two 32 -> 1 heads: the reference computes dense() - 99 and then += 99 at the active sites
                    if (valid && c0 == 0) {
                        const int s = a.coords[row * 3], y = a.coords[row * 3 + 1], x = a.coords[row * 3 + 2];
                        a.map[((size_t)s * a.mapH + y) * a.mapW + x] = (v[0] - 99.0f) + 99.0f;
                    }
                    continue;
                }
                if (a.stats) {
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 16; ++i) sts_f32(stg + (i * 36 + lane) * 4, valid ? v[i] : 0.f);
                    __syncwarp();
                    const int col = lane & 15;
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 z = lds_f32x4(stg + (col * 36 + 4 * j) * 4);
                        acc += lane < 16 ? (z.x + z.y) + (z.z + z.w) : (z.x * z.x + z.y * z.y) + (z.z * z.z + z.w * z.w);
                    }
                    part[(lane >> 4) * a.Cout + c0 + col] += acc;
                }
                if (valid) {
                    uint32_t o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                        o[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    __half* orow = a.out + (size_t)row * a.out_stride + a.c_off + c0;
                    reinterpret_cast<uint4*>(orow)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    reinterpret_cast<uint4*>(orow)[1] = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + 8 * grp);
        }
        if (a.stats) {
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            const int et = tid - (PRODUCERS + 32);
            float* dst = a.stats + (size_t)(blockIdx.x % STAT_COPIES) * 2 * a.Cout;
            for (int i = et; i < 2 * a.Cout; i += 32 * EPI_WARPS) {
                float tot = 0.f;
#pragma unroll
                for (int w = 0; w < EPI_WARPS; ++w) tot += s_part[w * 2 * a.Cout + i];
                atomicAdd(dst + i, tot);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

}  // namespace

namespace mg {

// Launches the persistent kernel when the layer qualifies; *handled tells the caller (mg_sparse_conv) whether it did.
int sparse_conv_persistent_launch(const mg_sparse_conv_desc* d, void* stream, bool* handled) {
    *handled = false;
    const char* e = std::getenv("MAGGIE_B200_NO_PERSISTENT_SPARSE");
    if (e && e[0] == '1') return MG_OK;
    const int cout_pad = (d->Cout + 15) / 16 * 16;     // the caller packs the weights with rows padded to 16
    // 32 or 64 input channels: a 64-channel row is two "virtual taps" of 32 channels (the weight pack is already laid out
    // that way: K index = tap * Cin + ci)
    if ((d->Cin != 32 && d->Cin != 64) || cout_pad > 64 || (!d->map && d->Cout % 16) || (d->map && d->stats)) return MG_OK;
    const int cl = d->Cin == 64 ? 1 : 0, VT = d->T << cl;
    if (VT > 9) return MG_OK;
    const int n_tiles = ceil_div(d->No, 128);
    if (n_tiles < 2 * kNumSMs) return MG_OK;              // small site lists: the one-tile-per-CTA kernel has more parallelism
    PArgs a;
    a.src = static_cast<const __half*>(d->src), a.src_stride = d->src_stride;
    a.table = d->table, a.T = VT, a.Tt = d->T, a.cl = cl, a.No = d->No, a.Cout = cout_pad, a.Cout_real = d->Cout;
    a.map = d->map, a.coords = d->coords, a.mapH = d->mapH, a.mapW = d->mapW;
    a.w = static_cast<const __half*>(d->w), a.bias = d->bias;
    a.out = static_cast<__half*>(d->out), a.out_stride = d->out_stride, a.c_off = d->c_off;
    a.stats = d->stats, a.pre_act = d->pre_act, a.n_tiles = n_tiles;
    a.b_tile = ((cout_pad * ROWB + 1023) / 1024) * 1024;
    a.acc_cols = cout_pad < 32 ? 32 : cout_pad;
    a.tmem_cols = 2 * a.acc_cols;                          // 64 or 128
    const size_t smem = 1024 + (size_t)VT * a.b_tile + 2 * (size_t)VT * 128 * ROWB + 256 + EPI_WARPS * 16 * 36 * 4 +
                        EPI_WARPS * 2 * cout_pad * 4;
    if (smem > 224 * 1024) return MG_OK;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(sparse_conv_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) {
            set_error("mg_sparse_conv: cannot raise dynamic shared memory limit (persistent kernel)");
            return MG_ERR_CUDA;
        }
        attr_set = true;
    }
    MG_LAUNCH(sparse_conv_persistent_kernel, std::min(n_tiles, kNumSMs), THREADS, smem, stream, a);
    MG_CHECK_LAUNCH("mg_sparse_conv(persistent)");
    *handled = true;
    return MG_OK;
}

}  // namespace mg

// ================================================================================================ K9c: weight gradient
// Persistent weight gradient of the same layers:  dW[co][t*32 + ci] = sum_p dout[p][co] * src[table[p][t]][ci].
// GEMM with the sites as K: both operand tiles are MN-major rows of 64 bytes (32 channels).  One CTA per SM accumulates
// its share of the site tiles for ALL taps in TMEM ([128 x T*32] fp32, M padded with zero atoms) and flushes ONCE with
// 16-byte vector reductions; the gathers of the next tile are in flight while the current one is multiplied.
namespace {

struct GArgs {
    const __half* dout; int dout_stride, Cout, cout_eff;   // cout_eff = Cout rounded up to 32 (columns are readable)
    const __half* src; int src_stride;
    const int32_t* table; int T, No, n_tiles;
    float* dw;                                             // [Cout][T*32]
    int tmem_cols;
    int Tt, cl;          // table taps, log2(Cin / 32): T counts VIRTUAL taps (see PArgs)
};

constexpr int G_THREADS = PRODUCERS + 32;   // warps 0..3 gather (+ epilogue at the end), warp 4 MMA

__global__ void __launch_bounds__(G_THREADS, 1)
sparse_wgrad_persistent_kernel(const GArgs a) {
    mg::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int atom = 128 * ROWB, a_buf = 4 * atom, b_buf = a.T * atom;   // A: 4 atoms of 32 channels (M = 128), B: T tiles
    uint8_t* sA = smem;                         // [2][4 atoms]; atoms beyond cout_eff stay zero
    uint8_t* sB = sA + 2 * (size_t)a_buf;       // [2][T]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * (size_t)b_buf);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full = bar0, empty = bar0 + 16, tfull = bar0 + 32;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) mbar_init(full + 8 * i, PRODUCERS), mbar_init(empty + 8 * i, 1);
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    for (int i = tid; i < 2 * a_buf / 16; i += G_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp < 4) {
        const int sub = tid & 3, rr = tid >> 2, real_atoms = a.cout_eff / 32;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1, ph = (it >> 1) & 1;
            const int tile0 = tile * 128;
            int idx[9][4];
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int p = tile0 + rr + i * 32;
                    idx[t][i] = -1;
                    if (t < a.T && p < a.No) idx[t][i] = a.table ? __ldg(a.table + (size_t)p * a.Tt + (t >> a.cl)) : p;
                }
            mbar_wait(empty + 8 * buf, ph ^ 1);
            const int cmask = (1 << a.cl) - 1;
            // A: d_out rows, one 32-channel atom per 64 bytes of a row
            const uint32_t abase = smem_u32(sA + (size_t)buf * a_buf);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rr + i * 32, p = tile0 + r;
                for (int j = 0; j < real_atoms; ++j) {
                    const bool ok = p < a.No;
                    const __half* srcp = ok ? a.dout + (size_t)p * a.dout_stride + j * 32 + sub * 8 : a.dout;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(abase + j * atom + r * ROWB + swz64(r, sub) * 16),
                                 "l"(srcp), "r"(ok ? 16u : 0u) : "memory");
                }
            }
            // B: gathered source rows, one tile per tap
            const uint32_t bbase = smem_u32(sB + (size_t)buf * b_buf);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                if (t < a.T) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = rr + i * 32;
                        const __half* srcp = idx[t][i] >= 0 ? a.src + (size_t)idx[t][i] * a.src_stride + (t & cmask) * 32 + sub * 8 : a.src;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(bbase + t * atom + r * ROWB + swz64(r, sub) * 16),
                                     "l"(srcp), "r"(idx[t][i] >= 0 ? 16u : 0u) : "memory");
                    }
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full + 8 * buf);
        }
        // epilogue (once): D[co][t*32 + ci] -> vector reductions; accumulator row = output channel = TMEM lane
        if (my_tiles > 0) {
            const int q = warp, co = q * 32 + lane;
            mbar_wait(tfull, 0);
            tc_fence_after();
            if (q * 32 < a.Cout) {
                for (int c0 = 0; c0 < a.T * 32; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
                    tmem_ld_wait();
                    if (co < a.Cout) {
                        float* drow = a.dw + (size_t)co * a.T * 32 + c0;
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            red_add_v4(drow + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                       __uint_as_float(r[i + 3]));
                    }
                }
            }
        }
    } else {
        // whole warp, warp-uniform operands, one elected lane issues
        const uint32_t tmem_u = uniform_u32(tmem_base);
        const uint32_t idesc = instr_desc_f16(128, 32, 1, 1);   // both operands MN-major
        const uint32_t lay = swizzle_layout(ROWB);
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        int buf = 0, ph = 0;
        bool first = true;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            mbar_wait(full + 8 * buf, ph);
            tc_fence_after();
            const uint32_t abase = sA_u + buf * a_buf, bbase = sB_u + buf * b_buf;
            if (elect_one()) {
                for (int t = 0; t < a.T; ++t) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {   // 128 sites = 8 x UMMA_K(16)
                        const uint64_t da = smem_desc(abase + k * 16 * ROWB, atom, 8 * ROWB, lay);
                        const uint64_t db = smem_desc(bbase + t * atom + k * 16 * ROWB, atom, 8 * ROWB, lay);
                        mma_f16(tmem_u + t * 32, da, db, idesc, (first && k == 0) ? 0u : 1u);
                    }
                }
                mma_commit(empty + 8 * buf);
            }
            first = false;
            if (buf) ph ^= 1;
            buf ^= 1;
        }
        if (my_tiles > 0 && elect_one()) mma_commit(tfull);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

}  // namespace

namespace mg {

int sparse_wgrad_persistent_launch(const void* dout, int dout_stride, int Cout, const void* src, int src_stride, int Cin,
                                   const int32_t* table, int T, int No, float* dw, void* stream, bool* handled) {
    *handled = false;
    const char* e = std::getenv("MAGGIE_B200_NO_PERSISTENT_SPARSE");
    if (e && e[0] == '1') return MG_OK;
    const int cout_eff = (Cout + 31) / 32 * 32;
    if ((Cin != 32 && Cin != 64) || cout_eff > 128 || dout_stride < cout_eff) return MG_OK;
    const int cl = Cin == 64 ? 1 : 0, VT = T << cl;       // virtual taps of 32 channels (dw column = vt * 32 + ci)
    if (VT > 9) return MG_OK;
    const int n_tiles = ceil_div(No, 128);
    if (n_tiles < 2 * kNumSMs) return MG_OK;
    GArgs a;
    a.dout = static_cast<const __half*>(dout), a.dout_stride = dout_stride, a.Cout = Cout, a.cout_eff = cout_eff;
    a.src = static_cast<const __half*>(src), a.src_stride = src_stride;
    a.table = table, a.T = VT, a.Tt = T, a.cl = cl, a.No = No, a.n_tiles = n_tiles, a.dw = dw;
    a.tmem_cols = 32;
    while (a.tmem_cols < VT * 32) a.tmem_cols <<= 1;
    const size_t smem = 1024 + 2 * (size_t)(4 + VT) * 128 * ROWB + 256;
    if (smem > 224 * 1024) return MG_OK;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(sparse_wgrad_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) {
            set_error("mg_sparse_wgrad: cannot raise dynamic shared memory limit (persistent kernel)");
            return MG_ERR_CUDA;
        }
        attr_set = true;
    }
    MG_LAUNCH(sparse_wgrad_persistent_kernel, std::min(n_tiles, kNumSMs), G_THREADS, smem, stream, a);
    MG_CHECK_LAUNCH("mg_sparse_wgrad(persistent)");
    *handled = true;
    return MG_OK;
}

}  // namespace mg
