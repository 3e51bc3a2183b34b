// K9: sparse refinement kernels - rulebook convolutions on the active-site lists, no spconv.
//
//   out[p][co] = sum_t sum_ci  src[table[p][t]][ci] * W[co][t][ci]  (+ bias)        table entry -1 = no neighbour
//
// gather -> dense tile -> tensor-core MMA -> scatter:  one CTA owns 128 output sites.  Four producer warps gather the
// neighbour rows of one tap (coalesced 16-byte pieces of each source row) straight into shared memory in the
// 64/128-byte-swizzled K-major operand layout that tcgen05 expects (the same layout TMA produces in the dense
// kernel), the weight pack is resident in shared memory, one thread issues tcgen05.mma into a TMEM accumulator,
// and the same four warps run the epilogue (bias, BatchNorm1d statistic partials, fp16 rows at a column offset so
// concatenations never materialise, or the fp32 (-99 filled) logit map of the two heads).
// The same kernel computes SubMConv2d (neighbour table), SparseInverseConv2d (parent table), 1x1 / Linear layers
// (no table) and all their data gradients (transposed table + transposed pack).  The weight gradient kernel uses
// the sites as the GEMM K dimension with both row tiles as MN-major operands (like the dense K4).
#include "common.cuh"
#include "ptx.cuh"

namespace {

using namespace mg::ptx;

constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;

// ---- swizzled row tile: 128 rows of ROWB bytes, chunk j of row r lives at chunk (j ^ f(r)) ---------------------
__device__ __forceinline__ uint32_t swz_chunk(int r, int j, int rowb) {
    return rowb == 128 ? (uint32_t)(j ^ (r & 7)) : (rowb == 64 ? (uint32_t)(j ^ ((r >> 1) & 3)) : (uint32_t)(j ^ ((r >> 2) & 1)));
}
__device__ __forceinline__ void st_tile(uint8_t* tile, int r, int j, int rowb, uint4 v) {
    *reinterpret_cast<uint4*>(tile + (size_t)r * rowb + swz_chunk(r, j, rowb) * 16) = v;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ================================================================================================ gather rows
__global__ void __launch_bounds__(256)
gather_rows_kernel(const __half* __restrict__ dense, const int32_t* __restrict__ coords, int n, int n_i, int H, int W, int C,
                   __half* __restrict__ out, int out_stride, int c_off) {
    mg::pdl_prologue();
    const int G = C >> 3;
    const size_t total = (size_t)n * G;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(v / G), g = (int)(v - (size_t)r * G);
        const int s = coords[r * 3], y = coords[r * 3 + 1], x = coords[r * 3 + 2];
        const size_t pix = ((size_t)(s / n_i) * H + y) * W + x;
        *reinterpret_cast<uint4*>(out + (size_t)r * out_stride + c_off + g * 8) =
            __ldg(reinterpret_cast<const uint4*>(dense + pix * C + g * 8));
    }
}

// d(dense)[frame,y,x,:] += g[row, c_off:c_off+C]   (several instance slots may share a pixel -> half2 atomics)
__global__ void __launch_bounds__(256)
scatter_rows_add_kernel(const __half* __restrict__ g, int g_stride, int c_off, const int32_t* __restrict__ coords, int n,
                        int n_i, int H, int W, int C, __half* __restrict__ ddense) {
    mg::pdl_prologue();
    const int G = C >> 1;
    const size_t total = (size_t)n * G;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(v / G), c2 = (int)(v - (size_t)r * G);
        const int s = coords[r * 3], y = coords[r * 3 + 1], x = coords[r * 3 + 2];
        const size_t pix = ((size_t)(s / n_i) * H + y) * W + x;
        const __half2 val = *reinterpret_cast<const __half2*>(g + (size_t)r * g_stride + c_off + c2 * 2);
        atomicAdd(reinterpret_cast<__half2*>(ddense + pix * C) + c2, val);
    }
}

// ================================================================================================ forward / dgrad
struct SArgs {
    const __half* src; int src_stride;
    const int32_t* table; int T, No, Cin, Cout;      // Cout = padded N of the MMA (multiple of 16)
    const __half* w;                                 // [Cout][T*Cin]
    const float* bias;
    __half* out; int out_stride, c_off, Cout_real;
    float* stats;
    float* map; const int32_t* coords; int mapH, mapW;  // head mode: fp32 logit map [slots,mapH,mapW], col 0 only
    int BK, kblocks, rowb, stages, pre_act;
};

__global__ void __launch_bounds__(160, 1)
sparse_conv_kernel(const SArgs a) {
    mg::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int steps = a.T * a.kblocks;
    const int b_tile = ((a.Cout * a.rowb + 1023) / 1024) * 1024;   // one [Cout x BK] weight sub-tile, 1 KB aligned
    const int a_tile = 128 * a.rowb;
    uint8_t* sB = smem;
    uint8_t* sA = sB + (size_t)steps * b_tile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)a.stages * a_tile);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * a.stages + 1);
    // the epilogue's transposition / partial-sum buffers alias the first operand stages (free once the accumulator is
    // complete): the 64 -> 64 3x3 layers then fit nine stages and take the all-gathers-in-flight path
    float* s_stage = reinterpret_cast<float*>(sA);
    float* s_part = s_stage + 4 * 32 * 17;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * a.stages, tfull = empty0 + 8 * a.stages;
    const int tile0 = blockIdx.x * 128;
    const int cpr = a.rowb >> 4;  // 16-byte chunks per row
    const uint32_t tmem_cols = a.Cout < 32 ? 32 : a.Cout;

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full0 + 8 * s, 128);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    // resident weights: sub-tile (t, kb) = W[:, t*Cin + kb*BK : +BK] as Cout rows of rowb bytes, swizzled
    {
        const int Ktot = a.T * a.Cin;
        const int chunks = steps * a.Cout * cpr;
        // asynchronous copies: the whole pack (up to 74 KB) is in flight at once instead of one dependent
        // load -> store round trip per 16 bytes and thread
        for (int i = tid; i < chunks; i += 160) {
            const int j = i % cpr, n = (i / cpr) % a.Cout, st = i / (cpr * a.Cout);
            const int t = st / a.kblocks, kb = st - t * a.kblocks;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB + (size_t)st * b_tile) + n * a.rowb +
                         swz_chunk(n, j, a.rowb) * 16), "l"(a.w + (size_t)n * Ktot + t * a.Cin + kb * a.BK + j * 8) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===== gather producers =====
        const int sub = tid % cpr, rr = tid / cpr, rstep = 128 / cpr;
        if (a.stages == steps && a.kblocks == 1 && steps <= 9) {
            // Every tap has its own stage: nothing is recycled, so ALL gathers of the tile are put in flight at once -
            // first every neighbour index of this thread's rows, then one 16-byte cp.async per (tap, row) straight into
            // the swizzled operand tile (zero-fill form for missing neighbours) - and the stages are handed to the MMA
            // thread in order as their copy groups land.  (The staged loop below serialises two dependent global-load
            // latencies per tap: 18 latencies per tile instead of 2.)
            // (64-channel rows, 8 chunks of 16 bytes: 8 rows per thread and tap - this path served 32-channel rows only
            //  at first, and the 3x3 64-channel layers on the OS4 / OS8 site lists took 45-117 us in the staged loop)
            int idx[9][8];
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int p = tile0 + rr + i * rstep;
                    idx[t][i] = -1;
                    if (t < steps && i < cpr && p < a.No) idx[t][i] = a.table ? __ldg(a.table + (size_t)p * a.T + t) : p;
                }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                if (t < steps) {
                    const uint32_t tile = smem_u32(sA + (size_t)t * a_tile);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < cpr) {
                            const int r = rr + i * rstep;
                            const __half* srcp = idx[t][i] >= 0 ? a.src + (size_t)idx[t][i] * a.src_stride + sub * 8 : a.src;
                            const uint32_t nbytes = idx[t][i] >= 0 ? 16u : 0u;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile + r * a.rowb + swz_chunk(r, sub, a.rowb) * 16),
                                         "l"(srcp), "r"(nbytes) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
#define MG_HAND_OVER_STAGE(T)                                                         \
    if (T < steps) {                                                                  \
        asm volatile("cp.async.wait_group %0;" ::"n"(8 - T) : "memory");              \
        fence_async_smem();                                                           \
        mbar_arrive(full0 + 8 * T);                                                   \
    }
            MG_HAND_OVER_STAGE(0) MG_HAND_OVER_STAGE(1) MG_HAND_OVER_STAGE(2) MG_HAND_OVER_STAGE(3) MG_HAND_OVER_STAGE(4)
            MG_HAND_OVER_STAGE(5) MG_HAND_OVER_STAGE(6) MG_HAND_OVER_STAGE(7) MG_HAND_OVER_STAGE(8)
#undef MG_HAND_OVER_STAGE
        } else {
            // General path (128-byte rows, several K blocks, or more steps than stages): steps go in batches of `stages`;
            // within a batch every row copy is an asynchronous 16-byte cp.async, the neighbour indices of the next step
            // are fetched while the copies of the current one are being issued, and the stages are handed to the MMA
            // thread in order as their copy groups land.
            for (int st0 = 0; st0 < steps; st0 += a.stages) {
                const int nb = min(a.stages, steps - st0);
                int idx[8], idx_n[8];
                auto load_idx = [&](int st, int (&dst)[8]) {
                    const int t = st / a.kblocks;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int p = tile0 + rr + i * rstep;
                        dst[i] = -1;
                        if (i < cpr && p < a.No) dst[i] = a.table ? __ldg(a.table + (size_t)p * a.T + t) : p;
                    }
                };
                load_idx(st0, idx_n);
                for (int b = 0; b < nb; ++b) {
                    const int st = st0 + b, s = st % a.stages, ph = (st / a.stages) & 1;
                    const int kb = st - (st / a.kblocks) * a.kblocks;
#pragma unroll
                    for (int i = 0; i < 8; ++i) idx[i] = idx_n[i];
                    if (b + 1 < nb) load_idx(st + 1, idx_n);
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    const uint32_t tile = smem_u32(sA + (size_t)s * a_tile);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < cpr) {
                            const int r = rr + i * rstep;
                            const __half* srcp = idx[i] >= 0 ? a.src + (size_t)idx[i] * a.src_stride + kb * a.BK + sub * 8 : a.src;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile + r * a.rowb + swz_chunk(r, sub, a.rowb) * 16),
                                         "l"(srcp), "r"(idx[i] >= 0 ? 16u : 0u) : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
                for (int b = 0; b < nb; ++b) {
                    switch (nb - 1 - b) {   // cp.async.wait_group takes an immediate
                        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
                        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
                        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
                        case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
                        default: asm volatile("cp.async.wait_group 8;" ::: "memory"); break;
                    }
                    fence_async_smem();
                    mbar_arrive(full0 + 8 * ((st0 + b) % a.stages));
                }
            }
        }
        // ===== epilogue =====
        const int q = warp, row = tile0 + q * 32 + lane;
        const bool valid = row < a.No;
        float* stg = s_stage + q * 32 * 17;
        float* part = s_part + q * 2 * a.Cout;
        mbar_wait(tfull, 0);
        tc_fence_after();
        for (int c0 = 0; c0 < a.Cout; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(r[i]);
                if (a.bias && c0 + i < a.Cout_real) v[i] += __ldg(a.bias + c0 + i);
                if (a.pre_act == 1) v[i] = fmaxf(v[i], 0.f);
            }
            if (a.stats) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 16; ++i) stg[lane * 17 + i] = valid ? v[i] : 0.f;
                __syncwarp();
                const int col = lane & 15;
                float acc = 0.f;
                if (lane < 16) {
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) acc += stg[k * 17 + col];
                } else {
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) { const float z = stg[k * 17 + col]; acc += z * z; }
                }
                part[(lane >> 4) * a.Cout + c0 + col] = acc;
            }
            if (valid) {
                if (a.map) {
                    if (c0 == 0) {
                        const int s = a.coords[row * 3], y = a.coords[row * 3 + 1], x = a.coords[row * 3 + 2];
                        // the reference computes dense()-99 then += 99 at the active sites (fp32 rounding included)
                        a.map[((size_t)s * a.mapH + y) * a.mapW + x] = (v[0] - 99.0f) + 99.0f;
                    }
                } else if (c0 < a.Cout_real) {
                    uint4 o0, o1;
                    __half2 h;
                    h = __floats2half2_rn(v[0], v[1]);   o0.x = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[2], v[3]);   o0.y = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[4], v[5]);   o0.z = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[6], v[7]);   o0.w = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[8], v[9]);   o1.x = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[10], v[11]); o1.y = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[12], v[13]); o1.z = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(v[14], v[15]); o1.w = *reinterpret_cast<uint32_t*>(&h);
                    __half* orow = a.out + (size_t)row * a.out_stride + a.c_off + c0;
                    reinterpret_cast<uint4*>(orow)[0] = o0;
                    reinterpret_cast<uint4*>(orow)[1] = o1;
                }
            }
        }
        if (a.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            float* dst = a.stats + (size_t)(blockIdx.x % STAT_COPIES) * 2 * a.Cout_real;
            for (int i = tid; i < 2 * a.Cout; i += 128) {
                const int kind = i / a.Cout, c = i - kind * a.Cout;
                if (c < a.Cout_real) {
                    const float tot = s_part[0 * 2 * a.Cout + i] + s_part[1 * 2 * a.Cout + i] + s_part[2 * 2 * a.Cout + i] +
                                      s_part[3 * 2 * a.Cout + i];
                    atomicAdd(dst + kind * a.Cout_real + c, tot);
                }
            }
        }
    } else {
        // ===== MMA issuer: whole warp, warp-uniform operands (uniform registers, no elect / R2UR waterfall per
        // tcgen05.mma), one elected lane issues =====
        const uint32_t tmem_u = uniform_u32(tmem_base);
        const uint32_t idesc = instr_desc_f16(128, a.Cout, 0, 0);
        const uint32_t layout = swizzle_layout(a.rowb), sbo = 8 * a.rowb;
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        const int ksteps = a.BK / 16;
        int s = 0, ph = 0;
        for (int st = 0; st < steps; ++st) {
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            const uint64_t da = smem_desc(sA_u + s * a_tile, 0, sbo, layout), db = smem_desc(sB_u + st * b_tile, 0, sbo, layout);
            if (elect_one()) {
                mma_f16(tmem_u, da, db, idesc, st != 0);
                for (int k = 1; k < ksteps; ++k) mma_f16(tmem_u, da + 2 * k, db + 2 * k, idesc, 1u);
                mma_commit(empty0 + 8 * s);
            }
            if (++s == a.stages) s = 0, ph ^= 1;
        }
        if (elect_one()) mma_commit(tfull);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ================================================================================================ weight gradient
struct SWArgs {
    const __half* dout; int dout_stride, Cout;       // [No][>=Cout]
    const __half* src; int src_stride, Cin;
    const int32_t* table; int T, No;
    float* dw;                                       // [Cout][T*Cin] fp32, atomically accumulated
    int taps_per_cta, tiles_per_cta, n_tiles;
    int rowb_a, atoms_a, rowb_b, atoms_b;            // row bytes / number of MN atoms (A padded to M = 128)
};

__global__ void __launch_bounds__(160, 1)
sparse_wgrad_kernel(const SWArgs a) {
    mg::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int STAGES = 2;
    const int a_atom = 128 * a.rowb_a, a_tile = a.atoms_a * a_atom;      // atoms beyond Cout stay zero
    const int b_atom = 128 * a.rowb_b, b_tile = a.atoms_b * b_atom;
    uint8_t* sA = smem;                                                  // [STAGES] d_out tiles
    uint8_t* sB = sA + STAGES * a_tile;                                  // [STAGES][taps_per_cta] gathered src tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)STAGES * a.taps_per_cta * b_tile);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, tfull = empty0 + 8 * STAGES;
    const int t0 = blockIdx.y * a.taps_per_cta, nt = min(a.taps_per_cta, a.T - t0);
    const int tile_begin = blockIdx.x * a.tiles_per_cta, tile_end = min(tile_begin + a.tiles_per_cta, a.n_tiles);
    const int nk = tile_end - tile_begin;
    const uint32_t cols_needed = (uint32_t)(nt * a.Cin);
    uint32_t tmem_cols = 32;
    while (tmem_cols < cols_needed) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 128);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    for (int i = tid; i < STAGES * a_tile / 16; i += 160) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (nk > 0) {
        if (warp < 4) {
            const int cpa = a.rowb_a >> 4, cpb = a.rowb_b >> 4;
            const int ca_chunks = a.Cout >> 3, cb_chunks = a.Cin >> 3;   // 16-byte chunks per source row
            for (int i = 0; i < nk; ++i) {
                const int s = i % STAGES, ph = (i / STAGES) & 1;
                const int row0 = (tile_begin + i) * 128;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                // Both operand tiles are filled with 16-byte cp.async (zero-fill form for rows past the end / missing
                // neighbours): all neighbour indices of the tile first, then every copy of every tap in flight at once.
                // (Register-staged loads serialised two dependent global-load latencies per tap.)
                // A: d_out rows (K index = site)
                const uint32_t ta = smem_u32(sA + (size_t)s * a_tile);
#pragma unroll 4
                for (int e = tid; e < 128 * ca_chunks; e += 128) {
                    const int r = e / ca_chunks, c = e - r * ca_chunks;
                    const int p = row0 + r;
                    const bool ok = p < a.No;
                    const __half* srcp = ok ? a.dout + (size_t)p * a.dout_stride + c * 8 : a.dout;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ta + (c / cpa) * a_atom + r * a.rowb_a +
                                 swz_chunk(r, c % cpa, a.rowb_a) * 16), "l"(srcp), "r"(ok ? 16u : 0u) : "memory");
                }
                // B: gathered source rows, one tile per tap
                if (cb_chunks <= 8) {
                    // <= 8 (row, chunk) pairs per thread: fetch every neighbour index of the tile first (one latency),
                    // then put every copy in flight (64-channel rows: 8 pairs; they went through the loop below at first,
                    // one index round trip per pair)
                    int idx[8][9];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int e0 = tid + i * 128, r = e0 / cb_chunks, p = row0 + r;
#pragma unroll
                        for (int tt = 0; tt < 9; ++tt) {
                            idx[i][tt] = -1;
                            if (e0 < 128 * cb_chunks && tt < nt && p < a.No) idx[i][tt] = a.table ? __ldg(a.table + (size_t)p * a.T + t0 + tt) : p;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int e0 = tid + i * 128, r = e0 / cb_chunks, c = e0 - r * cb_chunks;
                        if (e0 < 128 * cb_chunks) {
#pragma unroll
                            for (int tt = 0; tt < 9; ++tt) {
                                if (tt < nt) {
                                    const uint32_t tb = smem_u32(sB + (size_t)(s * a.taps_per_cta + tt) * b_tile);
                                    const __half* srcp = idx[i][tt] >= 0 ? a.src + (size_t)idx[i][tt] * a.src_stride + c * 8 : a.src;
                                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tb + (c / cpb) * b_atom + r * a.rowb_b +
                                                 swz_chunk(r, c % cpb, a.rowb_b) * 16), "l"(srcp), "r"(idx[i][tt] >= 0 ? 16u : 0u) : "memory");
                                }
                            }
                        }
                    }
                } else
                for (int e0 = tid; e0 < 128 * cb_chunks; e0 += 128) {
                    const int r = e0 / cb_chunks, c = e0 - r * cb_chunks;
                    const int p = row0 + r;
                    int idx[9];
#pragma unroll
                    for (int tt = 0; tt < 9; ++tt) {
                        idx[tt] = -1;
                        if (tt < nt && p < a.No) idx[tt] = a.table ? __ldg(a.table + (size_t)p * a.T + t0 + tt) : p;
                    }
#pragma unroll
                    for (int tt = 0; tt < 9; ++tt) {
                        if (tt < nt) {
                            const uint32_t tb = smem_u32(sB + (size_t)(s * a.taps_per_cta + tt) * b_tile);
                            const __half* srcp = idx[tt] >= 0 ? a.src + (size_t)idx[tt] * a.src_stride + c * 8 : a.src;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tb + (c / cpb) * b_atom + r * a.rowb_b +
                                         swz_chunk(r, c % cpb, a.rowb_b) * 16), "l"(srcp), "r"(idx[tt] >= 0 ? 16u : 0u) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                fence_async_smem();
                mbar_arrive(full0 + 8 * s);
            }
            // epilogue: D[co][tap*Cin + ci] -> atomic add
            const int q = warp, co = q * 32 + lane;
            mbar_wait(tfull, 0);
            tc_fence_after();
            for (int c0 = 0; c0 < nt * a.Cin; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
                tmem_ld_wait();
                if (co < a.Cout) {
                    float* drow = a.dw + (size_t)co * a.T * a.Cin + (size_t)t0 * a.Cin + c0;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        red_add_v4(drow + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                   __uint_as_float(r[i + 3]));
                }
            }
        } else {
            // whole warp, warp-uniform operands, one elected lane issues
            const uint32_t tmem_u = uniform_u32(tmem_base);
            const uint32_t idesc = instr_desc_f16(128, a.Cin, 1, 1);
            const uint32_t layA = swizzle_layout(a.rowb_a), layB = swizzle_layout(a.rowb_b);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            int s = 0, ph = 0;
            for (int i = 0; i < nk; ++i) {
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t abase = sA_u + s * a_tile;
                if (elect_one()) {
                    for (int tt = 0; tt < nt; ++tt) {
                        const uint32_t bbase = sB_u + (s * a.taps_per_cta + tt) * b_tile;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint64_t da = smem_desc(abase + k * 16 * a.rowb_a, a_atom, 8 * a.rowb_a, layA);
                            const uint64_t db = smem_desc(bbase + k * 16 * a.rowb_b, b_atom, 8 * a.rowb_b, layB);
                            mma_f16(tmem_u + tt * a.Cin, da, db, idesc, (i | k) != 0);
                        }
                    }
                    mma_commit(empty0 + 8 * s);
                }
                if (++s == STAGES) s = 0, ph ^= 1;
            }
            if (elect_one()) mma_commit(tfull);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

int rowbytes_for(int c) { return c >= 64 ? 128 : (c >= 32 ? 64 : 32); }

}  // namespace

extern "C" int mg_gather_rows(const void* dense, const int32_t* coords, int n, int n_i, int H, int W, int C, void* out,
                              int out_stride, int c_off, void* stream) {
    MG_REQUIRE(dense && out && (coords || n == 0), "mg_gather_rows: null pointer");
    MG_REQUIRE(C % 8 == 0 && out_stride % 8 == 0 && c_off % 8 == 0 && n_i > 0, "mg_gather_rows: C, out_stride, c_off must be multiples of 8");
    if (n == 0) return MG_OK;
    const int grid = (int)std::min<size_t>(((size_t)n * (C / 8) + 255) / 256, (size_t)mg::kNumSMs * 8);
    MG_LAUNCH(gather_rows_kernel, grid, 256, 0, stream, static_cast<const __half*>(dense), coords, n, n_i, H, W, C,
              static_cast<__half*>(out), out_stride, c_off);
    MG_CHECK_LAUNCH("mg_gather_rows");
    return MG_OK;
}

extern "C" int mg_scatter_rows_add(const void* g, int g_stride, int c_off, const int32_t* coords, int n, int n_i, int H, int W,
                                   int C, void* ddense, void* stream) {
    MG_REQUIRE(g && ddense && (coords || n == 0), "mg_scatter_rows_add: null pointer");
    MG_REQUIRE(C % 2 == 0 && g_stride % 2 == 0 && c_off % 2 == 0 && n_i > 0, "mg_scatter_rows_add: even C / stride / offset required");
    if (n == 0) return MG_OK;
    const int grid = (int)std::min<size_t>(((size_t)n * (C / 2) + 255) / 256, (size_t)mg::kNumSMs * 8);
    MG_LAUNCH(scatter_rows_add_kernel, grid, 256, 0, stream, static_cast<const __half*>(g), g_stride, c_off, coords, n, n_i, H,
              W, C, static_cast<__half*>(ddense));
    MG_CHECK_LAUNCH("mg_scatter_rows_add");
    return MG_OK;
}

namespace mg {
int sparse_conv_persistent_launch(const mg_sparse_conv_desc* d, void* stream, bool* handled);   // k9b_sparse_persistent.cu
int sparse_wgrad_persistent_launch(const void* dout, int dout_stride, int Cout, const void* src, int src_stride, int Cin,
                                   const int32_t* table, int T, int No, float* dw, void* stream, bool* handled);
}

extern "C" int mg_sparse_conv(const mg_sparse_conv_desc* d, void* stream) {
    MG_REQUIRE(d && d->src && d->w && (d->out || d->map), "mg_sparse_conv: null pointer");
    MG_REQUIRE(d->T >= 1 && d->T <= 9 && (d->table || d->T == 1), "mg_sparse_conv: T=%d needs a table", d->T);
    MG_REQUIRE(d->Cin % 32 == 0 && d->Cin <= 128, "mg_sparse_conv: Cin must be 32, 64, 96 or 128 (got %d)", d->Cin);
    MG_REQUIRE(d->Cout >= 1 && d->Cout <= 128, "mg_sparse_conv: Cout out of range (%d)", d->Cout);
    MG_REQUIRE(d->src_stride % 8 == 0, "mg_sparse_conv: src_stride must be a multiple of 8");
    MG_REQUIRE(!d->map || d->coords, "mg_sparse_conv: head mode needs coords");
    MG_REQUIRE(d->map || (d->Cout % 16 == 0 && d->out_stride % 8 == 0 && d->c_off % 8 == 0), "mg_sparse_conv: row output needs Cout %% 16 == 0");
    if (d->No <= 0) return MG_OK;
    {
        // 32-channel layers on the large site lists: persistent kernel (K9b)
        bool handled = false;
        const int rc = mg::sparse_conv_persistent_launch(d, stream, &handled);
        if (rc != MG_OK || handled) return rc;
    }
    SArgs a;
    a.src = static_cast<const __half*>(d->src), a.src_stride = d->src_stride;
    a.table = d->table, a.T = d->T, a.No = d->No, a.Cin = d->Cin;
    a.Cout_real = d->Cout, a.Cout = (d->Cout + 15) / 16 * 16;
    a.w = static_cast<const __half*>(d->w);  // caller packs [Cout padded to 16][T*Cin]
    a.bias = d->bias, a.out = static_cast<__half*>(d->out), a.out_stride = d->out_stride, a.c_off = d->c_off;
    a.stats = d->stats, a.map = d->map, a.coords = d->coords, a.mapH = d->mapH, a.mapW = d->mapW, a.pre_act = d->pre_act;
    a.BK = d->Cin % 64 == 0 ? 64 : 32;
    a.kblocks = d->Cin / a.BK, a.rowb = a.BK * 2;
    const int steps = a.T * a.kblocks;
    const int b_tile = ((a.Cout * a.rowb + 1023) / 1024) * 1024;
    // Stage recycling costs a tcgen05.commit -> mbarrier -> producer round trip (~1.3 us measured in K2b): give every
    // tap its own stage whenever two CTAs still fit per SM, so that the gathers of a tile never wait for the MMAs.
    const size_t fixed_smem = 1024 + (size_t)steps * b_tile + 256;   // (epilogue scratch aliases the first operand stages)
    a.stages = std::min(4, std::max(2, steps));
    while (a.stages < steps && fixed_smem + (size_t)(a.stages + 1) * 128 * a.rowb <= 112 * 1024) ++a.stages;
    if (a.rowb == 128)   // 64-channel layers run on few sites (OS4 / OS8): latency matters there, not occupancy
        while (a.stages < std::min(steps, 9) && fixed_smem + (size_t)(a.stages + 1) * 128 * a.rowb <= 227 * 1024) ++a.stages;
    const size_t smem = fixed_smem + (size_t)a.stages * 128 * a.rowb;
    MG_REQUIRE(smem <= 227 * 1024, "mg_sparse_conv: weight pack does not fit in shared memory (%zu B)", smem);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(sparse_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            mg::set_error("mg_sparse_conv: cannot raise dynamic shared memory limit");
            return MG_ERR_CUDA;
        }
        attr_set = true;
    }
    MG_LAUNCH(sparse_conv_kernel, mg::ceil_div(d->No, 128), 160, smem, stream, a);
    MG_CHECK_LAUNCH("mg_sparse_conv");
    return MG_OK;
}

extern "C" int mg_sparse_wgrad(const void* dout, int dout_stride, int Cout, const void* src, int src_stride, int Cin,
                               const int32_t* table, int T, int No, float* dw, void* stream) {
    MG_REQUIRE(dout && src && dw && (table || T == 1), "mg_sparse_wgrad: null pointer");
    MG_REQUIRE(Cout % 8 == 0 && Cout <= 128 && Cin % 16 == 0 && Cin <= 128 && (Cin & (Cin - 1)) == 0 && T >= 1 && T <= 9,
               "mg_sparse_wgrad: unsupported shape Cout=%d Cin=%d T=%d", Cout, Cin, T);
    MG_REQUIRE(dout_stride % 8 == 0 && src_stride % 8 == 0, "mg_sparse_wgrad: strides must be multiples of 8");
    if (No <= 0) return MG_OK;
    {
        bool handled = false;
        const int rc = mg::sparse_wgrad_persistent_launch(dout, dout_stride, Cout, src, src_stride, Cin, table, T, No, dw, stream, &handled);
        if (rc != MG_OK || handled) return rc;
    }
    SWArgs a;
    a.dout = static_cast<const __half*>(dout), a.dout_stride = dout_stride, a.Cout = Cout;
    a.src = static_cast<const __half*>(src), a.src_stride = src_stride, a.Cin = Cin;
    a.table = table, a.T = T, a.No = No, a.dw = dw;
    a.rowb_a = rowbytes_for(Cout), a.atoms_a = 128 / (a.rowb_a / 2);   // M padded to 128 channels
    a.rowb_b = rowbytes_for(Cin), a.atoms_b = Cin / (a.rowb_b / 2);
    const int a_tile = a.atoms_a * 128 * a.rowb_a, b_tile = a.atoms_b * 128 * a.rowb_b;
    a.taps_per_cta = std::min(T, std::min(512 / Cin, (int)((200 * 1024 - 2 * a_tile) / (2 * b_tile))));
    MG_REQUIRE(a.taps_per_cta >= 1, "mg_sparse_wgrad: tile does not fit");
    const int tap_groups = mg::ceil_div(T, a.taps_per_cta);
    a.n_tiles = mg::ceil_div(No, 128);
    int splits = std::max(1, std::min(a.n_tiles, mg::ceil_div(2 * mg::kNumSMs, tap_groups)));
    a.tiles_per_cta = mg::ceil_div(a.n_tiles, splits);
    splits = mg::ceil_div(a.n_tiles, a.tiles_per_cta);
    const size_t smem = 1024 + 2 * (size_t)a_tile + 2 * (size_t)a.taps_per_cta * b_tile + 256;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(sparse_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            mg::set_error("mg_sparse_wgrad: cannot raise dynamic shared memory limit");
            return MG_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid(splits, tap_groups);
    MG_LAUNCH(sparse_wgrad_kernel, grid, 160, smem, stream, a);
    MG_CHECK_LAUNCH("mg_sparse_wgrad");
    return MG_OK;
}
