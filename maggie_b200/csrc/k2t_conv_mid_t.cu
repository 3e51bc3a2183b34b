// K2t: TRANSPOSED halo-patch convolution for the mid-resolution layers (Ci >= 128, Co a multiple of 128, width <= 64,
// stride 1, taps within a small halo): D[Cout x pixels] = W[Cout x K] * X[K x pixels].
//
// Why transposed.  tools/mma_bench.cu measures the issue cost of one tcgen05.mma (M = 128, K = 16, both operands in shared
// memory) on B200 as ~37 + 0.31 * N cycles up to N ~ 200 and 0.55 * N beyond: below N = 256 the instruction is bound by
// the shared-memory reads of its operands (4 KB of A + 32 * N bytes of B at 128 B/clk), not by the tensor pipe
// (N = 64: 57 cycles for 32 cycles of math, N = 128: 77 for 64, N = 256: 141 for 128).  K2h issues N = 64 MMAs (pixels
// as M, five 128-pixel blocks x 64 output channels per TMEM set) and measured 12.8 us of MMA time for 5.9 us of math
// on 64^2 128->128.  Here the WEIGHTS are the M = 128 operand and the PIXELS of the halo patch are the N operand, and N is
// free in steps of 16 up to 256: a slab of R image rows (R * (W + 2) linear patch pixels, <= 512 = all TMEM columns) is
// covered by one or two MMAs of N = 128..256 per k step - 86-91 % of the tensor rate instead of 56 %, and no rounding
// of the slab to 128-pixel blocks (K2h computes 640 accumulator rows for 512 pixels).
//   * B operand (activations): per 64-channel chunk ONE 4-D TMA box {64 ch, P = W + 2h pixels, R + 2h rows} brings the
//     halo patch of the slab as a linear array of pixels (128-byte swizzled rows, hardware zero fill = padding).  The
//     operand of tap (dy, dx) is the same buffer read from pixel (h + dy) * P + dx on (shifted descriptor start, as in
//     K2b / K2h); the whole patch stays resident for all taps and k steps of the item.
//   * A operand (weights): [128 Cout x 64 ch] blocks of one (tap, chunk) stream through a ring (16 KB each).
//   * accumulator: lane = output channel, column = linear patch pixel.  Epilogue (8 warps): tcgen05.ld gives a thread 16
//     pixels of ITS channel - BatchNorm sum / sum of squares are plain per-thread accumulations, bias / affine are
//     per-thread scalars; the fp16 results are transposed through shared memory (the patch buffer, free by then) into a
//     dense [R][W][128] box and leave with ONE TMA store (pad columns and rows beyond the image are never staged).
// One CTA per SM, persistent over (slab, 128-channel tile) items; slab height chosen by a small cost model
// (MMA issue cost vs. the ~39 B/clk an SM takes in through TMA vs. number of waves).
// Measured (CUDA-graph replay, 8 frames): 64^2 128->128 12.8 us (755 TFLOP/s; K2h 19.8, generic K2 14.0; with the
// BatchNorm statistics 13.2 vs 22.2 / 18.9), 32^2 256->256 13.1 us (K2h 22.2, K2 18.8), 16^2 512->512 19.9 us (30.9 / 27.0).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

#include <cstdlib>

namespace {

using namespace mg::ptx;

constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;                         // warp 0: TMA producer, 1: MMA issue + TMEM, 2-3: idle
constexpr int THREADS = 32 * (FIRST_EPI_WARP + EPI_WARPS);
constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;
constexpr int MAX_CHUNKS = 8, MAX_NB = 10;
constexpr int BM = 128;                                   // output channels per item
constexpr int CH = 64;                                    // channels per stage: 128-byte swizzled rows
constexpr int B_BYTES = BM * CH * 2;                      // one weight block

struct TArgs {
    int n_taps, tap_off[9], tap_koff[9];                  // tap_off: patch-pixel offset (hal + dy) * P + dx of output pixel 0
    int H, W, Ci, Co;
    int P, R, hal, rblocks, n_ntiles, n_items;
    int chunks, cols, nparts, part_n[2], NB, a_bytes, a_al, tmem_cols;
    int pre_act, post_act;
    float* stats;
    const float *bias, *scale, *shift;
    const __half* res;
    int c_off;
    unsigned long long* trace;   // profiling only (mg_conv_midt_trace): per CTA 8 globaltimer stamps, see the kernel
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define MIDT_STAMP(slot)                                                         \
    do {                                                                         \
        if (a.trace && lane == 0) a.trace[blockIdx.x * 8 + (slot)] = gtime();    \
    } while (0)

__device__ __forceinline__ float act_apply(float v, int act) {
    return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v > 0.f ? v : 0.2f * v) : v);
}
__device__ __forceinline__ void sts_u16(uint32_t addr, unsigned short v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// EPI: 0 / 1 / 2 = lean epilogue (pre-activation none / ReLU / LeakyReLU fixed at compile time: training forward and data
// gradients), 3 = general (bias, eval BN affine, residual, post-activation), 4 = bias only (Linear layers over dense rows:
// the run-time switches of the general form tripled the instruction-bound epilogue, 6.7 -> 18 us on 32768 x 128 -> 128).
template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
conv_midt_tcgen05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                         const __grid_constant__ CUtensorMap tmY, const TArgs a) {
    mg::pdl_launch();
    if (a.trace && threadIdx.x == 32) a.trace[blockIdx.x * 8 + 0] = gtime();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // [1 KB guard][patch: chunks x a_al][weight ring: NB x 16 KB][barriers][tmem slot]
    // (tap (-h, -h) of output pixel 0 reads up to h pixels before the patch, the last MMA up to 15 + 2 h P pixels past a
    //  chunk - into the next chunk or the weight ring: those columns are pad pixels / beyond the slab and never stored)
    uint8_t* sX = smem + 1024;
    uint8_t* sW = sX + a.chunks * a.a_al;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + a.NB * B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + MAX_CHUNKS + 2 * MAX_NB + 4);
    int* s_tab = reinterpret_cast<int*>(tmem_slot + 4);      // [cols]: staging byte offset of accumulator column m, -1 = pad

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t x_full = bar0, w_full = x_full + 8 * MAX_CHUNKS, w_empty = w_full + 8 * MAX_NB,
                   t_full = w_empty + 8 * MAX_NB, t_empty = t_full + 8, p_free = t_empty + 8;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmY);
        for (int i = 0; i < a.chunks; ++i) mbar_init(x_full + 8 * i, 1);
        for (int i = 0; i < a.NB; ++i) mbar_init(w_full + 8 * i, 1), mbar_init(w_empty + 8 * i, 1);
        mbar_init(t_full, 1), mbar_init(t_empty, EPI_WARPS), mbar_init(p_free, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mg::pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) MIDT_STAMP(1);

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer: the patch chunks of an item + the (chunk, tap) weight blocks through the ring =====
            int s = 0, ph = 0, it = 0;
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
                const int pt = item / a.n_ntiles, n0 = (item - pt * a.n_ntiles) * BM;
                const int img = pt / a.rblocks, rb = pt - img * a.rblocks;
                if (it > 0) mbar_wait(p_free, (it - 1) & 1);    // the previous item's output staging aliases the patch
                int next_chunk = 0, issued = 0;
                for (int c = 0; c < a.chunks; ++c) {
                    for (int t = 0; t < a.n_taps; ++t) {
                        // patch chunk c goes out just before its first weight block; the others follow once the ring is
                        // primed (the first MMA needs chunk 0 + block 0 only)
                        while (next_chunk < a.chunks && (next_chunk <= c || issued >= a.NB)) {
                            mbar_expect_tx(x_full + 8 * next_chunk, a.a_bytes);
                            tma_load_4d(smem_u32(sX + next_chunk * a.a_al), &tmX, x_full + 8 * next_chunk, next_chunk * CH,
                                        -a.hal, rb * a.R - a.hal, img);
                            ++next_chunk;
                        }
                        mbar_wait(w_empty + 8 * s, ph ^ 1);
                        mbar_expect_tx(w_full + 8 * s, B_BYTES);
                        tma_load_2d(smem_u32(sW + s * B_BYTES), &tmW, w_full + 8 * s, a.tap_koff[t] + c * CH, n0);
                        ++issued;
                        if (++s == a.NB) s = 0, ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: whole warp runs the loop (uniform operands), one elected lane issues =====
        const uint32_t tmem_u = uniform_u32(tmem_base);
        const uint32_t idesc0 = instr_desc_f16(BM, a.part_n[0], 0, 0), idesc1 = instr_desc_f16(BM, a.part_n[1], 0, 0);
        const uint32_t lay = swizzle_layout(128), sbo = 1024;
        const uint32_t sX_u = smem_u32(sX), sW_u = smem_u32(sW);
        const uint32_t part1_rows = (uint32_t)(a.part_n[0] * 128) >> 4;
        int sb = 0, pb = 0, it = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
            if (it > 0) {
                mbar_wait(t_empty, (it - 1) & 1);
                tc_fence_after();
            }
            for (int c = 0; c < a.chunks; ++c) {
                mbar_wait(x_full + 8 * c, it & 1);
                tc_fence_after();
                if (it == 0 && c == 0) MIDT_STAMP(2);
                const uint64_t x_desc0 = smem_desc(sX_u + c * a.a_al, 0, sbo, lay);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    if (t < a.n_taps) {
                        mbar_wait(w_full + 8 * sb, pb);
                        tc_fence_after();
                        if (it == 0 && c == 0 && t == 0) MIDT_STAMP(3);
                        const uint64_t w_desc = smem_desc(sW_u + sb * B_BYTES, 0, sbo, lay);
                        const uint64_t xd = x_desc0 + (int64_t)((a.tap_off[t] * 128) >> 4);
                        const uint32_t first = (c | t) == 0 ? 0u : 1u;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < CH / 16; ++k) mma_f16(tmem_u, w_desc + 2 * k, xd + 2 * k, idesc0, first | (uint32_t)k);
                        }
                        if (a.nparts > 1) {
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < CH / 16; ++k)
                                    mma_f16(tmem_u + a.part_n[0], w_desc + 2 * k, xd + part1_rows + 2 * k, idesc1, first | (uint32_t)k);
                            }
                        }
                        if (elect_one()) mma_commit(w_empty + 8 * sb);
                        if (++sb == a.NB) sb = 0, pb ^= 1;
                    }
                }
            }
            if (elect_one()) mma_commit(t_full);
            if (it == 0) MIDT_STAMP(4);
        }
    } else if (warp >= FIRST_EPI_WARP) {
        // ===== epilogue: warps 4..11; TMEM lane quarter = warp % 4 (32 output channels), the two warps of a quarter
        // take alternate 16-pixel column groups.  The loop is instruction-bound (34 816 accumulators per item), so
        // everything that depends on the column only - is it a real pixel, where does it go in the staging box - comes
        // from a table built once per CTA while the first MMAs run, and the TMEM loads are double-buffered. =====
        const int q = warp & 3, half = (warp - FIRST_EPI_WARP) >> 2;
        const int co_l = q * 32 + lane;
        const uint32_t stage = smem_u32(sX);                   // [R][W][128] fp16, dense (the TMA store box)
        for (int m = threadIdx.x - 32 * FIRST_EPI_WARP; m < a.cols; m += 32 * EPI_WARPS) {
            const int row = m / a.P, j = m - row * a.P;
            s_tab[m] = (row < a.R && j >= a.hal && j < a.hal + a.W) ? (row * a.W + j - a.hal) * (BM * 2) : -1;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t st_l = stage + co_l * 2;
        const int ngroups = a.cols >> 4;
        int it = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
            const int pt = item / a.n_ntiles, n0 = (item - pt * a.n_ntiles) * BM;
            const int img = pt / a.rblocks, rb = pt - img * a.rblocks;
            const int y0 = rb * a.R;
            const int co = n0 + co_l;
            const uint32_t lim = (uint32_t)(min(a.R, a.H - y0) * a.W * (BM * 2));   // rows of the slab inside the image
            float bias = 0.f, sc = 1.f, sh = 0.f;
            const __half* res_l = nullptr;
            if constexpr (EPI == 3) {
                if (a.bias) bias = __ldg(a.bias + co);
                if (a.scale) sc = __ldg(a.scale + co), sh = __ldg(a.shift + co);
                if (a.res) res_l = a.res + ((size_t)img * a.H + y0) * a.W * a.Co + co;
            } else if constexpr (EPI == 4) {
                bias = __ldg(a.bias + co);
            }
            float s_sum = 0.f, s_sq = 0.f;
            auto process = [&](const uint32_t (&r)[16], int g) {
                const int4* t4 = reinterpret_cast<const int4*>(s_tab + g * 16);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int4 o4 = t4[c4];
                    const int off[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if ((uint32_t)off[i] < lim) {
                            float v = __uint_as_float(r[4 * c4 + i]);
                            if constexpr (EPI == 3) {
                                v = act_apply(v + bias, a.pre_act);
                            } else if constexpr (EPI == 1) {
                                v = fmaxf(v, 0.f);
                            } else if constexpr (EPI == 2) {
                                v = v > 0.f ? v : 0.2f * v;
                            } else if constexpr (EPI == 4) {
                                v += bias;
                            }
                            s_sum += v, s_sq = fmaf(v, v, s_sq);
                            if constexpr (EPI == 3) {
                                v = fmaf(v, sc, sh);
                                if (res_l) v += __half2float(__ldg(res_l + (size_t)(off[i] >> 8) * a.Co));
                                v = act_apply(v, a.post_act);
                            }
                            sts_u16(st_l + (uint32_t)off[i], __half_as_ushort(__float2half_rn(v)));
                        }
                    }
                }
            };
            mbar_wait(t_full, it & 1);
            tc_fence_after();
            if (it == 0 && warp == FIRST_EPI_WARP) MIDT_STAMP(5);
            uint32_t ra[16], rb2[16];
            int g = half;
            if (g < ngroups) tmem_ld16(tl + g * 16, ra);
            while (g < ngroups) {
                tmem_ld_wait();
                if (g + 2 < ngroups) tmem_ld16(tl + (g + 2) * 16, rb2);
                process(ra, g);
                g += 2;
                if (g >= ngroups) break;
                tmem_ld_wait();
                if (g + 2 < ngroups) tmem_ld16(tl + (g + 2) * 16, ra);
                process(rb2, g);
                g += 2;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty);               // accumulators drained: the next item's MMAs may start
            if (a.stats) {
                float* dst = a.stats + (size_t)(blockIdx.x % STAT_COPIES) * 2 * a.Co;
                atomicAdd(dst + co, s_sum);
                atomicAdd(dst + a.Co + co, s_sq);
            }
            fence_async_smem();                                // staging writes -> visible to the TMA store
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            if (it == 0 && warp == FIRST_EPI_WARP) MIDT_STAMP(6);
            if (warp == FIRST_EPI_WARP && lane == 0) {
                tma_store_4d(&tmY, stage, a.c_off + n0, 0, y0, img);
                tma_store_commit_wait_read();                  // staging (= patch buffer) may be overwritten again
                mbar_arrive(p_free);
            }
        }
        if (warp == FIRST_EPI_WARP && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
        MIDT_STAMP(7);
    }
}

std::atomic<unsigned long long> g_midt_launches{0};
unsigned long long* g_midt_trace = nullptr;   // device buffer [148][8] set by mg_conv_midt_trace (profiling only)

bool midt_disabled() {
    const char* e = std::getenv("MAGGIE_B200_NO_MIDT_CONV");
    return e && e[0] == '1';
}
bool midt_1x1_disabled() {
    const char* e = std::getenv("MAGGIE_B200_NO_MIDT_1X1");
    return e && e[0] == '1';
}
int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// issue cost (cycles) of one M = 128, K = 16 MMA with N columns, both operands in shared memory (tools/mma_bench.cu)
double mma_cycles(int n) { return std::max(37.0 + 0.31 * n, 0.55 * n); }

}  // namespace

namespace mg {

// MG_OK + *handled = true: launched here;  MG_OK + *handled = false: not eligible (K2h / generic kernel);  else error.
int conv_midt_launch(const mg_conv_desc* d, void* stream, bool* handled) {
    *handled = false;
    if (midt_disabled()) return MG_OK;
    if (d->sy != 1 || d->sx != 1 || d->oys != 1 || d->oxs != 1 || d->oy0 != 0 || d->ox0 != 0 || d->res_up) return MG_OK;
    if (d->Hg != d->Hi || d->Wg != d->Wi || d->Ho != d->Hi || d->Wo != d->Wi || d->n_phases > 1) return MG_OK;
    if (d->n_taps > 9 || (d->n_taps < 4 && d->n_taps != 1) || d->Ci % CH || d->Ci < 128 || d->Ci > CH * MAX_CHUNKS || d->Co % BM) return MG_OK;
    if (d->Wi > 64 || d->Wi < 8 || d->Hi < 4) return MG_OK;
    int hal = 0;
    for (int t = 0; t < d->n_taps; ++t) hal = std::max(hal, std::max(std::abs(d->tap_dy[t]), std::abs(d->tap_dx[t])));
    if (hal > 2 || (hal == 0) != (d->n_taps == 1)) return MG_OK;   // 1x1 layers (no halo, P = W) or taps within +-2 pixels
    if (d->n_taps == 1 && midt_1x1_disabled()) return MG_OK;
    if (d->res && (d->c_off != 0 || d->Cs != d->Co)) return MG_OK;
    if (d->Cs % 8 || d->c_off % 8) return MG_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) return MG_OK;

    TArgs a;
    a.n_taps = d->n_taps;
    a.H = d->Hi, a.W = d->Wi, a.Ci = d->Ci, a.Co = d->Co, a.hal = hal;
    a.P = d->Wi + 2 * hal;
    a.n_ntiles = d->Co / BM;
    a.chunks = d->Ci / CH;
    const int fixed = 1024 /*align*/ + 1024 /*guard*/ + (MAX_CHUNKS + 2 * MAX_NB + 4) * 8 + 64 + 512 * 4 /*column table*/;
    const int budget = 226 * 1024 - fixed;
    // slab height: cost model over R (see the header comment)
    const double ksteps = (double)d->n_taps * d->Ci / 16.0;
    const int forced_R = env_int("MAGGIE_B200_MIDT_ROWS", 0);
    int best_R = 0, max_R = 0;
    double best_cost = 1e30;
    for (int R = 1; R <= d->Hi; ++R) {
        if (forced_R && R != forced_R) continue;
        const int cols = (R * a.P + 15) & ~15;
        if (cols > 512) break;
        const int a_al = ((R + 2 * hal) * a.P * 128 + 1023) & ~1023;
        if ((budget - a.chunks * a_al) / B_BYTES < 3) break;
        max_R = R;                                                   // (the patch of R rows fits)
        if (R * d->Wi * BM * 2 > a.chunks * a_al) continue;          // output staging must fit the patch buffer
        const int rblocks = ceil_div(d->Hi, R);
        if (ceil_div(d->Hi, rblocks) != R) continue;                 // even split only
        const int items = d->N * rblocks * a.n_ntiles;
        const int waves = ceil_div(items, kNumSMs);
        const int n0 = cols <= 256 ? cols : ((cols / 2 + 15) & ~15);
        const double mma = ksteps * (mma_cycles(n0) + (cols > 256 ? mma_cycles(cols - n0) : 0.0));
        // measured (tools/midt_probe.py): an SM takes in ~39 B/clk of TMA loads (weights of the item + its patch), the
        // epilogue costs ~10 clk per accumulator column, set-up + first loads + store + exit ~10 k clk per launch
        const double load = ((double)BM * d->n_taps * d->Ci * 2 + (double)a.chunks * a_al) / 39.0;
        const double cost = waves * (std::max(mma, load) + 10.0 * cols) + 10000.0;
        if (cost < best_cost) best_cost = cost, best_R = R;
    }
    // slabs of fewer than 4 rows (the patch of a wide-Ci layer does not fit otherwise) re-stream the weights too often: the
    // generic kernel is faster there (64^2 256->128: 21 us vs 24 us)
    if (!best_R || (max_R < 4 && max_R < d->Hi && !forced_R)) return MG_OK;
    const int R = best_R;
    a.R = R;
    a.rblocks = ceil_div(d->Hi, R);
    a.n_items = d->N * a.rblocks * a.n_ntiles;
    a.cols = (R * a.P + 15) & ~15;
    if (a.cols <= 256) {
        a.nparts = 1, a.part_n[0] = a.cols, a.part_n[1] = 16;
    } else {
        a.nparts = 2, a.part_n[0] = (a.cols / 2 + 15) & ~15, a.part_n[1] = a.cols - a.part_n[0];
    }
    a.a_bytes = (R + 2 * hal) * a.P * 128;
    a.a_al = (a.a_bytes + 1023) & ~1023;
    a.NB = std::min(MAX_NB, (budget - a.chunks * a.a_al) / B_BYTES);
    a.tmem_cols = 32;
    while (a.tmem_cols < a.cols) a.tmem_cols <<= 1;
    for (int t = 0; t < d->n_taps; ++t) {
        a.tap_off[t] = (hal + d->tap_dy[t]) * a.P + d->tap_dx[t];
        a.tap_koff[t] = d->tap_koff[t];
    }
    a.pre_act = d->pre_act, a.post_act = d->post_act, a.stats = d->stats, a.bias = d->bias;
    a.scale = d->scale, a.shift = d->shift, a.res = static_cast<const __half*>(d->res);
    a.c_off = d->c_off;
    a.trace = g_midt_trace;

    CUtensorMap tmX, tmW, tmY;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->Ci * 2, (cuuint64_t)d->Wi * d->Ci * 2, (cuuint64_t)d->Hi * d->Wi * d->Ci * 2};
        cuuint32_t box[4] = {(cuuint32_t)CH, (cuuint32_t)a.P, (cuuint32_t)(R + 2 * hal), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->x), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return MG_OK;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Co};
        cuuint64_t strides[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)CH, (cuuint32_t)BM};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return MG_OK;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Cs, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->Cs * 2, (cuuint64_t)d->Wi * d->Cs * 2, (cuuint64_t)d->Hi * d->Wi * d->Cs * 2};
        cuuint32_t box[4] = {(cuuint32_t)BM, (cuuint32_t)d->Wi, (cuuint32_t)R, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d->out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return MG_OK;
    }
    const size_t smem = (size_t)fixed + (size_t)a.chunks * a.a_al + (size_t)a.NB * B_BYTES;
    const bool lean = !d->bias && !d->scale && !d->res && !d->post_act;
    const bool bias_only = d->bias && !d->scale && !d->res && !d->post_act && !d->pre_act;
    const int epi = lean ? d->pre_act : (bias_only ? 4 : 3);
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TArgs);
    static const KernelFn table[5] = {conv_midt_tcgen05_kernel<0>, conv_midt_tcgen05_kernel<1>, conv_midt_tcgen05_kernel<2>,
                                      conv_midt_tcgen05_kernel<3>, conv_midt_tcgen05_kernel<4>};
    static bool attr_set = false;
    if (!attr_set) {
        for (int j = 0; j < 5; ++j)
            if (cudaFuncSetAttribute(table[j], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
                set_error("mg_conv_fprop: cannot raise dynamic shared memory limit (transposed mid kernel)");
                return MG_ERR_CUDA;
            }
        attr_set = true;
    }
    const int grid = std::min(a.n_items, kNumSMs);
    const KernelFn fn = table[epi];
    MG_LAUNCH(fn, grid, THREADS, smem, stream, tmX, tmW, tmY, a);
    MG_CHECK_LAUNCH("mg_conv_fprop(mid, transposed)");
    g_midt_launches.fetch_add(1, std::memory_order_relaxed);
    *handled = true;
    return MG_OK;
}

}  // namespace mg

extern "C" unsigned long long mg_conv_midt_launches(void) { return g_midt_launches.load(); }

// Profiling aid (tools/midt_probe.py): K2t launches write 8 globaltimer stamps per CTA into `buf` (device, >= 148 * 8
// uint64): kernel entry, set-up + griddepcontrol.wait done, first patch chunk landed, first weight block landed, all MMAs of
// the first item issued, accumulators complete, epilogue of the first item staged, kernel exit.  NULL switches it off.
extern "C" void mg_conv_midt_trace(unsigned long long* buf) { g_midt_trace = buf; }
