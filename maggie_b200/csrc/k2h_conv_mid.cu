// K2h: halo-patch, weight-streaming, persistent variant of the dense convolution for the MID-RESOLUTION layers
// (Ci in {128, 256, 512}, width <= 64; stride 1; taps within +-1 pixel: the 3x3 "same" convs of the encoder trunk at 64^2 /
// 32^2 / 16^2, of the dense decoder, and their data gradients - 25 of the 62 conv layers and most of the conv FLOPs).
//
// Why: the generic kernel (K2) stages a [128 pixels x 64 channels] activation tile AND a [128 x 64] weight tile for every
// (tap, 64-channel chunk) of every 128 x 128 output tile: 64 FLOP per staged byte.  At the ~12 TB/s the L2 can deliver
// to the SMs that caps the tensor pipe at ~55 % and K2 measured 395 TFLOP/s (24.4 us) on 64^2 128->128 (ncu: tensor pipe
// 20 %, L2 -> SM traffic 151 MB for 8.4 MB of unique operands).  Here
//   * A (activations): per channel chunk ONE 4-D TMA box {CH channels, P = W + 2 pixels, R + 2 rows} brings a halo patch of
//     a whole R-row slab of the image into shared memory as a LINEAR array of pixels (hardware zero fill = conv padding).
//     Accumulator row m of a 128-row MMA block is patch pixel (start + m); the operand of tap (dy, dx) is the SAME buffer
//     read from start + dy*P + dx (the swizzle XOR is a function of the shared-memory address, so a descriptor that does
//     not start on an 8-row atom boundary reads what TMA stored - as in K2b).  All 9 taps and all M blocks of the slab
//     reuse the patch: activation traffic drops ~7x.
//   * B (weights): [BN x CH] blocks of one (tap, chunk) stream through their own ring; each block feeds ALL M blocks of the
//     slab (up to 5 x 128 pixels), so weight traffic per FLOP drops ~5x.  ~200-300 FLOP per staged byte in total.
//   * one CTA per SM, persistent over (slab, BN-column tile) work items; accumulators of all M blocks of an item side by
//     side in TMEM (<= 512 columns; two sets when they fit, so the epilogue of item i overlaps the MMAs of item i + 1);
//     separate producer threads for the A and the B ring (an A patch is 40 - 85 KB: it must be requested chunks ahead,
//     independently of the fine-grained weight ring).
// The two pad columns of every patch row and the rows of the last block beyond the slab are computed and discarded.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

#include <cstdlib>

namespace {

using namespace mg::ptx;

constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;                         // warp 0: A producer, 1: MMA issue + TMEM, 2: B producer, 3: idle
constexpr int THREADS = 32 * (FIRST_EPI_WARP + EPI_WARPS);
constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;
constexpr int MAX_NA = 4, MAX_NB = 12;

struct MArgs {
    int n_taps, tap_dy[9], tap_dx[9], tap_koff[9];
    int H, W, Ci, Co;
    int P, R, rblocks, mblocks, n_ptiles, n_ntiles, n_items;
    int CH, chunks, BN, NA, NB;                            // channels per stage (32 | 64), Ci / CH, N tile, ring depths
    int row_bytes, a_bytes, a_al, b_bytes, set_cols, sets, tmem_cols;
    __half* out;
    int Cs, c_off;
    int pre_act, post_act;
    float* stats;
    const float *bias, *scale, *shift;
    const __half* res;
    unsigned long long* trace;   // profiling only (mg_conv_mid_trace): per CTA 8 globaltimer stamps, see the kernel
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define MID_STAMP(slot)                                                          \
    do {                                                                         \
        if (a.trace && lane == 0) a.trace[blockIdx.x * 8 + (slot)] = gtime();    \
    } while (0)

__device__ __forceinline__ float act_apply(float v, int act) {
    return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v > 0.f ? v : 0.2f * v) : v);
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// KS = CH / 16: UMMA K steps per (tap, chunk) stage.  EPI: 0 / 1 / 2 = lean epilogue (pre-activation none / ReLU /
// LeakyReLU fixed at compile time: training forward and data gradients), 3 = general (bias, eval BN affine, residual, ...).
template <int KS, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
conv_mid_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const MArgs a) {
    mg::pdl_launch();
    if (a.trace && threadIdx.x == 32) a.trace[blockIdx.x * 8 + 0] = gtime();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // [1 KB guard][A ring][B ring][barriers][tmem slot][epilogue staging][stat partials]
    // (the first M block of a patch reads one pixel before it, the last one up to 128 pixels past it - into the next A
    //  stage or the B ring: those accumulator rows are pad columns / rows beyond the slab and are never stored)
    uint8_t* sA = smem + 1024;
    uint8_t* sB = sA + a.NA * a.a_al;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + a.NB * a.b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_NA + 2 * MAX_NB + 4);
    float* s_stage = reinterpret_cast<float*>(tmem_slot + 4);   // [EPI_WARPS][16 columns][36]: transposition buffer
    float* s_part = s_stage + EPI_WARPS * 16 * 36;              // [EPI_WARPS][2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t a_full = bar0, a_empty = a_full + 8 * MAX_NA, b_full = a_empty + 8 * MAX_NA, b_empty = b_full + 8 * MAX_NB,
                   t_full = b_empty + 8 * MAX_NB, t_empty = t_full + 16;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int i = 0; i < a.NA; ++i) mbar_init(a_full + 8 * i, 1), mbar_init(a_empty + 8 * i, 1);
        for (int i = 0; i < a.NB; ++i) mbar_init(b_full + 8 * i, 1), mbar_init(b_empty + 8 * i, 1);
        for (int i = 0; i < 2; ++i) mbar_init(t_full + 8 * i, 1), mbar_init(t_empty + 8 * i, EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    for (int i = threadIdx.x; i < EPI_WARPS * 2 * a.BN; i += THREADS) s_part[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mg::pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) MID_STAMP(1);

    if (warp == 0) {
        if (lane == 0) {
            // ===== A producer: one halo patch per (item, channel chunk) =====
            int s = 0, ph = 0;
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
                const int pt = item / a.n_ntiles;
                const int img = pt / a.rblocks, rb = pt - img * a.rblocks;
                for (int c = 0; c < a.chunks; ++c) {
                    mbar_wait(a_empty + 8 * s, ph ^ 1);
                    mbar_expect_tx(a_full + 8 * s, a.a_bytes);
                    tma_load_4d(smem_u32(sA + s * a.a_al), &tmA, a_full + 8 * s, c * a.CH, -1, rb * a.R - 1, img);
                    if (++s == a.NA) s = 0, ph ^= 1;
                }
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {
            // ===== B producer: one [BN x CH] weight block per (item, chunk, tap) =====
            int s = 0, ph = 0;
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
                const int n0 = (item % a.n_ntiles) * a.BN;
                for (int c = 0; c < a.chunks; ++c) {
                    for (int t = 0; t < a.n_taps; ++t) {
                        mbar_wait(b_empty + 8 * s, ph ^ 1);
                        mbar_expect_tx(b_full + 8 * s, a.b_bytes);
                        tma_load_2d(smem_u32(sB + s * a.b_bytes), &tmB, b_full + 8 * s, a.tap_koff[t] + c * a.CH, n0);
                        if (++s == a.NB) s = 0, ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // Every operand of tcgen05.mma must sit in UNIFORM registers.  Values the compiler cannot prove warp-uniform (the
        // TMEM base read from shared memory, an offset table indexed at run time) make it wrap EVERY MMA in an
        // elect / R2UR.BROADCAST waterfall loop (~100 cycles per MMA: 3x the tensor time of an N = 64 MMA; measured with
        // tools/mma_bench.cu and visible in the SASS).  Hence: the TMEM base goes through a shuffle (uniform by
        // construction), the tap offsets are compile-time-unrolled reads of kernel parameters, ring indices are
        // counters (no run-time modulo), and the whole warp runs the loop with one elected lane issuing.
        const uint32_t tmem_u = uniform_u32(tmem_base);
        const uint32_t idesc = instr_desc_f16(128, a.BN, 0, 0);
        const uint32_t lay = swizzle_layout(a.row_bytes), sbo = 8 * a.row_bytes;
        const uint32_t mb_step = (128 * a.row_bytes) >> 4;
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        int sa = 0, pa = 0, sb = 0, pb = 0, set = 0, tph = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            mbar_wait(t_empty + 8 * set, tph ^ 1);       // (all lanes poll: no divergence in this warp)
            tc_fence_after();
            const uint32_t d0 = tmem_u + set * a.set_cols;
            for (int c = 0; c < a.chunks; ++c) {
                mbar_wait(a_full + 8 * sa, pa);
                tc_fence_after();
                if (item == (int)blockIdx.x && c == 0) MID_STAMP(2);
                const uint64_t a_desc0 = smem_desc(sA_u + sa * a.a_al, 0, sbo, lay);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    if (t < a.n_taps) {
                        mbar_wait(b_full + 8 * sb, pb);
                        tc_fence_after();
                        if (item == (int)blockIdx.x && c == 0 && t == 0) MID_STAMP(3);
                        const int a_off = (((1 + a.tap_dy[t]) * a.P + a.tap_dx[t]) * a.row_bytes) >> 4;
                        const uint64_t b_desc = smem_desc(sB_u + sb * a.b_bytes, 0, sbo, lay);
                        uint64_t da = a_desc0 + (int64_t)a_off;
                        uint32_t d_tmem = d0;
                        const uint32_t first = (c | t) == 0 ? 0u : 1u;
                        for (int mb = 0; mb < a.mblocks; ++mb, da += mb_step, d_tmem += a.BN) {
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < KS; ++k) mma_f16(d_tmem, da + 2 * k, b_desc + 2 * k, idesc, first | (uint32_t)k);
                            }
                        }
                        if (elect_one()) mma_commit(b_empty + 8 * sb);   // this weight block is free when the MMAs above retire
                        if (++sb == a.NB) sb = 0, pb ^= 1;
                    }
                }
                if (elect_one()) mma_commit(a_empty + 8 * sa);           // ... and so is the patch after its last tap
                if (++sa == a.NA) sa = 0, pa ^= 1;
            }
            if (elect_one()) mma_commit(t_full + 8 * set);               // all accumulators of the item are complete
            if (item == (int)blockIdx.x) MID_STAMP(4);
            if (a.sets == 2) {
                if (set) tph ^= 1;
                set ^= 1;
            } else {
                tph ^= 1;
            }
        }
    } else if (warp >= FIRST_EPI_WARP) {
        // ===== epilogue: warps 4..11, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3, ew = warp - FIRST_EPI_WARP, grp = ew >> 2;
        constexpr int EPI_GROUPS = EPI_WARPS / 4;
        const uint32_t stg = smem_u32(s_stage + ew * 16 * 36);
        float* part = s_part + ew * 2 * a.BN;
        int it = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
            const int pt = item / a.n_ntiles, n0 = (item - pt * a.n_ntiles) * a.BN;
            const int img = pt / a.rblocks, rb = pt - img * a.rblocks;
            const int y0 = rb * a.R;
            const int set = a.sets == 2 ? (it & 1) : 0;
            const int tph = a.sets == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(t_full + 8 * set, tph);
            tc_fence_after();
            if (item == (int)blockIdx.x && warp == FIRST_EPI_WARP) MID_STAMP(5);
            for (int mb = grp; mb < a.mblocks; mb += EPI_GROUPS) {
                const int m = mb * 128 + q * 32 + lane;
                const int row = m / a.P, j = m - row * a.P;
                const int y = y0 + row, x = j - 1;
                const bool valid = row < a.R && j >= 1 && j <= a.W && y < a.H;
                const size_t pix = ((size_t)img * a.H + y) * a.W + x;
                __half* orow = a.out + pix * a.Cs + a.c_off + n0;
                const __half* rrow = a.res ? a.res + pix * a.Co + n0 : nullptr;
                for (int c0 = 0; c0 < a.BN; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + set * a.set_cols + mb * a.BN + c0, r);
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i] = __uint_as_float(r[i]);
                        if constexpr (EPI == 3) {
                            if (a.bias) v[i] += __ldg(a.bias + n0 + c0 + i);
                            v[i] = act_apply(v[i], a.pre_act);
                        } else if constexpr (EPI == 1) {
                            v[i] = fmaxf(v[i], 0.f);
                        } else if constexpr (EPI == 2) {
                            v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
                        }
                    }
                    if (a.stats) {
                        // per-channel sum / sum of squares over this warp's 32 rows (transposed through shared memory)
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 16; ++i) sts_f32(stg + (i * 36 + lane) * 4, valid ? v[i] : 0.f);
                        __syncwarp();
                        const int col = lane & 15;
                        float acc = 0.f;
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float4 z = lds_f32x4(stg + (col * 36 + 4 * jj) * 4);
                            acc += lane < 16 ? (z.x + z.y) + (z.z + z.w) : (z.x * z.x + z.y * z.y) + (z.z * z.z + z.w * z.w);
                        }
                        part[(lane >> 4) * a.BN + c0 + col] += acc;   // one owner lane per entry
                    }
                    if (valid) {
                        if constexpr (EPI == 3) {
                            if (a.scale) {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    v[i] = fmaf(v[i], __ldg(a.scale + n0 + c0 + i), __ldg(a.shift + n0 + c0 + i));
                            }
                            if (rrow) {
                                const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rrow + c0)),
                                            r1 = __ldg(reinterpret_cast<const uint4*>(rrow + c0) + 1);
                                const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[i]));
                                    v[2 * i] += f.x, v[2 * i + 1] += f.y;
                                }
                            }
                            if (a.post_act) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = act_apply(v[i], a.post_act);
                            }
                        }
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                            o[i] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        reinterpret_cast<uint4*>(orow + c0)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        reinterpret_cast<uint4*>(orow + c0)[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (item == (int)blockIdx.x && warp == FIRST_EPI_WARP) MID_STAMP(6);
            if (lane == 0) mbar_arrive(t_empty + 8 * set);
            if (a.stats && a.n_ntiles > 1) {
                // the partial sums belong to this item's channel tile: flush them before the next item changes n0
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                const int et = threadIdx.x - 32 * FIRST_EPI_WARP;
                float* dst = a.stats + (size_t)(blockIdx.x % STAT_COPIES) * 2 * a.Co;
                for (int i = et; i < 2 * a.BN; i += 32 * EPI_WARPS) {
                    const int kind = i / a.BN, cc = i - kind * a.BN;
                    float tot = 0.f;
#pragma unroll
                    for (int w = 0; w < EPI_WARPS; ++w) tot += s_part[w * 2 * a.BN + i], s_part[w * 2 * a.BN + i] = 0.f;
                    atomicAdd(dst + kind * a.Co + n0 + cc, tot);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            }
        }
        if (a.stats && a.n_ntiles == 1) {
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            const int et = threadIdx.x - 32 * FIRST_EPI_WARP;
            float* dst = a.stats + (size_t)(blockIdx.x % STAT_COPIES) * 2 * a.Co;
            for (int i = et; i < 2 * a.BN; i += 32 * EPI_WARPS) {
                float tot = 0.f;
#pragma unroll
                for (int w = 0; w < EPI_WARPS; ++w) tot += s_part[w * 2 * a.BN + i];
                atomicAdd(dst + i, tot);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
        MID_STAMP(7);
    }
}

std::atomic<unsigned long long> g_mid_launches{0};
unsigned long long* g_mid_trace = nullptr;    // device buffer [148][8] set by mg_conv_mid_trace (profiling only)

// Opt-in since K2t: after the warp-uniform MMA issue fix the generic kernel K2 (two co-resident CTAs per SM) is faster than
// K2h on every mid-resolution layer (64^2 128->128: 14.0 vs 19.8 us, 64^2 256->128: 21.2 vs 32.0 us, CUDA-graph replay), and
// K2t is faster than both where it applies.  MAGGIE_B200_MID_CONV=h routes the K2h-eligible layers K2t does not take here.
bool mid_disabled() {
    const char* e = std::getenv("MAGGIE_B200_MID_CONV");
    return !(e && e[0] == 'h');
}

int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

}  // namespace

namespace mg {

// MG_OK + *handled = true: launched here;  MG_OK + *handled = false: not eligible (generic kernel);  else error.
int conv_mid_launch(const mg_conv_desc* d, void* stream, bool* handled) {
    *handled = false;
    if (mid_disabled()) return MG_OK;
    if (d->sy != 1 || d->sx != 1 || d->oys != 1 || d->oxs != 1 || d->oy0 != 0 || d->ox0 != 0 || d->res_up) return MG_OK;
    if (d->Hg != d->Hi || d->Wg != d->Wi || d->Ho != d->Hi || d->Wo != d->Wi || d->n_phases > 1) return MG_OK;
    if (d->n_taps > 9 || d->Ci % 64 || d->Ci < 128 || d->Co % 64 || d->Wi > 64 || d->Wi < 8 || d->Hi < 4) return MG_OK;
    for (int t = 0; t < d->n_taps; ++t)
        if (d->tap_dy[t] < -1 || d->tap_dy[t] > 1 || d->tap_dx[t] < -1 || d->tap_dx[t] > 1) return MG_OK;
    if (d->n_taps < 4) return MG_OK;          // 1x1 layers have no tap reuse to exploit: generic kernel
    if (d->res && (d->c_off != 0 || d->Cs != d->Co)) return MG_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) return MG_OK;

    MArgs a;
    a.n_taps = d->n_taps;
    for (int t = 0; t < d->n_taps; ++t) a.tap_dy[t] = d->tap_dy[t], a.tap_dx[t] = d->tap_dx[t], a.tap_koff[t] = d->tap_koff[t];
    a.H = d->Hi, a.W = d->Wi, a.Ci = d->Ci, a.Co = d->Co;
    a.P = d->Wi + 2;
    a.BN = 64;
    a.n_ntiles = d->Co / a.BN;
    // slab height: the largest R with <= 5 accumulator blocks (5 x 64 columns fit one TMEM set) that still yields enough
    // work items to occupy the chip; rows are split evenly over the slabs
    const int max_blocks = std::max(1, std::min(5, env_int("MAGGIE_B200_MID_BLOCKS", 5)));
    int R = std::min(d->Hi, (max_blocks * 128) / a.P);
    while (R > 1 && d->N * ceil_div(d->Hi, R) * a.n_ntiles < kNumSMs * 3 / 4 && ceil_div(R, 2) * a.P > 128) R = ceil_div(R, 2);
    a.rblocks = ceil_div(d->Hi, R);
    R = ceil_div(d->Hi, a.rblocks);
    a.R = R;
    a.mblocks = ceil_div(R * a.P, 128);
    a.n_ptiles = d->N * a.rblocks;
    a.n_items = a.n_ptiles * a.n_ntiles;
    a.set_cols = a.mblocks * a.BN;
    a.sets = 2 * a.set_cols <= 512 ? 2 : 1;
    a.tmem_cols = 32;
    while (a.tmem_cols < a.sets * a.set_cols) a.tmem_cols <<= 1;
    if (a.tmem_cols > 512) return MG_OK;

    const int fixed = 1024 /*align*/ + 1024 /*guard*/ + (2 * MAX_NA + 2 * MAX_NB + 4) * 8 + 16 + EPI_WARPS * 16 * 36 * 4 +
                      EPI_WARPS * 2 * a.BN * 4 + 256;
    const int budget = 226 * 1024 - fixed;
    // stage granularity: 64 channels (128-byte swizzled rows) when two patches + >= 4 weight blocks fit, else 32 channels
    bool ok = false;
    for (int ch : {64, 32}) {
        if (env_int("MAGGIE_B200_MID_CH", ch) != ch) continue;
        a.CH = ch, a.row_bytes = ch * 2, a.chunks = d->Ci / ch;
        a.a_bytes = (R + 2) * a.P * a.row_bytes;
        a.a_al = (a.a_bytes + 1023) & ~1023;
        a.b_bytes = a.BN * a.row_bytes;
        for (int na = std::min(MAX_NA, std::min(a.chunks + 1, ch == 64 ? 3 : 4)); na >= 2 && !ok; --na) {
            const int left = budget - na * a.a_al;
            // the last patch's over-read (<= 128 pixel rows) must stay inside the allocation: it lands in the B ring
            const int nb = std::min(MAX_NB, left / a.b_bytes);
            if (nb >= 4 && nb * a.b_bytes >= 128 * a.row_bytes) a.NA = na, a.NB = nb, ok = true;
        }
        if (ok) break;
    }
    if (!ok) return MG_OK;
    a.out = static_cast<__half*>(d->out);
    a.Cs = d->Cs, a.c_off = d->c_off;
    a.pre_act = d->pre_act, a.post_act = d->post_act, a.stats = d->stats, a.bias = d->bias;
    a.scale = d->scale, a.shift = d->shift, a.res = static_cast<const __half*>(d->res);
    a.trace = g_mid_trace;

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->Ci * 2, (cuuint64_t)d->Wi * d->Ci * 2, (cuuint64_t)d->Hi * d->Wi * d->Ci * 2};
        cuuint32_t box[4] = {(cuuint32_t)a.CH, (cuuint32_t)a.P, (cuuint32_t)(R + 2), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(a.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return MG_OK;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Co};
        cuuint64_t strides[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)a.CH, (cuuint32_t)a.BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(a.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return MG_OK;
    }
    const size_t smem = (size_t)fixed + (size_t)a.NA * a.a_al + (size_t)a.NB * a.b_bytes;
    const bool lean = !d->bias && !d->scale && !d->res && !d->post_act;
    const int epi = lean ? d->pre_act : 3;
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const MArgs);
    static const KernelFn table[2][4] = {
        {conv_mid_tcgen05_kernel<2, 0>, conv_mid_tcgen05_kernel<2, 1>, conv_mid_tcgen05_kernel<2, 2>, conv_mid_tcgen05_kernel<2, 3>},
        {conv_mid_tcgen05_kernel<4, 0>, conv_mid_tcgen05_kernel<4, 1>, conv_mid_tcgen05_kernel<4, 2>, conv_mid_tcgen05_kernel<4, 3>}};
    static bool attr_set = false;
    if (!attr_set) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 4; ++j)
                if (cudaFuncSetAttribute(table[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
                    set_error("mg_conv_fprop: cannot raise dynamic shared memory limit (mid kernel)");
                    return MG_ERR_CUDA;
                }
        attr_set = true;
    }
    const int grid = std::min(a.n_items, kNumSMs);
    const KernelFn fn = table[a.CH == 64 ? 1 : 0][epi];
    MG_LAUNCH(fn, grid, THREADS, smem, stream, tmA, tmB, a);
    MG_CHECK_LAUNCH("mg_conv_fprop(mid)");
    g_mid_launches.fetch_add(1, std::memory_order_relaxed);
    *handled = true;
    return MG_OK;
}

}  // namespace mg

extern "C" unsigned long long mg_conv_mid_launches(void) { return g_mid_launches.load(); }

// Profiling aid (tools/mid_probe.py): K2h launches write 8 globaltimer stamps per CTA into `buf` (device, >= 148 * 8
// uint64): kernel entry, after griddepcontrol.wait, first patch landed, first weight block landed, all MMAs of the first
// item issued, accumulators complete, epilogue of the first item done, kernel exit.  NULL switches it off.
extern "C" void mg_conv_mid_trace(unsigned long long* buf) { g_mid_trace = buf; }
