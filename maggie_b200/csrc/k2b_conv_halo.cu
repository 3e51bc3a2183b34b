// K2b: halo-resident, weight-stationary, persistent variant of the dense convolution for the HIGH-RESOLUTION,
// LOW-CHANNEL layers (Ci, Co <= 64; stride 1; taps within +-1 pixel: 3x3 "same" convs and their data gradients).
// These layers are HBM-bound (32 -> 32 at 512^2: 144 FLOP/B against a ridge of ~210), and the generic kernel (K2)
// re-stages the activation tile once per tap - 9x the L2 -> SM traffic - and re-loads the weights for every tile.
//
//   * A (activations): ONE 4-D TMA box {C channels, P = Ws + 2 pixels, R + 2 rows} per tile brings a halo patch into
//     shared memory as a LINEAR array of pixels, one swizzled row of C channels (32 / 64 / 128 bytes) per pixel; the
//     hardware zero fill supplies the conv padding (coordinates start at x = -1, y = -1).  Accumulator row m of a
//     128-row MMA is patch pixel (start + m), so the operand of tap (dy, dx) is the SAME buffer read from
//     start + dy*P + dx.  The swizzle XOR is a function of the shared-memory ADDRESS, hence a descriptor whose start
//     is not on an 8-row atom boundary reads exactly what TMA stored there (verified bit-exact against K2).  No im2col,
//     no per-tap staging; the two pad columns of every patch row are computed and discarded (1.5 % of the MMAs).
//   * B (weights): all taps of the layer stay resident in shared memory for the CTA's lifetime (<= 74 KB).
//   * Persistent CTAs (one per SM) walk over the tiles; the halo buffer is double buffered and so is the accumulator
//     SET of a whole tile (all its 128-pixel blocks side by side in TMEM, 2 x <= 256 columns): ONE tcgen05.commit /
//     mbarrier round trip per tile instead of one per block (measured: a commit -> wait -> arrive -> wait round trip
//     costs ~1.3 us, which dominated everything else), so TMA, tcgen05.mma and the epilogue of consecutive tiles overlap.
//   * The single MMA-issuing thread is the critical resource (18 small MMAs per 128 pixels): operand descriptors are
//     precomputed per tap, the K loop is unrolled at compile time, one 32-bit add per MMA.
//   * With one CTA per SM the epilogue (tcgen05.ld, activation, BatchNorm partial sums, fp16 stores) is the busiest
//     role (ncu: 70 % of the stall samples with 4 epilogue warps): THREE epilogue warpgroups take the 128-pixel blocks
//     of a tile round robin, so every SM sub-partition has three epilogue warps to interleave.
//   * BatchNorm statistics are accumulated per warp over ALL tiles of the CTA and flushed with one set of atomics.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

#include <cstdlib>

namespace {

using namespace mg::ptx;

constexpr int EPI_GROUPS = 3;
constexpr int EPI_WARPS = 4 * EPI_GROUPS;
constexpr int THREADS = 64 + 32 * EPI_WARPS;   // warp 0: TMA, warp 1: MMA issue + TMEM, warps 2..13: epilogue
constexpr int SETS = 2;        // accumulator sets (one tile each) in TMEM
constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;

struct HArgs {
    int n_taps, tap_dy[9], tap_dx[9], tap_koff[9];
    int H, W, C, Co;
    int Ws, P, R, strips, rblocks, n_tiles, mblocks;
    int row_bytes, a_bytes, b_tap_bytes, acc_cols, set_cols, tmem_cols;
    int debug;   // profiling experiments only (MAGGIE_B200_HALO_DEBUG): 1 = one tap per block, 4 = no stores
    __half* out;
    int Cs, c_off;
    int pre_act, post_act;
    float* stats;
    const float *bias, *scale, *shift;
    const __half* res;
};

__device__ __forceinline__ float act_apply(float v, int act) {
    return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v > 0.f ? v : 0.2f * v) : v);
}

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// KS = C / 16: UMMA K steps per tap.  EPI: epilogue specialisation - 0 / 1 / 2 = lean (no bias / affine / residual /
// post-activation; pre-activation none / ReLU / LeakyReLU fixed at compile time: training forward and data gradients),
// 3 = general (everything decided at run time: eval-mode fused BatchNorm, residual, bias).
template <int KS, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
conv_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HArgs a) {
    mg::pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // [1 KB guard][A0][A1][4 KB guard][B taps][barriers][tmem slot][epilogue staging][stat partials]
    // (guards: the first / last 128-row block of a patch reads up to one pixel before / 127 pixels past it; those
    //  accumulator rows are pad columns or rows beyond the tile and are never stored)
    const int a_al = (a.a_bytes + 1023) & ~1023;
    uint8_t* sA = smem + 1024;
    uint8_t* sB = sA + 2 * a_al + 4096;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + a.n_taps * a.b_tap_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    float* s_stage = reinterpret_cast<float*>(tmem_slot + 4);   // [EPI_WARPS][16 columns][36]: transposition buffer
    float* s_part = s_stage + EPI_WARPS * 16 * 36;              // [EPI_WARPS][2][Co]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t a_full = bar0, a_empty = bar0 + 16, b_full = bar0 + 32, t_full = bar0 + 40, t_empty = t_full + 8 * SETS;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int i = 0; i < 2; ++i) mbar_init(a_full + 8 * i, 1), mbar_init(a_empty + 8 * i, 1);
        mbar_init(b_full, 1);
        for (int i = 0; i < SETS; ++i) mbar_init(t_full + 8 * i, 1), mbar_init(t_empty + 8 * i, EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    for (int i = threadIdx.x; i < EPI_WARPS * 2 * a.Co; i += THREADS) s_part[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mg::pdl_wait();   // set-up overlapped with the previous kernel's tail; its data is needed from here on
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_img = a.rblocks * a.strips;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: the weights once, then one halo patch per tile =====
            mbar_expect_tx(b_full, a.n_taps * a.b_tap_bytes);
            for (int t = 0; t < a.n_taps; ++t) tma_load_2d(smem_u32(sB + t * a.b_tap_bytes), &tmB, b_full, a.tap_koff[t], 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1, ph = (it >> 1) & 1;
                const int img = tile / tiles_img, rem = tile - img * tiles_img;
                const int rb = rem / a.strips, st = rem - rb * a.strips;
                mbar_wait(a_empty + 8 * buf, ph ^ 1);
                mbar_expect_tx(a_full + 8 * buf, a.a_bytes);
                tma_load_4d(smem_u32(sA + buf * a_al), &tmA, a_full + 8 * buf, 0, st * a.Ws - 1, rb * a.R - 1, img);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop with warp-uniform operands (uniform registers; otherwise ptxas wraps
        // EVERY tcgen05.mma in an elect / R2UR.BROADCAST waterfall of ~100 cycles - 2-3x the tensor time of these small
        // MMAs, measured with tools/mma_bench.cu), one elected lane issues =====
        const uint32_t tmem_u = uniform_u32(tmem_base);
        const uint32_t idesc = instr_desc_f16(128, a.Co, 0, 0);
        const uint32_t lay = swizzle_layout(a.row_bytes), sbo = 8 * a.row_bytes;   // A and B rows are both C*2 bytes
        const int ntap = (a.debug & 1) ? 1 : a.n_taps;
        const uint32_t mb_step = (128 * a.row_bytes) >> 4;
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        mbar_wait(b_full, 0);
        int buf = 0, ph = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            mbar_wait(t_empty + 8 * buf, ph ^ 1);   // halo buffer and accumulator set alternate together
            mbar_wait(a_full + 8 * buf, ph);
            tc_fence_after();
            if (elect_one()) {
                uint64_t a_desc = smem_desc(sA_u + buf * a_al, 0, sbo, lay);
                uint32_t d_tmem = tmem_u + buf * a.set_cols;
                for (int mb = 0; mb < a.mblocks; ++mb, a_desc += mb_step, d_tmem += a.acc_cols) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        if (t < ntap) {
                            // per-tap constants straight from the kernel parameters (compile-time indices): A start offset
                            // in 16-byte units relative to the patch (may be -1 pixel), B descriptor of the resident tap
                            const int a_off = (((1 + a.tap_dy[t]) * a.P + a.tap_dx[t]) * a.row_bytes) >> 4;
                            const uint64_t da = a_desc + (int64_t)a_off;
                            const uint64_t db = smem_desc(sB_u + t * a.b_tap_bytes, 0, sbo, lay);
#pragma unroll
                            for (int k = 0; k < KS; ++k) mma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (t | k) != 0);
                        }
                    }
                }
                mma_commit(t_full + 8 * buf);    // the whole tile's accumulators are complete ...
                mma_commit(a_empty + 8 * buf);   // ... and its halo patch can be overwritten
            }
            if (buf) ph ^= 1;
            buf ^= 1;
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3, ew = warp - 2, grp = ew >> 2;   // TMEM lane quarter, epilogue warp, warpgroup
        const uint32_t stg = smem_u32(s_stage + ew * 16 * 36);
        float* part = s_part + ew * 2 * a.Co;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int img = tile / tiles_img, rem = tile - img * tiles_img;
            const int rb = rem / a.strips, st = rem - rb * a.strips;
            const int y0 = rb * a.R, xs = st * a.Ws;
            const int buf = it & 1, ph = (it >> 1) & 1;
            mbar_wait(t_full + 8 * buf, ph);
            tc_fence_after();
            for (int mb = grp; mb < a.mblocks; mb += EPI_GROUPS) {
                const int m = mb * 128 + q * 32 + lane;
                const int row = m / a.P, j = m - row * a.P;
                const int y = y0 + row, x = xs + j - 1;
                const bool valid = row < a.R && j >= 1 && j <= a.Ws && x < a.W && y < a.H;
                const size_t pix = ((size_t)img * a.H + y) * a.W + x;
                __half* orow = a.out + pix * a.Cs + a.c_off;
                const __half* rrow = a.res ? a.res + pix * a.Co : nullptr;
                for (int c0 = 0; c0 < a.Co; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * a.set_cols + mb * a.acc_cols + c0, r);
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i] = __uint_as_float(r[i]);
                        if constexpr (EPI == 3) {
                            if (a.bias) v[i] += __ldg(a.bias + c0 + i);
                            v[i] = act_apply(v[i], a.pre_act);
                        } else if constexpr (EPI == 1) {
                            v[i] = fmaxf(v[i], 0.f);
                        } else if constexpr (EPI == 2) {
                            v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
                        }
                    }
                    if (a.stats) {
                        // per-channel sum / sum of squares over this warp's 32 rows: transpose through shared memory
                        // ([column][row], pitch 36: conflict-free scalar stores, 16-byte loads), lanes 0..15 sum column
                        // `lane`, lanes 16..31 sum its squares
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
                        part[(lane >> 4) * a.Co + c0 + col] += acc;   // one owner lane per entry: no race
                    }
                    if (valid && !(a.debug & 4)) {
                        if constexpr (EPI == 3) {
                        if (a.scale) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], __ldg(a.scale + c0 + i), __ldg(a.shift + c0 + i));
                        }
                        if (rrow) {
                            const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rrow + c0)), r1 = __ldg(reinterpret_cast<const uint4*>(rrow + c0) + 1);
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
            if (lane == 0) mbar_arrive(t_empty + 8 * buf);   // this warp's quarter of the accumulator set is free again
        }
        if (a.stats) {
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");  // the epilogue warps only
            const int et = threadIdx.x - 64;
            float* dst = a.stats + (size_t)(blockIdx.x % STAT_COPIES) * 2 * a.Co;
            for (int i = et; i < 2 * a.Co; i += 32 * EPI_WARPS) {
                float tot = 0.f;
#pragma unroll
                for (int w = 0; w < EPI_WARPS; ++w) tot += s_part[w * 2 * a.Co + i];
                atomicAdd(dst + i, tot);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

std::atomic<unsigned long long> g_halo_launches{0};

bool halo_disabled() {   // read on every call (a getenv is ~100 ns): tests and A/B timings flip it at run time
    const char* e = std::getenv("MAGGIE_B200_NO_HALO_CONV");
    return e && e[0] == '1';
}

}  // namespace

namespace mg {

// Returns MG_OK with *handled = true when the layer was launched on the halo-resident kernel, MG_OK with
// *handled = false when it is not eligible (the caller then uses the generic kernel), or an error code.
int conv_halo_launch(const mg_conv_desc* d, void* stream, bool* handled) {
    *handled = false;
    if (halo_disabled()) return MG_OK;
    if (d->sy != 1 || d->sx != 1 || d->oys != 1 || d->oxs != 1 || d->oy0 != 0 || d->ox0 != 0 || d->res_up) return MG_OK;
    if (d->Hg != d->Hi || d->Wg != d->Wi || d->Ho != d->Hi || d->Wo != d->Wi) return MG_OK;
    // (16-channel operands - 32-byte swizzle - do not survive the shifted descriptor start: left to the generic kernel)
    if (d->n_taps > 9 || (d->Ci != 32 && d->Ci != 64) || (d->Co != 32 && d->Co != 64)) return MG_OK;
    if (d->Wi < 64 || d->Hi < 8) return MG_OK;
    for (int t = 0; t < d->n_taps; ++t)
        if (d->tap_dy[t] < -1 || d->tap_dy[t] > 1 || d->tap_dx[t] < -1 || d->tap_dx[t] > 1) return MG_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) return MG_OK;

    HArgs a;
    a.n_taps = d->n_taps;
    for (int t = 0; t < d->n_taps; ++t) a.tap_dy[t] = d->tap_dy[t], a.tap_dx[t] = d->tap_dx[t], a.tap_koff[t] = d->tap_koff[t];
    a.H = d->Hi, a.W = d->Wi, a.C = d->Ci, a.Co = d->Co;
    a.Ws = d->Wi <= 254 ? d->Wi : 128;
    a.P = a.Ws + 2;
    a.strips = ceil_div(d->Wi, a.Ws);
    a.b_tap_bytes = d->Co * d->Ci * 2;
    if (a.b_tap_bytes % 1024) return MG_OK;               // swizzle atoms of consecutive taps must stay 1 KB aligned
    a.acc_cols = d->Co;
    const int fixed = 1024 /*align*/ + 1024 + 4096 + a.n_taps * a.b_tap_bytes + 256 + EPI_WARPS * 16 * 36 * 4 + EPI_WARPS * 2 * d->Co * 4 + 1024;
    const int budget = 224 * 1024 - fixed;
    int R = 0;
    a.row_bytes = d->Ci * 2;                              // 32 / 64 / 128-byte swizzled rows (activation pixels, weight rows)
    auto patch_bytes = [&](int r) { return (r + 2) * a.P * a.row_bytes; };
    for (int r = 16; r >= 1; --r) {   // largest row block whose two patches fit and whose accumulator set is <= 256 columns
        const int ab = (patch_bytes(r) + 1023) & ~1023;
        if (2 * ab <= budget && ceil_div(r * a.P, 128) * a.acc_cols <= 256) { R = r; break; }
    }
    if (R == 0) return MG_OK;
    R = std::min(R, d->Hi);
    a.R = R;
    a.rblocks = ceil_div(d->Hi, R);
    a.n_tiles = d->N * a.rblocks * a.strips;
    a.mblocks = ceil_div(R * a.P, 128);
    a.a_bytes = patch_bytes(R);
    a.set_cols = a.mblocks * a.acc_cols;
    a.tmem_cols = 32;
    while (a.tmem_cols < SETS * a.set_cols) a.tmem_cols <<= 1;
    {
        const char* e = std::getenv("MAGGIE_B200_HALO_DEBUG");
        a.debug = e ? std::atoi(e) : 0;
    }
    a.out = static_cast<__half*>(d->out);
    a.Cs = d->Cs, a.c_off = d->c_off;
    a.pre_act = d->pre_act, a.post_act = d->post_act, a.stats = d->stats, a.bias = d->bias;
    a.scale = d->scale, a.shift = d->shift, a.res = static_cast<const __half*>(d->res);
    if (a.res && (d->c_off != 0 || d->Cs != d->Co)) return MG_OK;

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->Ci * 2, (cuuint64_t)d->Wi * d->Ci * 2, (cuuint64_t)d->Hi * d->Wi * d->Ci * 2};
        cuuint32_t box[4] = {(cuuint32_t)d->Ci, (cuuint32_t)a.P, (cuuint32_t)(R + 2), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(a.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return MG_OK;              // not expressible: let the generic kernel take it
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Co};
        cuuint64_t strides[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)d->Ci, (cuuint32_t)d->Co};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(a.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return MG_OK;
    }
    const size_t smem = (size_t)fixed + 2 * (size_t)((a.a_bytes + 1023) & ~1023);
    const bool lean = !d->bias && !d->scale && !d->res && !d->post_act;
    const int epi = lean ? d->pre_act : 3;
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const HArgs);
    static const KernelFn table[2][4] = {
        {conv_halo_tcgen05_kernel<2, 0>, conv_halo_tcgen05_kernel<2, 1>, conv_halo_tcgen05_kernel<2, 2>, conv_halo_tcgen05_kernel<2, 3>},
        {conv_halo_tcgen05_kernel<4, 0>, conv_halo_tcgen05_kernel<4, 1>, conv_halo_tcgen05_kernel<4, 2>, conv_halo_tcgen05_kernel<4, 3>}};
    static bool attr_set = false;
    if (!attr_set) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 4; ++j)
                if (cudaFuncSetAttribute(table[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) {
                    set_error("mg_conv_fprop: cannot raise dynamic shared memory limit (halo kernel)");
                    return MG_ERR_CUDA;
                }
        attr_set = true;
    }
    const int grid = std::min(a.n_tiles, kNumSMs);
    const KernelFn fn = table[d->Ci == 64 ? 1 : 0][epi];
    MG_LAUNCH(fn, grid, THREADS, smem, stream, tmA, tmB, a);
    MG_CHECK_LAUNCH("mg_conv_fprop(halo)");
    g_halo_launches.fetch_add(1, std::memory_order_relaxed);
    *handled = true;
    return MG_OK;
}

}  // namespace mg

extern "C" unsigned long long mg_conv_halo_launches(void) { return g_halo_launches.load(); }
