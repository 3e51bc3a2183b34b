// K2s (EXPERIMENTAL, opt-in, not yet validated on hardware: taken only when the caller passes a split-K workspace, which
// the host side does only under MAGGIE_B200_CONV_SPLITK=1): split-K variant of K2 for the layers whose launch has fewer
// CTAs than the GPU has CTA slots.
//
// Why: K2 on the low-resolution layers is bound by the LATENCY of its staging chain, not by bandwidth or math
// (profiles/README.md): one CTA streams the whole K range of its 128 x BN tile (36 - 72 k-blocks of 32 KB at 32^2 / 16^2)
// through a 2-3 stage ring at ~2.5 us per ring cycle, while the launch has only 32 - 128 CTAs for 296 slots.  Splitting
// the K range of a tile over S CTAs shortens every chain S-fold and fills the idle SMs.
//
// How: blockIdx.z = split.  Every CTA accumulates its k-blocks in TMEM exactly as K2 does, then adds its fp32 partial
// tile into the tile's workspace slot with 16-byte vector reductions (L2 atomics), fences, and takes a ticket.  The CTA
// that draws the last ticket re-reads the summed tile (L2), runs K2's epilogue on it (bias, activation, BatchNorm
// statistics, affine, residual, fp16 store) and leaves the slot and the ticket counter zeroed for the next launch.  No
// CTA ever waits for another one, so no co-residency is required.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace {

using namespace mg::ptx;

constexpr int BM = 128;
constexpr int THREADS = 192;       // warp 0: TMA, warp 1: MMA + TMEM, warps 2..5: epilogue
constexpr int MAX_TAPS = MG_CONV_MAX_TAPS;
constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;
constexpr int COUNTER_BYTES = 16 * 1024;   // ticket counters at the head of the workspace (4096 tiles)

struct SArgs {
    int n_taps;
    int tap_dy[MAX_TAPS], tap_dx[MAX_TAPS], tap_koff[MAX_TAPS];
    int sy, sx, Hg, Wg, th, tw, tiles_y, tiles_x;
    int BK, kchunks, BN, Co, stages, swizzle, splits;
    __half* out;
    int Ho, Wo, Cs, c_off, oys, oy0, oxs, ox0;
    int pre_act, post_act;
    float* stats;
    const float* bias;
    const float* scale;
    const float* shift;
    const __half* res;
    int res_up;
    unsigned int* counters;   // [tiles * n-tiles], zero between launches
    float* ws;                // [tiles * n-tiles][128][BN] fp32, zero between launches
};

__device__ __forceinline__ float apply_act(float v, int act) {
    return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v > 0.f ? v : 0.2f * v) : v);
}

__global__ void __launch_bounds__(THREADS, 2)
conv_splitk_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SArgs a) {
    mg::pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = BM * a.BK * 2, b_bytes = a.BN * a.BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = sA + a.stages * a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + a.stages * b_bytes);  // full[stages], empty[stages], tmem_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * a.stages + 1);
    uint32_t* s_ticket = tmem_slot + 1;
    float* s_stage = reinterpret_cast<float*>(tmem_slot + 4);               // [4 warps][32][17]
    float* s_part = s_stage + 4 * 32 * 17;                                  // [4 warps][2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * a.stages, tfull = empty0 + 8 * a.stages;

    int t = blockIdx.x;
    const int tx = t % a.tiles_x;
    t /= a.tiles_x;
    const int ty = t % a.tiles_y, img = t / a.tiles_y;
    const int y0 = ty * a.th, x0 = tx * a.tw, n0 = blockIdx.y * a.BN;
    const int nkb_all = a.n_taps * a.kchunks;
    const int kb0 = (int)(((long long)blockIdx.z * nkb_all) / a.splits), kb1 = (int)(((long long)(blockIdx.z + 1) * nkb_all) / a.splits);
    const int nkb = kb1 - kb0;   // >= 1 (the host guarantees nkb_all >= 2 * splits)

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), a.BN < 32 ? 32 : a.BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mg::pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int kb = kb0 + i;
                const int s = i % a.stages, ph = (i / a.stages) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                mbar_expect_tx(full0 + 8 * s, a_bytes + b_bytes);
                const int tap = kb / a.kchunks, c = kb - tap * a.kchunks;
                tma_load_4d(smem_u32(sA + s * a_bytes), &tmA, full0 + 8 * s, c * a.BK, x0 * a.sx + a.tap_dx[tap],
                            y0 * a.sy + a.tap_dy[tap], img);
                tma_load_2d(smem_u32(sB + s * b_bytes), &tmB, full0 + 8 * s, a.tap_koff[tap] + c * a.BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc_f16(BM, a.BN, 0, 0);
            const uint32_t layout = swizzle_layout(a.swizzle), sbo = 8 * a.swizzle;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % a.stages, ph = (i / a.stages) & 1;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t abase = smem_u32(sA + s * a_bytes), bbase = smem_u32(sB + s * b_bytes);
                for (int k = 0; k < a.BK / 16; ++k) {
                    const uint64_t da = smem_desc(abase + k * 32, 0, sbo, layout);
                    const uint64_t db = smem_desc(bbase + k * 32, 0, sbo, layout);
                    mma_f16(tmem_base, da, db, idesc, (i | k) != 0);
                }
                mma_commit(empty0 + 8 * s);
            }
            mma_commit(tfull);
        }
    } else {
        const int q = warp & 3;
        const int m = q * 32 + lane;                  // accumulator row = pixel within the tile
        const int et = threadIdx.x - 64;              // 0..127
        const size_t tile_id = (size_t)blockIdx.x * gridDim.y + blockIdx.y;
        float* wrow = a.ws + (tile_id * BM + m) * a.BN;
        mbar_wait(tfull, 0);
        tc_fence_after();
        // ---- phase 1: this CTA's partial tile -> workspace (vector reductions at the L2)
        for (int c0 = 0; c0 < a.BN; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                red_add_v4(wrow + c0 + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                           __uint_as_float(r[i + 3]));
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) *s_ticket = atomicAdd(a.counters + tile_id, 1u);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*s_ticket == (unsigned)a.splits - 1u) {
            // ---- phase 2 (the CTA that completed the tile): K2's epilogue on the summed tile; the slot is left zeroed
            __threadfence();
            const int py = y0 + m / a.tw, px = x0 + m % a.tw;
            const bool valid = (py < a.Hg) && (px < a.Wg);
            const int oy = py * a.oys + a.oy0, ox = px * a.oxs + a.ox0;
            __half* orow = a.out + (((size_t)img * a.Ho + oy) * a.Wo + ox) * a.Cs + a.c_off + n0;
            const __half* rrow = nullptr;
            if (a.res)
                rrow = a.res + (a.res_up ? (((size_t)img * (a.Ho >> 1) + (oy >> 1)) * (a.Wo >> 1) + (ox >> 1))
                                         : (((size_t)img * a.Ho + oy) * a.Wo + ox)) * a.Co + n0;
            float* stg = s_stage + q * 32 * 17;
            float* part = s_part + q * 2 * a.BN;
            for (int c0 = 0; c0 < a.BN; c0 += 16) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 z = __ldcg(reinterpret_cast<const float4*>(wrow + c0 + i));
                    __stcg(reinterpret_cast<float4*>(wrow + c0 + i), make_float4(0.f, 0.f, 0.f, 0.f));
                    v[i] = z.x, v[i + 1] = z.y, v[i + 2] = z.z, v[i + 3] = z.w;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (a.bias) v[i] += __ldg(a.bias + n0 + c0 + i);
                    v[i] = apply_act(v[i], a.pre_act);
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
                        for (int rr = 0; rr < 32; ++rr) acc += stg[rr * 17 + col];
                    } else {
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr) { const float z = stg[rr * 17 + col]; acc += z * z; }
                    }
                    part[(lane >> 4) * a.BN + c0 + col] = acc;
                }
                if (valid && n0 + c0 < a.Co) {
                    if (a.scale) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], __ldg(a.scale + n0 + c0 + i), __ldg(a.shift + n0 + c0 + i));
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
                        for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], a.post_act);
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
            if (et == 0) a.counters[tile_id] = 0u;
            if (a.stats) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                float* dst = a.stats + (size_t)(blockIdx.x % STAT_COPIES) * 2 * a.Co;
                for (int i = et; i < 2 * a.BN; i += 128) {
                    const int kind = i / a.BN, c = i - kind * a.BN;
                    if (n0 + c < a.Co) {
                        const float tot = s_part[0 * 2 * a.BN + i] + s_part[1 * 2 * a.BN + i] + s_part[2 * 2 * a.BN + i] +
                                          s_part[3 * 2 * a.BN + i];
                        atomicAdd(dst + kind * a.Co + n0 + c, tot);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.BN < 32 ? 32 : a.BN);
    }
}

}  // namespace

namespace mg {

// MG_OK + *handled = true when the layer ran on the split-K kernel; *handled = false: not eligible (the caller continues
// with K2).  Only called when the descriptor carries a workspace.
int conv_splitk_launch(const mg_conv_desc* d, void* stream, bool* handled) {
    *handled = false;
    if (!d->splitk_ws || d->n_phases > 1) return MG_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) return MG_OK;
    SArgs a;
    a.n_taps = d->n_taps;
    for (int t = 0; t < d->n_taps; ++t) a.tap_dy[t] = d->tap_dy[t], a.tap_dx[t] = d->tap_dx[t], a.tap_koff[t] = d->tap_koff[t];
    a.sy = d->sy, a.sx = d->sx, a.Hg = d->Hg, a.Wg = d->Wg;
    if (d->Wg > 8) a.th = 8, a.tw = 16; else a.th = 16, a.tw = 8;
    a.tiles_y = ceil_div(d->Hg, a.th), a.tiles_x = ceil_div(d->Wg, a.tw);
    if (d->Ci % 64) return MG_OK;
    a.BK = 64, a.kchunks = d->Ci / 64, a.swizzle = 128;
    a.Co = d->Co;
    if (d->Co % 128) return MG_OK;
    a.BN = 128;
    const int tiles = d->N * a.tiles_y * a.tiles_x, ntiles = d->Co / a.BN, ctas = tiles * ntiles;
    const int nkb = d->n_taps * a.kchunks;
    int splits = std::min(8, std::min((2 * kNumSMs) / ctas, nkb / 4));
    if (splits < 2 || tiles * ntiles > COUNTER_BYTES / 4) return MG_OK;
    if ((size_t)COUNTER_BYTES + (size_t)ctas * BM * a.BN * sizeof(float) > (size_t)d->splitk_ws_bytes) return MG_OK;
    a.splits = splits;
    a.out = static_cast<__half*>(d->out);
    a.Ho = d->Ho, a.Wo = d->Wo, a.Cs = d->Cs, a.c_off = d->c_off;
    a.oys = d->oys, a.oy0 = d->oy0, a.oxs = d->oxs, a.ox0 = d->ox0;
    a.pre_act = d->pre_act, a.post_act = d->post_act, a.stats = d->stats, a.bias = d->bias;
    a.scale = d->scale, a.shift = d->shift, a.res = static_cast<const __half*>(d->res), a.res_up = d->res_up;
    if (d->res && (d->c_off != 0 || d->Cs != d->Co)) return MG_OK;
    a.counters = reinterpret_cast<unsigned int*>(d->splitk_ws);
    a.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->splitk_ws) + COUNTER_BYTES);

    const int a_bytes = BM * a.BK * 2, b_bytes = a.BN * a.BK * 2;
    const int fixed = 1024 + 256 + 4 * 32 * 17 * 4 + 4 * 2 * a.BN * 4;
    a.stages = std::max(2, std::min(4, (100 * 1024 - fixed) / (a_bytes + b_bytes)));
    const size_t smem = (size_t)fixed + (size_t)a.stages * (a_bytes + b_bytes);

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->Ci * 2, (cuuint64_t)d->Wi * d->Ci * 2, (cuuint64_t)d->Hi * d->Wi * d->Ci * 2};
        cuuint32_t box[4] = {(cuuint32_t)a.BK, (cuuint32_t)(a.tw * a.sx), (cuuint32_t)(a.th * a.sy), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)a.sx, (cuuint32_t)a.sy, 1};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->x), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(a.swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return MG_OK;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Co};
        cuuint64_t strides[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)a.BK, (cuuint32_t)a.BN};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(a.swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return MG_OK;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_splitk_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            set_error("mg_conv_fprop: cannot raise dynamic shared memory limit (split-K kernel)");
            return MG_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid(tiles, ntiles, splits);
    MG_LAUNCH(conv_splitk_tcgen05_kernel, grid, THREADS, smem, stream, tmA, tmB, a);
    MG_CHECK_LAUNCH("mg_conv_fprop(split-K)");
    *handled = true;
    return MG_OK;
}

}  // namespace mg
