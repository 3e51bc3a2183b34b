// K4: convolution weight gradient on tcgen05 tensor cores.
//
//   dW[co][tap][ci] = sum over pixels p   dY[p][co] * X[p shifted by tap][ci]
//
// GEMM with K = pixels.  Both operands are NHWC, i.e. channel-contiguous = "MN-major" UMMA operands: a TMA box
// {64 channels, tw, th} of a 128-pixel tile lands in shared memory as 128 rows of 128 bytes (one row per pixel =
// one K index) with the 128-byte swizzle, which is exactly the canonical MN-major SWIZZLE_128B atom stack
// (SBO = 8 rows = 1024 B between K groups, LBO = 128 rows = 16 KB between 64-channel atoms).  The shifted X box
// uses the same hardware zero fill / element strides as the forward kernel.  M = 128 output channels (rows of a
// narrower layer are zero-filled by TMA), N = up to 256 input channels, accumulators [128 x N] fp32 in TMEM.
// Split-K over pixel tiles across CTAs; partial results are added to the fp32 dW pack with TMA reduce-add
// (cp.reduce.async.bulk.tensor) from a swizzled staging tile.
// (Measured and dropped: a halo-patch form - one X patch per dy group, the three dx taps read from shifted descriptor starts
//  and sharing the dY tile, 3x fewer staged bytes per FLOP - was correct but not faster on any mid-resolution layer, with 8 x 16
//  and with 4 x 16 pixel tiles: 64^2 128->128 16.2 vs 16.3 us, 32^2 256->256 24.2 vs 14.1 us.)
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace {

using namespace mg::ptx;

constexpr int THREADS = 192;
constexpr int MAX_TAPS = MG_CONV_MAX_TAPS;

struct WArgs {
    int tap_dy[MAX_TAPS], tap_dx[MAX_TAPS], tap_koff[MAX_TAPS];
    int n_taps, taps_per_cta;
    int sy, sx, ays, ay0, axs, ax0;
    int th, tw, tiles_y, tiles_x, n_tiles, tiles_per_cta;
    int Co, Ci, Ktot, BN, atomw_b, swz_b, n_atoms_b, co_tiles, stages, a_atoms;  // a_atoms: real 64-channel atoms of dY (1|2)
    float* dw;
};

// One CTA: a range of 128-pixel tiles (split-K) x a group of taps x one (Co tile, Ci tile).  Per pixel tile the dY
// tile is loaded ONCE and multiplied with the shifted X tile of every tap of the group (accumulators side by side in
// TMEM: taps_per_cta * BN columns).  A layer with <= 64 output channels has only one real 64-channel atom of dY; the
// second atom of the M = 128 operand is a shared zero region reached through the descriptor's leading-byte offset.
__global__ void __launch_bounds__(THREADS, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ CUtensorMap tmDW, const WArgs a) {
    mg::pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int atom_bytes = 128 * 128;
    const int a_bytes = a.a_atoms * atom_bytes;
    const int b_atom_bytes = 128 * a.swz_b, b_tile = a.n_atoms_b * b_atom_bytes, b_bytes = a.taps_per_cta * b_tile;
    uint8_t* sA = smem;
    uint8_t* sB = sA + a.stages * a_bytes;
    uint8_t* sZ = sB + a.stages * b_bytes;                      // 16 KB of zeros (only read when a_atoms == 1)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sZ + atom_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * a.stages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * a.stages, tfull = empty0 + 8 * a.stages;
    const int t0 = blockIdx.y * a.taps_per_cta, nt = min(a.taps_per_cta, a.n_taps - t0);
    const int co0 = (blockIdx.z % a.co_tiles) * 128, ci0 = (blockIdx.z / a.co_tiles) * a.BN;
    const int t_begin = blockIdx.x * a.tiles_per_cta, t_end = min(t_begin + a.tiles_per_cta, a.n_tiles);
    const int nk = t_end - t_begin;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(nt * a.BN)) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDY);
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmDW);
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    if (a.a_atoms == 1) {
        for (int i = threadIdx.x; i < atom_bytes / 16; i += THREADS) reinterpret_cast<uint4*>(sZ)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mg::pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    if (nk > 0) {
        if (warp == 0) {
            if (lane == 0) {
                for (int i = 0; i < nk; ++i) {
                    const int s = i % a.stages, ph = (i / a.stages) & 1;
                    int t = t_begin + i;
                    const int tx = t % a.tiles_x;
                    t /= a.tiles_x;
                    const int ty = t % a.tiles_y, img = t / a.tiles_y;
                    const int y0 = ty * a.th, x0 = tx * a.tw;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    mbar_expect_tx(full0 + 8 * s, a_bytes + nt * b_tile);
                    const uint32_t dstA = smem_u32(sA + s * a_bytes), dstB = smem_u32(sB + s * b_bytes);
                    for (int j = 0; j < a.a_atoms; ++j)
                        tma_load_4d(dstA + j * atom_bytes, &tmDY, full0 + 8 * s, co0 + 64 * j, x0 * a.axs + a.ax0,
                                    y0 * a.ays + a.ay0, img);
                    for (int tt = 0; tt < nt; ++tt)
                        for (int j = 0; j < a.n_atoms_b; ++j)
                            tma_load_4d(dstB + tt * b_tile + j * b_atom_bytes, &tmX, full0 + 8 * s, ci0 + a.atomw_b * j,
                                        x0 * a.sx + a.tap_dx[t0 + tt], y0 * a.sy + a.tap_dy[t0 + tt], img);
                }
            }
        } else if (warp == 1) {
            // whole warp, warp-uniform operands, one elected lane issues (no elect / R2UR waterfall per tcgen05.mma)
            const uint32_t tmem_u = uniform_u32(tmem_base);
            const uint32_t idesc = instr_desc_f16(128, a.BN, 1, 1);  // both operands MN-major
            const uint32_t layB = swizzle_layout(a.swz_b);
            const uint32_t zbase = smem_u32(sZ), sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            int s = 0, ph = 0;
            for (int i = 0; i < nk; ++i) {
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t abase = sA_u + s * a_bytes;
                const uint32_t lbo_a = a.a_atoms == 2 ? (uint32_t)atom_bytes : zbase - abase;
                if (elect_one()) {
                    // tap outer, k inner (measured: switching accumulator / B tile on every MMA is ~25 % slower)
                    for (int tt = 0; tt < nt; ++tt) {
                        const uint32_t bbase = sB_u + s * b_bytes + tt * b_tile;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {  // 128 pixels = 8 x UMMA_K(16)
                            const uint64_t da = smem_desc(abase + k * 16 * 128, lbo_a, 8 * 128, 2);
                            const uint64_t db = smem_desc(bbase + k * 16 * a.swz_b, b_atom_bytes, 8 * a.swz_b, layB);
                            mma_f16(tmem_u + tt * a.BN, da, db, idesc, (i | k) != 0);
                        }
                    }
                    mma_commit(empty0 + 8 * s);
                }
                if (++s == a.stages) s = 0, ph ^= 1;
            }
            if (elect_one()) mma_commit(tfull);
        } else {
            const int q = warp & 3;
            const int co = co0 + q * 32 + lane;
            mbar_wait(tfull, 0);
            tc_fence_after();
            const int et = threadIdx.x - 64, arow = q * 32 + lane;            // epilogue thread index / accumulator lane = dW row
            if (a.BN < 32) {
                // 16-channel layers: plain vector reductions (a 32-column chunk does not exist)
                for (int tt = 0; tt < nt; ++tt) {
                    float* drow = a.dw + (size_t)co * a.Ktot + a.tap_koff[t0 + tt] + ci0;
                    uint32_t r[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tt * a.BN, r);
                    tmem_ld_wait();
                    if (co < a.Co) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            red_add_v4(drow + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                       __uint_as_float(r[i + 3]));
                    }
                }
            } else {
                // Split-K flush through TMA reduce-add.  (Per-thread red.global.add.v4 made every warp instruction 32 separate
                // 16-byte L2 atomics - lane = dW row, rows Ktot floats apart: 1.8 M atomic transactions per 64^2 128->128
                // launch.)  The [128 x 32] fp32 chunks of the accumulator go through a ring of four 128-byte-swizzled staging
                // tiles (the stage buffers are free once the accumulators are complete) and leave as
                // cp.reduce.async.bulk.tensor: one 128-byte line per dW row and chunk; rows beyond Co are clipped by the map.
                // The CTAs of a (tap group, channel tile) start at different chunks and wrap around.
                const int chunks_per_tap = a.BN >> 5, total = nt * chunks_per_tap;
                const int rot = (int)((blockIdx.x * 5u) % (unsigned)total);
                uint8_t* stg = sA;                                              // 4 x 16 KB, 1 KB aligned
                for (int it = 0; it < total; ++it) {
                    int g = it + rot;
                    if (g >= total) g -= total;
                    const int tt = g / chunks_per_tap, c0 = (g - tt * chunks_per_tap) << 5;
                    const int buf = it & 3;
                    if (it >= 4) {
                        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");   // this buffer's last store was read
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                    }
                    uint32_t r0[16], r1[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tt * a.BN + c0, r0);
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tt * a.BN + c0 + 16, r1);
                    tmem_ld_wait();
                    const uint32_t row = smem_u32(stg + buf * 16384) + arow * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((j ^ (arow & 7)) << 4)), "r"(r0[4 * j]),
                                     "r"(r0[4 * j + 1]), "r"(r0[4 * j + 2]), "r"(r0[4 * j + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((j + 4) ^ (arow & 7)) << 4)), "r"(r1[4 * j]),
                                     "r"(r1[4 * j + 1]), "r"(r1[4 * j + 2]), "r"(r1[4 * j + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et == 0) {
                        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(reinterpret_cast<uint64_t>(&tmDW)), "r"(smem_u32(stg + buf * 16384)),
                                       "r"(a.tap_koff[t0 + tt] + ci0 + c0), "r"(co0) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace

namespace mg {
int wgrad_halo_launch(const mg_wgrad_desc* d, void* stream, bool* handled);   // k4b_wgrad_halo.cu
}

extern "C" int mg_conv_wgrad(const mg_wgrad_desc* d, void* stream) {
    MG_REQUIRE(d && d->dy && d->x && d->dw, "mg_conv_wgrad: null pointer");
    MG_REQUIRE(d->n_taps >= 1 && d->n_taps <= MAX_TAPS, "mg_conv_wgrad: n_taps %d out of range", d->n_taps);
    MG_REQUIRE(d->Ci % 16 == 0 && d->Co % 8 == 0, "mg_conv_wgrad: Ci %% 16 and Co %% 8 required (Ci=%d Co=%d)", d->Ci, d->Co);
    MG_REQUIRE(d->sy >= 1 && d->sy <= 2 && d->ays >= 1 && d->ays <= 2, "mg_conv_wgrad: strides must be 1 or 2");
    if (!mg::get_encode()) {
        mg::set_error("mg_conv_wgrad: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return MG_ERR_CUDA;
    }
    {
        // high-resolution 32-channel 3x3 layers: persistent halo-resident kernel (K4b)
        bool handled = false;
        const int rc = mg::wgrad_halo_launch(d, stream, &handled);
        if (rc != MG_OK || handled) return rc;
    }
    WArgs a;
    for (int t = 0; t < d->n_taps; ++t) a.tap_dy[t] = d->tap_dy[t], a.tap_dx[t] = d->tap_dx[t], a.tap_koff[t] = d->tap_koff[t];
    a.sy = d->sy, a.sx = d->sx, a.ays = d->ays, a.ay0 = d->ay0, a.axs = d->axs, a.ax0 = d->ax0;
    if (d->Wg > 8) a.th = 8, a.tw = 16; else a.th = 16, a.tw = 8;
    a.tiles_y = mg::ceil_div(d->Hg, a.th), a.tiles_x = mg::ceil_div(d->Wg, a.tw);
    a.n_tiles = d->N * a.tiles_y * a.tiles_x;
    a.Co = d->Co, a.Ci = d->Ci, a.Ktot = d->Ktot, a.dw = d->dw;
    a.BN = d->Ci >= 256 ? 256 : d->Ci;       // Ci in {16,32,64,128,256,512,1280}
    MG_REQUIRE(d->Ci % a.BN == 0 && (a.BN % 64 == 0 || a.BN == 32 || a.BN == 16), "mg_conv_wgrad: unsupported Ci=%d", d->Ci);
    a.atomw_b = a.BN >= 64 ? 64 : a.BN;
    a.swz_b = a.atomw_b * 2;
    a.n_atoms_b = a.BN / a.atomw_b;
    a.co_tiles = mg::ceil_div(d->Co, 128);
    const int ci_tiles = d->Ci / a.BN;
    a.n_taps = d->n_taps;
    a.a_atoms = d->Co > 64 ? 2 : 1;
    const int a_bytes = a.a_atoms * 128 * 128, b_tile = a.n_atoms_b * 128 * a.swz_b;
    // taps per CTA: bounded by TMEM (512 columns) and by two pipeline stages in ~200 KB of shared memory
    a.taps_per_cta = std::max(1, std::min(std::min(d->n_taps, 512 / a.BN), (int)((100 * 1024 - a_bytes) / b_tile)));
    while (d->n_taps % a.taps_per_cta) --a.taps_per_cta;   // equal-sized tap groups (9 -> 9 | 3 | 1, 4 -> 4 | 2 | 1)
    if (a.BN == 64) a.taps_per_cta = 1;                    // measured: 64-channel layers prefer 2 small CTAs per SM
    a.stages = a.taps_per_cta == 1 && a.BN <= 128 ? std::max(2, std::min(3, (int)((94 * 1024) / (a_bytes + b_tile))))
                                                  : std::max(2, std::min(4, (int)((190 * 1024) / (a_bytes + a.taps_per_cta * b_tile))));
    while ((size_t)a.stages * (a_bytes + a.taps_per_cta * b_tile) < 64 * 1024) ++a.stages;   // the epilogue's staging ring lives there
    const size_t smem = 1024 + 256 + 128 * 128 + (size_t)a.stages * (a_bytes + a.taps_per_cta * b_tile);
    const int tap_groups = mg::ceil_div(d->n_taps, a.taps_per_cta);
    const int base_ctas = tap_groups * a.co_tiles * ci_tiles;
    const int target_ctas = smem > 113 * 1024 ? mg::kNumSMs : 2 * mg::kNumSMs;  // one wave of resident CTAs
    int splits = std::max(1, std::min(a.n_tiles, mg::ceil_div(target_ctas, base_ctas)));
    a.tiles_per_cta = mg::ceil_div(a.n_tiles, splits);
    splits = mg::ceil_div(a.n_tiles, a.tiles_per_cta);

    CUtensorMap tmDY, tmX;
    CUresult r = mg::encode_nhwc(&tmDY, d->dy, d->N, d->Hy, d->Wy, d->Co, 64, a.tw * a.axs, a.th * a.ays, a.axs, a.ays, 128);
    if (r != CUDA_SUCCESS) {
        mg::set_error("mg_conv_wgrad: cuTensorMapEncodeTiled(dY) failed (%d)", (int)r);
        return MG_ERR_CUDA;
    }
    r = mg::encode_nhwc(&tmX, d->x, d->N, d->Hi, d->Wi, d->Ci, a.atomw_b, a.tw * a.sx, a.th * a.sy, a.sx, a.sy, a.swz_b);
    if (r != CUDA_SUCCESS) {
        mg::set_error("mg_conv_wgrad: cuTensorMapEncodeTiled(X) failed (%d)", (int)r);
        return MG_ERR_CUDA;
    }
    CUtensorMap tmDW;
    {
        // dW pack [Co][Ktot] fp32: [128 rows x 32 floats] boxes, 128-byte swizzle (the reduce-add flush of the epilogue)
        cuuint64_t dims[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Co};
        cuuint64_t strides[1] = {(cuuint64_t)d->Ktot * 4};
        cuuint32_t box[2] = {32, 128};
        cuuint32_t estr[2] = {1, 1};
        r = mg::get_encode()(&tmDW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d->dw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            mg::set_error("mg_conv_wgrad: cuTensorMapEncodeTiled(dW) failed (%d)", (int)r);
            return MG_ERR_CUDA;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            mg::set_error("mg_conv_wgrad: cannot raise dynamic shared memory limit");
            return MG_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid(splits, tap_groups, a.co_tiles * ci_tiles);
    MG_LAUNCH(wgrad_tcgen05_kernel, grid, THREADS, smem, stream, tmDY, tmX, tmDW, a);
    MG_CHECK_LAUNCH("mg_conv_wgrad");
    return MG_OK;
}
