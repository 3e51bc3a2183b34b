// K4b: weight gradient of the high-resolution 32- and 64-channel 3x3 stride-1 layers (512^2 / 256^2 x 32, 128^2 x 64:
// HBM-bound, 2 x 134 MB per layer at 512^2), persistent and halo-resident like K2b.  With 64 input channels the
// [128 x 9 * 64] accumulator would not fit the 512 TMEM columns: the CTAs then form two populations, each loading and
// accumulating 32 of the channels (the X patch is still read once, the dY tile twice).
//
//   dW[co][(dy,dx)][ci] = sum over pixels p   dY[p][co] * X[p + (dy,dx)][ci]
//
//   * K = pixels.  One K block = one row of a 128-pixel strip.  A = the dY row tile (MN-major rows of 128 B = 64
//     channels, TMA zero-fills the channels beyond Co; the second 64-channel atom of the M = 128 operand is a shared zero
//     region reached through the leading-dimension offset, as in K4).
//   * B = the X halo patch of the tile ((R+2) x (128+2) pixels, one swizzled 64-byte row per pixel, ONE TMA box, hardware
//     zero fill = padding).  The operand of tap (dy, dx) is that patch read from row (r+1+dy)*P + (1+dx): a shift along
//     K.  The three dx taps of one dy are even fused into ONE MMA with N = 96: N-atom j of the MN-major operand is the
//     same patch shifted by j pixels, i.e. the leading-dimension (atom) stride of the descriptor is one row = 64 bytes.
//     24 MMAs per pixel row instead of 72 (MAGGIE_B200_WGRAD_HALO_MODE=1 selects the one-tap-per-MMA form).
//   * The [128 x 288] fp32 accumulator lives in TMEM for the CTA's whole lifetime (split-K over the CTA's tiles) and is
//     flushed ONCE with 16-byte vector reductions: X and dY are each read exactly once from HBM.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

using namespace mg::ptx;

constexpr int THREADS = 192;   // warp 0: TMA, warp 1: MMA + TMEM, warps 2..5: final epilogue
constexpr int WS = 128, P = WS + 2;
constexpr int XROW = 64;       // bytes per patch pixel (32 channels)
constexpr int YROW = 128;      // bytes per dY row (64-channel atom)

std::atomic<unsigned long long> g_wgrad_halo_launches{0};

struct WHArgs {
    int H, W, Co, R, strips, rblocks, n_tiles, mode;
    int x_bytes, y_bytes;
    int halves, Ci;            // 64 input channels: two CTA populations, each with the [128 x 288] accumulator of 32 of them
    float* dw;                 // [Co][9*Ci]
};

__global__ void __launch_bounds__(THREADS, 1)
wgrad_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WHArgs a) {
    mg::pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // [1 KB guard][X0][X1][2 KB guard][Y0][Y1][zero atom 16 KB][barriers][tmem slot]
    const int x_al = (a.x_bytes + 1023) & ~1023, y_al = (a.y_bytes + 1023) & ~1023;
    uint8_t* sX = smem + 1024;
    uint8_t* sY = sX + 2 * x_al + 2048;
    uint8_t* sZ = sY + 2 * y_al;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sZ + 128 * YROW);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t full = bar0, empty = bar0 + 16, tfull = bar0 + 32;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDY);
        prefetch_tmap(&tmX);
        for (int i = 0; i < 2; ++i) mbar_init(full + 8 * i, 1), mbar_init(empty + 8 * i, 1);
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    for (int i = threadIdx.x; i < 128 * YROW / 16; i += THREADS) reinterpret_cast<uint4*>(sZ)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mg::pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_img = a.rblocks * a.strips;
    const int half = blockIdx.x % a.halves, tile0 = blockIdx.x / a.halves, tile_step = gridDim.x / a.halves;
    const bool any = tile0 < a.n_tiles;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = tile0; tile < a.n_tiles; tile += tile_step, ++it) {
                const int buf = it & 1, ph = (it >> 1) & 1;
                const int img = tile / tiles_img, rem = tile - img * tiles_img;
                const int rb = rem / a.strips, st = rem - rb * a.strips;
                mbar_wait(empty + 8 * buf, ph ^ 1);
                mbar_expect_tx(full + 8 * buf, a.x_bytes + a.y_bytes);
                tma_load_4d(smem_u32(sX + buf * x_al), &tmX, full + 8 * buf, half * 32, st * WS - 1, rb * a.R - 1, img);
                tma_load_4d(smem_u32(sY + buf * y_al), &tmDY, full + 8 * buf, 0, st * WS, rb * a.R, img);
            }
        }
    } else if (warp == 1) {
        // whole warp, warp-uniform operands, one elected lane issues (no elect / R2UR waterfall per tcgen05.mma)
        const uint32_t tmem_u = uniform_u32(tmem_base);
        const uint32_t idesc96 = instr_desc_f16(128, 96, 1, 1), idesc32 = instr_desc_f16(128, 32, 1, 1);
        const uint32_t zbase = smem_u32(sZ), sX_u = smem_u32(sX), sY_u = smem_u32(sY);
        int buf = 0, ph = 0;
        bool first = true;
        for (int tile = tile0; tile < a.n_tiles; tile += tile_step) {
            mbar_wait(full + 8 * buf, ph);
            tc_fence_after();
            const uint32_t xbase = sX_u + buf * x_al, ybase = sY_u + buf * y_al;
            if (elect_one()) {
                for (int r = 0; r < a.R; ++r) {
                    const uint32_t arow = ybase + r * 128 * YROW;
                    const uint32_t lbo_a = zbase - arow;                 // atom 1 of M = 128: the shared zero region
#pragma unroll
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        const uint32_t brow = xbase + ((r + dyi) * P) * XROW;   // patch row r + 1 + dy, dx = -1 is pixel 0
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint64_t da = smem_desc(arow + k * 16 * YROW, lbo_a, 8 * YROW, 2);
                            const uint32_t acc = (first && r == 0 && k == 0) ? 0u : 1u;
                            if (a.mode == 0) {
                                // N = 96: atoms 0..2 = the patch shifted by 0 / 1 / 2 pixels (dx = -1, 0, +1)
                                const uint64_t db = smem_desc(brow + k * 16 * XROW, XROW, 8 * XROW, 4);
                                mma_f16(tmem_u + dyi * 96, da, db, idesc96, acc);
                            } else {
#pragma unroll
                                for (int dxi = 0; dxi < 3; ++dxi) {
                                    const uint64_t db = smem_desc(brow + dxi * XROW + k * 16 * XROW, 128 * XROW, 8 * XROW, 4);
                                    mma_f16(tmem_u + dyi * 96 + dxi * 32, da, db, idesc32, acc);
                                }
                            }
                        }
                    }
                }
                mma_commit(empty + 8 * buf);
            }
            first = false;
            if (buf) ph ^= 1;
            buf ^= 1;
        }
        if (any && elect_one()) mma_commit(tfull);
    } else if (any) {
        // final epilogue: accumulator row = output channel (TMEM lane), 288 columns = (dy, dx, ci of this CTA's 32 channels)
        const int q = warp & 3, co = q * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
        if (q * 32 < a.Co) {
            for (int c0 = 0; c0 < 288; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
                tmem_ld_wait();
                if (co < a.Co) {
                    float* drow = a.dw + (size_t)co * (9 * a.Ci) + (c0 >> 5) * a.Ci + half * 32 + (c0 & 31);
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        red_add_v4(drow + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                   __uint_as_float(r[i + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

namespace mg {

int wgrad_halo_launch(const mg_wgrad_desc* d, void* stream, bool* handled) {
    *handled = false;
    const char* e = std::getenv("MAGGIE_B200_NO_HALO_CONV");
    if (e && e[0] == '1') return MG_OK;
    if (d->n_taps != 9 || (d->Ci != 32 && d->Ci != 64) || d->Co > 64 || d->Co % 8 || d->Ktot != 9 * d->Ci) return MG_OK;
    if (d->sy != 1 || d->sx != 1 || d->ays != 1 || d->axs != 1 || d->ay0 != 0 || d->ax0 != 0) return MG_OK;
    if (d->Hy != d->Hi || d->Wy != d->Wi || d->Hg != d->Hi || d->Wg != d->Wi || d->Wi % WS || d->Hi < 2) return MG_OK;
    for (int t = 0; t < 9; ++t)
        if (d->tap_dy[t] != t / 3 - 1 || d->tap_dx[t] != t % 3 - 1 || d->tap_koff[t] != t * d->Ci) return MG_OK;
    if (!get_encode()) return MG_OK;
    WHArgs a;
    a.H = d->Hi, a.W = d->Wi, a.Co = d->Co, a.dw = d->dw;
    a.R = 2;
    a.strips = d->Wi / WS, a.rblocks = ceil_div(d->Hi, a.R);
    a.n_tiles = d->N * a.rblocks * a.strips;
    a.halves = d->Ci / 32, a.Ci = d->Ci;
    if (a.n_tiles * a.halves < 2 * kNumSMs) return MG_OK;
    a.x_bytes = (a.R + 2) * P * XROW, a.y_bytes = a.R * 128 * YROW;
    {
        const char* m = std::getenv("MAGGIE_B200_WGRAD_HALO_MODE");
        a.mode = m ? std::atoi(m) : 0;
    }
    CUtensorMap tmDY, tmX;
    if (encode_nhwc(&tmDY, d->dy, d->N, d->Hy, d->Wy, d->Co, 64, WS, a.R, 1, 1, 128) != CUDA_SUCCESS) return MG_OK;
    if (encode_nhwc(&tmX, d->x, d->N, d->Hi, d->Wi, d->Ci, 32, P, a.R + 2, 1, 1, 64) != CUDA_SUCCESS) return MG_OK;
    const size_t smem = 1024 + 1024 + 2 * (size_t)((a.x_bytes + 1023) & ~1023) + 2048 + 2 * (size_t)((a.y_bytes + 1023) & ~1023) +
                        128 * YROW + 256;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(wgrad_halo_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) {
            set_error("mg_conv_wgrad: cannot raise dynamic shared memory limit (halo kernel)");
            return MG_ERR_CUDA;
        }
        attr_set = true;
    }
    MG_LAUNCH(wgrad_halo_tcgen05_kernel, std::min(a.n_tiles * a.halves, kNumSMs / a.halves * a.halves), THREADS, smem, stream, tmDY,
              tmX, a);
    MG_CHECK_LAUNCH("mg_conv_wgrad(halo)");
    g_wgrad_halo_launches.fetch_add(1, std::memory_order_relaxed);
    *handled = true;
    return MG_OK;
}

}  // namespace mg

extern "C" unsigned long long mg_wgrad_halo_launches(void) { return g_wgrad_halo_launches.load(); }
