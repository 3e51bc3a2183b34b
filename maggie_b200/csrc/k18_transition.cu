// K18: the transition ("trimap") ground truth of the loaders on the GPU.
//
//   gen_transition_gt (dataloader/utils.py:15-35):  kernel = cv2 MORPH_ELLIPSE (k, k);  for every instance alpha (uint8):
//       trans = (cv2.dilate(alpha, kernel, iterations) - cv2.erode(alpha, kernel, iterations)) > 0
//       trans |= (alpha > 127) != (mask == 255)          (mask at full size or at 1/8, repeated 8 x 8)
//   The training loader calls it with k in 2..4 and 5..14 iterations, the evaluation loader with k = 25 and one iteration.
//
// Grey-level morphology on uint8 planes: one pass = max (dilate) and min (erode) over the ellipse footprint anchored at
// (k/2, k/2), pixels outside the image ignored (cv2's default border value for morphology).  A 32 x 32 tile with its halo
// sits in shared memory for both chains; `iterations` passes ping-pong between two pairs of planes, the last pass writes the
// {0,1} map.  Loader-side work (per batch, not per layer): a brute-force window loop is enough (k = 25: 377 taps).
#include "common.cuh"

namespace {

constexpr int TILE = 32, MAXK = 31, HALO = MAXK / 2 + 1, SW = TILE + 2 * HALO;

// Row span of cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)), row i: columns [j1, j2)  (as in k8_unknown.cu).
__device__ __forceinline__ void ellipse_span(int k, int i, int& j1, int& j2) {
    const int r = k / 2, c = k / 2;
    const double inv_r2 = r ? __ddiv_rn(1.0, (double)(r * r)) : 0.0;
    const int dy = i - r;
    const int dx = (int)rint(__dmul_rn((double)c, __dsqrt_rn(__dmul_rn((double)(r * r - dy * dy), inv_r2))));
    j1 = max(c - dx, 0);
    j2 = min(c + dx + 1, k);
}

// one morphology pass of both chains: dil_out = max over the footprint of dil_in, ero_out = min of ero_in.
// last != 0: instead write trans = (max - min > 0) | ((alpha > 127) != (mask == 255)) as uint8 {0,1}.
__global__ void __launch_bounds__(256)
morph_pass_kernel(const uint8_t* __restrict__ dil_in, const uint8_t* __restrict__ ero_in, uint8_t* __restrict__ dil_out,
                  uint8_t* __restrict__ ero_out, const uint8_t* __restrict__ alpha, const uint8_t* __restrict__ mask, int mask_div,
                  uint8_t* __restrict__ trans, int H, int W, int k, int last) {
    mg::pdl_prologue();
    __shared__ uint8_t s_d[SW][SW + 4], s_e[SW][SW + 4];
    __shared__ int s_j1[MAXK], s_j2[MAXK];
    const int a = k / 2, x0 = blockIdx.x * TILE, y0 = blockIdx.y * TILE;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid < k) ellipse_span(k, tid, s_j1[tid], s_j2[tid]);
    const int span = TILE + k;                       // rows / cols y0 - a .. y0 + TILE + (k - 1 - a)
    for (int i = tid; i < span * span; i += 256) {
        const int r = i / span, c = i - r * span;
        const int y = y0 - a + r, x = x0 - a + c;
        const bool in = y >= 0 && y < H && x >= 0 && x < W;
        s_d[r][c] = in ? dil_in[plane + (size_t)y * W + x] : 0;       // neutral elements: outside pixels never win
        s_e[r][c] = in ? ero_in[plane + (size_t)y * W + x] : 255;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ty = threadIdx.y + 8 * q, tx = threadIdx.x;
        const int y = y0 + ty, x = x0 + tx;
        if (y >= H || x >= W) continue;
        int mx = 0, mn = 255;
        for (int i = 0; i < k; ++i) {
            const int j1 = s_j1[i], j2 = s_j2[i];
            for (int j = j1; j < j2; ++j) {
                mx = max(mx, (int)s_d[ty + i][tx + j]);
                mn = min(mn, (int)s_e[ty + i][tx + j]);
            }
        }
        const size_t o = plane + (size_t)y * W + x;
        if (!last) {
            dil_out[o] = (uint8_t)mx, ero_out[o] = (uint8_t)mn;
        } else {
            int t = mx - mn > 0;
            if (mask) {
                const int hm = H / mask_div, wm = W / mask_div;
                const uint8_t m = mask[(size_t)blockIdx.z * hm * wm + (size_t)(y / mask_div) * wm + x / mask_div];
                t |= (alpha[o] > 127) != (m == 255);
            }
            trans[o] = (uint8_t)t;
        }
    }
}

}  // namespace

extern "C" int mg_transition_gt(const void* alpha_u8, const void* mask_u8, int mask_div, int planes, int H, int W, int k_size,
                                int iterations, void* tmp_u8, void* trans_u8, void* stream) {
    if (planes <= 0 || H <= 0 || W <= 0) return MG_OK;
    MG_REQUIRE(alpha_u8 && trans_u8, "mg_transition_gt: null pointer");
    MG_REQUIRE(k_size >= 1 && k_size <= MAXK, "mg_transition_gt: k_size %d out of range (1..%d)", k_size, MAXK);
    MG_REQUIRE(iterations >= 1 && iterations <= 64, "mg_transition_gt: iterations %d out of range", iterations);
    MG_REQUIRE(iterations == 1 || tmp_u8, "mg_transition_gt: %d iterations need the 4-plane-set scratch buffer", iterations);
    MG_REQUIRE(!mask_u8 || mask_div == 1 || (mask_div == 8 && H % 8 == 0 && W % 8 == 0),
               "mg_transition_gt: mask_div must be 1 or 8 (with H, W multiples of 8)");
    MG_REQUIRE(planes <= 65535, "mg_transition_gt: too many planes");
    const size_t n = (size_t)planes * H * W;
    uint8_t* t = static_cast<uint8_t*>(tmp_u8);
    uint8_t* buf[2][2] = {{t, t ? t + n : nullptr}, {t ? t + 2 * n : nullptr, t ? t + 3 * n : nullptr}};   // [ping/pong][dil/ero]
    const uint8_t* a = static_cast<const uint8_t*>(alpha_u8);
    const uint8_t *din = a, *ein = a;
    dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, planes), block(32, 8);
    for (int it = 0; it < iterations; ++it) {
        const int last = it == iterations - 1;
        uint8_t* dout = last ? nullptr : buf[it & 1][0];
        uint8_t* eout = last ? nullptr : buf[it & 1][1];
        MG_LAUNCH(morph_pass_kernel, grid, block, 0, stream, din, ein, dout, eout, a, static_cast<const uint8_t*>(mask_u8), mask_div,
                  static_cast<uint8_t*>(trans_u8), H, W, k_size, last);
        din = dout, ein = eout;
    }
    MG_CHECK_LAUNCH("mg_transition_gt");
    return MG_OK;
}
