// K7: alpha heads - bilinear upsampling (align_corners = False, integer scale 1 / 2 / 4 / 8) + (tanh + 1) / 2 + optional
// per-plane factor (the `valid` mask of the OS8 head), forward and backward, fp32.
//   forward : one thread per output pixel (coalesced stores; the <= 4 coarse taps come from L1 / L2)
//   backward: gather form - a block owns a tile of COARSE pixels, evaluates d alpha / d (interpolated logit) * upstream
//             gradient once per fine pixel of the tile's footprint into shared memory, and every coarse pixel then sums
//             its (2 S)^2 window with the separable bilinear weights: no atomics, no scatter.
// Same arithmetic as torch: src = (dst + 0.5) / S - 0.5 clamped at 0, x1 = x0 + (x0 < n - 1), weights (1 - l, l).
#include "common.cuh"

namespace {

__device__ __forceinline__ void src_index(int d, float rs, int n, int& i0, int& i1, float& l1) {
    float s = rs * ((float)d + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

__device__ __forceinline__ float bilerp(const float* __restrict__ p, int w, int y0, int y1, float ly, int x0, int x1, float lx) {
    const float h0 = 1.f - ly, w0 = 1.f - lx;
    return h0 * (w0 * __ldg(p + (size_t)y0 * w + x0) + lx * __ldg(p + (size_t)y0 * w + x1)) +
           ly * (w0 * __ldg(p + (size_t)y1 * w + x0) + lx * __ldg(p + (size_t)y1 * w + x1));
}

__global__ void __launch_bounds__(256)
upsample_tanh_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ plane_scale, float* __restrict__ out,
                         int planes, int h, int w, int S, int32_t* __restrict__ all_zero) {
    mg::pdl_prologue();
    bool nz = false;
    const int H = h * S, W = w * S;
    const float rs = 1.f / (float)S;
    const size_t total = (size_t)planes * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), pl = (int)(i / ((size_t)W * H));
        const float* p = logits + (size_t)pl * h * w;
        float v;
        if (S == 1) {
            v = __ldg(p + (size_t)y * w + x);
        } else {
            int y0, y1, x0, x1;
            float ly, lx;
            src_index(y, rs, h, y0, y1, ly);
            src_index(x, rs, w, x0, x1, lx);
            v = bilerp(p, w, y0, y1, ly, x0, x1, lx);
        }
        float a = (tanhf(v) + 1.0f) * 0.5f;
        if (plane_scale) a *= __ldg(plane_scale + pl);
        out[i] = a;
        nz |= a != 0.f;
    }
    // device-side form of the reference's `x_os8.sum() == 0` test (decoder/resnet_inst_matt_spconv.py:314): no host read
    if (all_zero && __any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) *all_zero = 0;
}

// Four consecutive output pixels per thread, 32-bit index arithmetic, one 16-byte store (W % 4 == 0 and < 2^31 outputs): the
// one-pixel form above spends most of its instructions on three 64-bit divisions per pixel.  Same arithmetic per pixel.
__global__ void __launch_bounds__(256)
upsample_tanh_fwd4_kernel(const float* __restrict__ logits, const float* __restrict__ plane_scale, float* __restrict__ out,
                          int planes, int h, int w, int S, int32_t* __restrict__ all_zero) {
    mg::pdl_prologue();
    bool nz = false;
    const unsigned H = h * S, W = w * S, W4 = W >> 2, total4 = (unsigned)planes * H * W4;
    const float rs = 1.f / (float)S;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const unsigned row = i / W4, pl = row / H;
        const int x = (int)(i - row * W4) << 2, y = (int)(row - pl * H);
        const float* p = logits + (size_t)pl * h * w;
        const float ps = plane_scale ? __ldg(plane_scale + pl) : 1.f;
        float a[4];
        if (S == 1) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)y * w + x));
            a[0] = v.x, a[1] = v.y, a[2] = v.z, a[3] = v.w;
        } else {
            int y0, y1;
            float ly;
            src_index(y, rs, h, y0, y1, ly);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int x0, x1;
                float lx;
                src_index(x + j, rs, w, x0, x1, lx);
                a[j] = bilerp(p, w, y0, y1, ly, x0, x1, lx);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a[j] = (tanhf(a[j]) + 1.0f) * 0.5f;
            if (plane_scale) a[j] *= ps;
            nz |= a[j] != 0.f;
        }
        *reinterpret_cast<float4*>(out + (size_t)row * W + x) = make_float4(a[0], a[1], a[2], a[3]);
    }
    if (all_zero && __any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) *all_zero = 0;
}

// grid (w / TC, h / TC, planes), block 256;  TC = coarse tile edge, footprint edge F = S * TC + S
template <int S, int TC>
__global__ void __launch_bounds__(256)
upsample_tanh_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ plane_scale, const float* __restrict__ g,
                         float* __restrict__ glogits, int h, int w) {
    mg::pdl_prologue();
    constexpr int F = S * TC + S;
    __shared__ float s_v[F][F + 1];
    const int H = h * S, W = w * S, pl = blockIdx.z;
    const int cy0 = blockIdx.y * TC, cx0 = blockIdx.x * TC;
    const int fy0 = S * cy0 - S / 2, fx0 = S * cx0 - S / 2;     // first fine pixel that can touch the tile
    const float* p = logits + (size_t)pl * h * w;
    const float* gp = g + (size_t)pl * H * W;
    const float ps = plane_scale ? __ldg(plane_scale + pl) : 1.f;
    const float rs = 1.f / (float)S;
    for (int i = threadIdx.x; i < F * F; i += 256) {
        const int ly_ = i / F, lx_ = i - ly_ * F;
        const int y = fy0 + ly_, x = fx0 + lx_;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            int y0, y1, x0, x1;
            float ly, lx;
            src_index(y, rs, h, y0, y1, ly);
            src_index(x, rs, w, x0, x1, lx);
            const float t = tanhf(bilerp(p, w, y0, y1, ly, x0, x1, lx));
            v = __ldg(gp + (size_t)y * W + x) * ps * 0.5f * (1.f - t * t);
        }
        s_v[ly_][lx_] = v;
    }
    // separable bilinear weights of the tile's coarse pixels over their 2S-wide windows, once per block (the same
    // src_index arithmetic; 0 outside the image): the window sums below are then two table reads and one FMA per tap
    __shared__ float s_wy[TC][2 * S], s_wx[TC][2 * S];
    for (int i = threadIdx.x; i < 2 * TC * 2 * S; i += 256) {
        const int which = i / (TC * 2 * S), j = i - which * (TC * 2 * S), c = j / (2 * S), r = j - c * (2 * S);
        const int cc = (which ? cx0 : cy0) + c, n = which ? w : h, N = which ? W : H;
        const int f = S * cc - S / 2 + r;
        float wt = 0.f;
        if (f >= 0 && f < N) {
            int i0, i1;
            float l;
            src_index(f, rs, n, i0, i1, l);
            wt = (i0 == cc ? 1.f - l : 0.f) + (i1 == cc ? l : 0.f);
        }
        (which ? s_wx : s_wy)[c][r] = wt;
    }
    __syncthreads();
    // 256 threads = TC*TC coarse pixels x PARTS row groups of the 2S-row window
    constexpr int PARTS = 256 / (TC * TC);
    const int c = threadIdx.x / PARTS, part = threadIdx.x - c * PARTS;
    const int cyl = c / TC, cxl = c % TC;
    const int cy = cy0 + cyl, cx = cx0 + cxl;
    float acc = 0.f;
    if (cy < h && cx < w) {
        // fine rows that can touch coarse row cy: [S*cy - S/2, S*cy + 3S/2), split over the PARTS threads of this pixel
        constexpr int ROWS = 2 * S / PARTS > 0 ? 2 * S / PARTS : 1;
        for (int ry = part * ROWS; ry < (part + 1) * ROWS && ry < 2 * S; ++ry) {
            const float wy = s_wy[cyl][ry];
            if (wy == 0.f) continue;
            const float* row = &s_v[S * cyl + ry][S * cxl];
            float racc = 0.f;
#pragma unroll
            for (int rx = 0; rx < 2 * S; ++rx) racc += s_wx[cxl][rx] * row[rx];
            acc += wy * racc;
        }
    }
#pragma unroll
    for (int d = PARTS / 2; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (part == 0 && cy < h && cx < w) glogits[((size_t)pl * h + cy) * w + cx] = acc;
}

__global__ void __launch_bounds__(256)
tanh_head_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ plane_scale, const float* __restrict__ g,
                     float* __restrict__ glogits, int planes, int hw) {
    mg::pdl_prologue();
    const size_t total = (size_t)planes * hw;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float t = tanhf(__ldg(logits + i));
        const float ps = plane_scale ? __ldg(plane_scale + i / hw) : 1.f;
        glogits[i] = __ldg(g + i) * ps * 0.5f * (1.f - t * t);
    }
}

}  // namespace

extern "C" int mg_upsample_tanh_fwd(const float* logits, const float* plane_scale, float* out, int planes, int h, int w, int S,
                                    int32_t* all_zero, void* stream) {
    MG_REQUIRE(S == 1 || S == 2 || S == 4 || S == 8, "mg_upsample_tanh_fwd: scale must be 1, 2, 4 or 8 (got %d)", S);
    if (planes <= 0 || h <= 0 || w <= 0) return MG_OK;
    MG_REQUIRE(logits && out, "mg_upsample_tanh_fwd: null pointer");
    const size_t total = (size_t)planes * h * S * w * S;
    if ((w * S) % 4 == 0 && total < ((size_t)1 << 31) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
        (S > 1 || (reinterpret_cast<uintptr_t>(logits) & 15) == 0)) {
        const int grid = (int)std::min<size_t>((total / 4 + 255) / 256, (size_t)mg::kNumSMs * 16);
        MG_LAUNCH(upsample_tanh_fwd4_kernel, grid, 256, 0, stream, logits, plane_scale, out, planes, h, w, S, all_zero);
        MG_CHECK_LAUNCH("mg_upsample_tanh_fwd");
        return MG_OK;
    }
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)mg::kNumSMs * 16);
    MG_LAUNCH(upsample_tanh_fwd_kernel, grid, 256, 0, stream, logits, plane_scale, out, planes, h, w, S, all_zero);
    MG_CHECK_LAUNCH("mg_upsample_tanh_fwd");
    return MG_OK;
}

extern "C" int mg_upsample_tanh_bwd(const float* logits, const float* plane_scale, const float* g, float* glogits, int planes,
                                    int h, int w, int S, void* stream) {
    MG_REQUIRE(S == 1 || S == 2 || S == 4 || S == 8, "mg_upsample_tanh_bwd: scale must be 1, 2, 4 or 8 (got %d)", S);
    if (planes <= 0 || h <= 0 || w <= 0) return MG_OK;
    MG_REQUIRE(logits && g && glogits, "mg_upsample_tanh_bwd: null pointer");
    MG_REQUIRE(planes <= 65535, "mg_upsample_tanh_bwd: too many planes");
    if (S == 1) {
        const size_t total = (size_t)planes * h * w;
        const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)mg::kNumSMs * 16);
        MG_LAUNCH(tanh_head_bwd_kernel, grid, 256, 0, stream, logits, plane_scale, g, glogits, planes, h * w);
    } else if (S == 2) {
        MG_LAUNCH((upsample_tanh_bwd_kernel<2, 16>), dim3(mg::ceil_div(w, 16), mg::ceil_div(h, 16), planes), 256, 0, stream, logits,
                  plane_scale, g, glogits, h, w);
    } else if (S == 4) {
        MG_LAUNCH((upsample_tanh_bwd_kernel<4, 8>), dim3(mg::ceil_div(w, 8), mg::ceil_div(h, 8), planes), 256, 0, stream, logits,
                  plane_scale, g, glogits, h, w);
    } else {
        MG_LAUNCH((upsample_tanh_bwd_kernel<8, 8>), dim3(mg::ceil_div(w, 8), mg::ceil_div(h, 8), planes), 256, 0, stream, logits,
                  plane_scale, g, glogits, h, w);
    }
    MG_CHECK_LAUNCH("mg_upsample_tanh_bwd");
    return MG_OK;
}
