// K16: the two ends of the path that touch loader / evaluator data; plain HBM streaming kernels.
//
//   mg_input_stage   : uint8 frames (HWC) + uint8 alphas + uint8 masks, as the dataset decodes them, -> the float tensors
//                      the model consumes: ToTensor + Normalize (dataloader/transforms.py:720-783: permute, /255,
//                      (x - mean) / std; alpha < 5 -> 0), alpha / 255, mask / 255 (dataloader/him.py:156-157) and the nearest
//                      1/8 mask down-sampling (him.py:175-176).  The host->device copy then moves 1 byte per value instead
//                      of 4, and no full-size float tensor is ever built on the CPU.
//   mg_alpha_finalize: evaluation tail - undo the padding / resizing of the test transforms
//                      (utils/postprocessing.py:36-64: crop, bilinear resize with align_corners = True) and clamp the
//                      near-0 / near-1 values (engine/test.py:141-142), one gather pass.
#include "common.cuh"

namespace {

struct Norm3 {
    float mean[3], std[3];
};

// one thread per pixel: 3 image bytes -> 3 planes; n_i alpha and mask bytes -> planes
__global__ void __launch_bounds__(256)
input_stage_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ alphas, const uint8_t* __restrict__ masks,
                   float* __restrict__ image, float* __restrict__ alpha, float* __restrict__ mask, Norm3 nm, int B, int n_i,
                   int H, int W, int mask_div) {
    mg::pdl_prologue();
    const size_t hw = (size_t)H * W, total = (size_t)B * hw;
    const int hm = H / mask_div, wm = W / mask_div;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / hw, p = i - b * hw;
        const uint8_t* f = frames + i * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) image[(b * 3 + c) * hw + p] = ((float)f[c] / 255.0f - nm.mean[c]) / nm.std[c];
        const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
        for (int k = 0; k < n_i; ++k) {
            const size_t q = (b * n_i + k) * hw + p;
            if (alphas) {
                const uint8_t a = alphas[q];
                alpha[q] = (a < 5 ? 0.0f : (float)a) / 255.0f;
            }
            if (masks) {
                const float m = (float)masks[q] / 255.0f;
                if (mask_div == 1) {
                    mask[q] = m;
                } else if (y % mask_div == 0 && x % mask_div == 0 && y / mask_div < hm && x / mask_div < wm) {
                    // F.interpolate(mode="nearest") to (H/8, W/8): source index = floor(dst * scale) = dst * 8
                    mask[((b * n_i + k) * hm + y / mask_div) * wm + x / mask_div] = m;
                }
            }
        }
    }
}

// out[pl][y][x] over the ORIGINAL size (Ho, Wo): bilinear (align_corners = True) sample of the (hc x wc) top-left crop of in
__global__ void __launch_bounds__(256)
alpha_finalize_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int h, int w, int hc, int wc, int Ho,
                      int Wo, int resize, float lo, float hi) {
    mg::pdl_prologue();
    const size_t total = (size_t)planes * Ho * Wo;
    const float sy = Ho > 1 ? (float)(hc - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(wc - 1) / (float)(Wo - 1) : 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho);
        const size_t pl = i / ((size_t)Wo * Ho);
        const float* src = in + pl * h * w;
        float v;
        if (!resize) {
            v = __ldg(src + (size_t)y * w + x);
        } else {
            const float fy = sy * (float)y, fx = sx * (float)x;
            const int y0 = (int)fy, x0 = (int)fx;
            const int y1 = y0 + (y0 < hc - 1 ? 1 : 0), x1 = x0 + (x0 < wc - 1 ? 1 : 0);
            const float ly = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), lx = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
            const float top = __fadd_rn(__fmul_rn(1.f - lx, __ldg(src + (size_t)y0 * w + x0)), __fmul_rn(lx, __ldg(src + (size_t)y0 * w + x1)));
            const float bot = __fadd_rn(__fmul_rn(1.f - lx, __ldg(src + (size_t)y1 * w + x0)), __fmul_rn(lx, __ldg(src + (size_t)y1 * w + x1)));
            v = __fadd_rn(__fmul_rn(1.f - ly, top), __fmul_rn(ly, bot));
        }
        if (v <= lo) v = 0.f;
        if (v >= hi) v = 1.f;
        out[i] = v;
    }
}

int stream_grid(size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)mg::kNumSMs * 16); }

}  // namespace

extern "C" int mg_input_stage(const void* frames_u8, const void* alphas_u8, const void* masks_u8, float* image, float* alpha,
                              float* mask, const float* mean3, const float* std3, int B, int n_i, int H, int W, int mask_div,
                              void* stream) {
    if (B <= 0 || H <= 0 || W <= 0) return MG_OK;
    MG_REQUIRE(frames_u8 && image && mean3 && std3, "mg_input_stage: null pointer");
    MG_REQUIRE((alphas_u8 == nullptr) == (alpha == nullptr) && (masks_u8 == nullptr) == (mask == nullptr),
               "mg_input_stage: alpha / mask inputs and outputs go together");
    MG_REQUIRE(n_i >= 0 && n_i <= 64, "mg_input_stage: bad instance count %d", n_i);
    MG_REQUIRE(mask_div == 1 || (mask_div == 8 && H % 8 == 0 && W % 8 == 0),
               "mg_input_stage: mask_div must be 1 or 8 (with H, W multiples of 8)");
    Norm3 nm;
    for (int c = 0; c < 3; ++c) {   // mean3 / std3 are HOST arrays (three floats each)
        nm.mean[c] = mean3[c], nm.std[c] = std3[c];
        MG_REQUIRE(nm.std[c] != 0.f, "mg_input_stage: std[%d] is zero", c);
    }
    MG_LAUNCH(input_stage_kernel, stream_grid((size_t)B * H * W), 256, 0, stream, static_cast<const uint8_t*>(frames_u8),
              static_cast<const uint8_t*>(alphas_u8), static_cast<const uint8_t*>(masks_u8), image, alpha, mask, nm, B, n_i, H, W,
              mask_div);
    MG_CHECK_LAUNCH("mg_input_stage");
    return MG_OK;
}

extern "C" int mg_alpha_finalize(const float* in, float* out, int planes, int h, int w, int pad_h, int pad_w, int out_h,
                                 int out_w, float lo, float hi, void* stream) {
    if (planes <= 0) return MG_OK;
    MG_REQUIRE(in && out, "mg_alpha_finalize: null pointer");
    const int hc = h - pad_h, wc = w - pad_w;
    MG_REQUIRE(pad_h >= 0 && pad_w >= 0 && hc > 0 && wc > 0, "mg_alpha_finalize: padding (%d, %d) does not fit %d x %d", pad_h,
               pad_w, h, w);
    const int resize = out_h > 0 && out_w > 0;
    const int Ho = resize ? out_h : hc, Wo = resize ? out_w : wc;
    MG_LAUNCH(alpha_finalize_kernel, stream_grid((size_t)planes * Ho * Wo), 256, 0, stream, in, out, planes, h, w, hc, wc, Ho, Wo,
              resize, lo, hi);
    MG_CHECK_LAUNCH("mg_alpha_finalize");
    return MG_OK;
}
