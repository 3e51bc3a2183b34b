// K11: the element-wise half of the video model (MaGGIe_Temp, BASELINE config C4) around the native convolutions.
//
//   ConvGRU step (module/conv_gru.py:50-70):   rz = sigmoid(conv_ih([x | h]));  c = tanh(conv_hh([x | r * h]));
//                                               h' = (1 - z) * h + z * c
//     gru_concat2      : cat1 = [x | h]                         (the operand of conv_ih; no torch.cat)
//     gru_gate1_fwd    : cat2 = [x | sigmoid(rz[:, :C]) * h]    (the operand of conv_hh, written in place of a second cat)
//     gru_gate2_fwd    : h'   = (1 - z) h + z tanh(c_pre),  z = sigmoid(rz[:, C:])
//     gru_gate2_bwd    : dh' -> d rz[:, C:], d c_pre, and the direct part of dh
//     gru_gate1_bwd    : d cat2 -> d rz[:, :C] and dpart = [d cat2[:, :C] | dh_direct + d cat2[:, C:] * r]; the data gradient
//                        of conv_ih then ADDS dpart in its epilogue (residual input) and leaves [dx | dh] in one tensor
//   Bidirectional temporal fusion (decoder/resnet_inst_matt_spconv_temp.py:122-142): the forward and the backward
//     recurrences  fp_i = fp_{i-1} (1 - s_i) + p_i s_i  over the frames of a clip and their average, for all instances of
//     a pixel in registers: one pass forward, one pass backward (recurrences recomputed, no saved intermediates).
// All tensors NHWC fp16 (GRU) / planar fp32 (fusion); fp32 arithmetic.  HBM-bound streaming kernels.
#include "common.cuh"

namespace {

struct H8 {
    uint4 u;
    __device__ __forceinline__ void to_float(float (&f)[8]) const {
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(h[i]);
            f[2 * i] = t.x, f[2 * i + 1] = t.y;
        }
    }
    __device__ __forceinline__ void from_float(const float (&f)[8]) {
        __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    }
};
__device__ __forceinline__ H8 ld8(const __half* p) {
    H8 v;
    v.u = __ldg(reinterpret_cast<const uint4*>(p));
    return v;
}
__device__ __forceinline__ void st8(__half* p, const H8& v) { *reinterpret_cast<uint4*>(p) = v.u; }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// one thread = 8 channels of one pixel; G = C / 8 groups per pixel
#define K11_FOR_EACH(P, G)                                                                             \
    const size_t total = (P) * (size_t)(G);                                                            \
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(256)
gru_concat2_kernel(const __half* __restrict__ x, const __half* __restrict__ h, __half* __restrict__ cat1, size_t P, int C) {
    mg::pdl_prologue();
    const int G = C >> 3;
    K11_FOR_EACH(P, 2 * G) {
        const size_t p = v / (2 * G);
        const int g = (int)(v - p * 2 * G);
        const __half* src = g < G ? x + p * C + g * 8 : h + p * C + (g - G) * 8;
        st8(cat1 + p * 2 * C + g * 8, ld8(src));
    }
}

__global__ void __launch_bounds__(256)
gru_gate1_fwd_kernel(const __half* __restrict__ rz, const __half* __restrict__ cat1, __half* __restrict__ cat2, size_t P, int C) {
    mg::pdl_prologue();
    const int G = C >> 3;
    K11_FOR_EACH(P, G) {
        const size_t p = v / G;
        const int c0 = (int)(v - p * G) << 3;
        const size_t row = p * 2 * C;
        st8(cat2 + row + c0, ld8(cat1 + row + c0));                         // x half: copied
        float r[8], h[8], o[8];
        ld8(rz + row + c0).to_float(r), ld8(cat1 + row + C + c0).to_float(h);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sigmoidf_(r[i]) * h[i];
        H8 w;
        w.from_float(o);
        st8(cat2 + row + C + c0, w);
    }
}

__global__ void __launch_bounds__(256)
gru_gate2_fwd_kernel(const __half* __restrict__ rz, const __half* __restrict__ cpre, const __half* __restrict__ cat1,
                     __half* __restrict__ hnew, size_t P, int C) {
    mg::pdl_prologue();
    const int G = C >> 3;
    K11_FOR_EACH(P, G) {
        const size_t p = v / G;
        const int c0 = (int)(v - p * G) << 3;
        float z[8], c[8], h[8], o[8];
        ld8(rz + p * 2 * C + C + c0).to_float(z), ld8(cpre + p * C + c0).to_float(c), ld8(cat1 + p * 2 * C + C + c0).to_float(h);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float zz = sigmoidf_(z[i]);
            o[i] = (1.f - zz) * h[i] + zz * tanhf(c[i]);
        }
        H8 w;
        w.from_float(o);
        st8(hnew + p * C + c0, w);
    }
}

// dh' [P][C] -> drz[:, C:] (z half), dcpre [P][C], dhd [P][C] = dh' (1 - z)
__global__ void __launch_bounds__(256)
gru_gate2_bwd_kernel(const __half* __restrict__ dhn, const __half* __restrict__ rz, const __half* __restrict__ cpre,
                     const __half* __restrict__ cat1, __half* __restrict__ drz, __half* __restrict__ dcpre,
                     __half* __restrict__ dhd, size_t P, int C) {
    mg::pdl_prologue();
    const int G = C >> 3;
    K11_FOR_EACH(P, G) {
        const size_t p = v / G;
        const int c0 = (int)(v - p * G) << 3;
        float g[8], z[8], c[8], h[8], oz[8], oc[8], oh[8];
        ld8(dhn + p * C + c0).to_float(g), ld8(rz + p * 2 * C + C + c0).to_float(z), ld8(cpre + p * C + c0).to_float(c);
        ld8(cat1 + p * 2 * C + C + c0).to_float(h);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float zz = sigmoidf_(z[i]), cc = tanhf(c[i]);
            oz[i] = g[i] * (cc - h[i]) * zz * (1.f - zz);
            oc[i] = g[i] * zz * (1.f - cc * cc);
            oh[i] = g[i] * (1.f - zz);
        }
        H8 w;
        w.from_float(oz), st8(drz + p * 2 * C + C + c0, w);
        w.from_float(oc), st8(dcpre + p * C + c0, w);
        w.from_float(oh), st8(dhd + p * C + c0, w);
    }
}

// dcat2 [P][2C] -> drz[:, :C] (r half), dpart [P][2C] = [dcat2[:, :C] | dhd + dcat2[:, C:] * r]
__global__ void __launch_bounds__(256)
gru_gate1_bwd_kernel(const __half* __restrict__ dcat2, const __half* __restrict__ rz, const __half* __restrict__ cat1,
                     const __half* __restrict__ dhd, __half* __restrict__ drz, __half* __restrict__ dpart, size_t P, int C) {
    mg::pdl_prologue();
    const int G = C >> 3;
    K11_FOR_EACH(P, G) {
        const size_t p = v / G;
        const int c0 = (int)(v - p * G) << 3;
        const size_t row = p * 2 * C;
        st8(dpart + row + c0, ld8(dcat2 + row + c0));
        float d[8], r[8], h[8], hd[8], orr[8], oh[8];
        ld8(dcat2 + row + C + c0).to_float(d), ld8(rz + row + c0).to_float(r), ld8(cat1 + row + C + c0).to_float(h);
        ld8(dhd + p * C + c0).to_float(hd);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float rr = sigmoidf_(r[i]);
            orr[i] = d[i] * h[i] * rr * (1.f - rr);
            oh[i] = hd[i] + d[i] * rr;
        }
        H8 w;
        w.from_float(orr), st8(drz + row + c0, w);
        w.from_float(oh), st8(dpart + row + C + c0, w);
    }
}

// ------------------------------------------------------------------------------------------------ temporal fusion
constexpr int MAX_F = 8;

// fd / bd [B][F][HW] logits (fd[:, 0] and bd[:, F-1] unused), preds / fused [B][F][n_i][HW] fp32
__global__ void __launch_bounds__(256)
temporal_fuse_fwd_kernel(const float* __restrict__ fd, const float* __restrict__ bd, const float* __restrict__ preds,
                         float* __restrict__ fused, int B, int F, int n_i, size_t HW) {
    mg::pdl_prologue();
    const size_t total = (size_t)B * HW;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const size_t b = v / HW, px = v - b * HW;
        float sf[MAX_F], sb[MAX_F];
#pragma unroll
        for (int i = 0; i < MAX_F; ++i) {
            sf[i] = (i >= 1 && i < F) ? sigmoidf_(__ldg(fd + (b * F + i) * HW + px)) : 0.f;
            sb[i] = (i < F - 1) ? sigmoidf_(__ldg(bd + (b * F + i) * HW + px)) : 0.f;
        }
        for (int j = 0; j < n_i; ++j) {
            float p[MAX_F], fp[MAX_F];
#pragma unroll
            for (int i = 0; i < MAX_F; ++i) p[i] = i < F ? __ldg(preds + ((b * F + i) * n_i + j) * HW + px) : 0.f;
            fp[0] = p[0];
#pragma unroll
            for (int i = 1; i < MAX_F; ++i) fp[i] = i < F ? fp[i - 1] * (1.f - sf[i]) + p[i] * sf[i] : 0.f;
            float bp = p[F - 1];
            fused[((b * F + F - 1) * n_i + j) * HW + px] = bp;
#pragma unroll
            for (int i = MAX_F - 2; i >= 0; --i) {
                if (i < F - 1) {
                    bp = bp * (1.f - sb[i]) + p[i] * sb[i];
                    fused[((b * F + i) * n_i + j) * HW + px] = i == 0 ? fp[0] : 0.5f * (fp[i] + bp);
                }
            }
        }
    }
}

// g = d fused -> dpreds, dfd, dbd (logit gradients; planes fd[:, 0] / bd[:, F-1] get zeros)
__global__ void __launch_bounds__(256)
temporal_fuse_bwd_kernel(const float* __restrict__ fd, const float* __restrict__ bd, const float* __restrict__ preds,
                         const float* __restrict__ g, float* __restrict__ dpreds, float* __restrict__ dfd,
                         float* __restrict__ dbd, int B, int F, int n_i, size_t HW) {
    mg::pdl_prologue();
    const size_t total = (size_t)B * HW;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const size_t b = v / HW, px = v - b * HW;
        float sf[MAX_F], sb[MAX_F], dsf[MAX_F], dsb[MAX_F];
#pragma unroll
        for (int i = 0; i < MAX_F; ++i) {
            sf[i] = (i >= 1 && i < F) ? sigmoidf_(__ldg(fd + (b * F + i) * HW + px)) : 0.f;
            sb[i] = (i < F - 1) ? sigmoidf_(__ldg(bd + (b * F + i) * HW + px)) : 0.f;
            dsf[i] = dsb[i] = 0.f;
        }
        for (int j = 0; j < n_i; ++j) {
            float p[MAX_F], fp[MAX_F], bp[MAX_F], gf[MAX_F], gb[MAX_F], dp[MAX_F];
#pragma unroll
            for (int i = 0; i < MAX_F; ++i) {
                const size_t o = ((b * F + i) * n_i + j) * HW + px;
                p[i] = i < F ? __ldg(preds + o) : 0.f;
                const float gi = i < F ? __ldg(g + o) : 0.f;
                // fused_0 = fp_0, fused_{F-1} = bp_{F-1}, the others average the two chains
                gf[i] = i == 0 ? gi : (i < F - 1 ? 0.5f * gi : 0.f);
                gb[i] = i == F - 1 ? gi : (i >= 1 && i < F - 1 ? 0.5f * gi : 0.f);
                dp[i] = 0.f;
            }
            fp[0] = p[0];
#pragma unroll
            for (int i = 1; i < MAX_F; ++i) fp[i] = i < F ? fp[i - 1] * (1.f - sf[i]) + p[i] * sf[i] : 0.f;
#pragma unroll
            for (int i = MAX_F - 1; i >= 0; --i) bp[i] = i == F - 1 ? p[i] : (i < F - 1 ? bp[i + 1 < MAX_F ? i + 1 : i] * (1.f - sb[i]) + p[i] * sb[i] : 0.f);
#pragma unroll
            for (int i = MAX_F - 1; i >= 1; --i) {
                if (i < F) {
                    dp[i] += gf[i] * sf[i];
                    dsf[i] += gf[i] * (p[i] - fp[i - 1]);
                    gf[i - 1] += gf[i] * (1.f - sf[i]);
                }
            }
            dp[0] += gf[0];
#pragma unroll
            for (int i = 0; i < MAX_F - 1; ++i) {
                if (i < F - 1) {
                    dp[i] += gb[i] * sb[i];
                    dsb[i] += gb[i] * (p[i] - bp[i + 1]);
                    gb[i + 1] += gb[i] * (1.f - sb[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < MAX_F; ++i) {
                if (i == F - 1) dp[i] += gb[i];
                if (i < F) dpreds[((b * F + i) * n_i + j) * HW + px] = dp[i];
            }
        }
#pragma unroll
        for (int i = 0; i < MAX_F; ++i) {
            if (i < F) {
                dfd[(b * F + i) * HW + px] = dsf[i] * sf[i] * (1.f - sf[i]);
                dbd[(b * F + i) * HW + px] = dsb[i] * sb[i] * (1.f - sb[i]);
            }
        }
    }
}

int grid_for(size_t items) { return (int)std::min<size_t>((items + 255) / 256, (size_t)mg::kNumSMs * 8); }

}  // namespace

extern "C" int mg_gru_concat2(const void* x, const void* h, void* cat1, size_t P, int C, void* stream) {
    MG_REQUIRE(x && h && cat1 && C % 8 == 0, "mg_gru_concat2: null pointer or C %% 8 != 0");
    if (P == 0) return MG_OK;
    MG_LAUNCH(gru_concat2_kernel, grid_for(P * (C / 4)), 256, 0, stream, static_cast<const __half*>(x), static_cast<const __half*>(h),
              static_cast<__half*>(cat1), P, C);
    MG_CHECK_LAUNCH("mg_gru_concat2");
    return MG_OK;
}

extern "C" int mg_gru_gate1_fwd(const void* rz, const void* cat1, void* cat2, size_t P, int C, void* stream) {
    MG_REQUIRE(rz && cat1 && cat2 && C % 8 == 0, "mg_gru_gate1_fwd: null pointer or C %% 8 != 0");
    if (P == 0) return MG_OK;
    MG_LAUNCH(gru_gate1_fwd_kernel, grid_for(P * (C / 8)), 256, 0, stream, static_cast<const __half*>(rz),
              static_cast<const __half*>(cat1), static_cast<__half*>(cat2), P, C);
    MG_CHECK_LAUNCH("mg_gru_gate1_fwd");
    return MG_OK;
}

extern "C" int mg_gru_gate2_fwd(const void* rz, const void* c_pre, const void* cat1, void* h_new, size_t P, int C, void* stream) {
    MG_REQUIRE(rz && c_pre && cat1 && h_new && C % 8 == 0, "mg_gru_gate2_fwd: null pointer or C %% 8 != 0");
    if (P == 0) return MG_OK;
    MG_LAUNCH(gru_gate2_fwd_kernel, grid_for(P * (C / 8)), 256, 0, stream, static_cast<const __half*>(rz),
              static_cast<const __half*>(c_pre), static_cast<const __half*>(cat1), static_cast<__half*>(h_new), P, C);
    MG_CHECK_LAUNCH("mg_gru_gate2_fwd");
    return MG_OK;
}

extern "C" int mg_gru_gate2_bwd(const void* dh_new, const void* rz, const void* c_pre, const void* cat1, void* drz, void* dc_pre,
                                void* dh_direct, size_t P, int C, void* stream) {
    MG_REQUIRE(dh_new && rz && c_pre && cat1 && drz && dc_pre && dh_direct && C % 8 == 0, "mg_gru_gate2_bwd: null pointer or C %% 8 != 0");
    if (P == 0) return MG_OK;
    MG_LAUNCH(gru_gate2_bwd_kernel, grid_for(P * (C / 8)), 256, 0, stream, static_cast<const __half*>(dh_new),
              static_cast<const __half*>(rz), static_cast<const __half*>(c_pre), static_cast<const __half*>(cat1),
              static_cast<__half*>(drz), static_cast<__half*>(dc_pre), static_cast<__half*>(dh_direct), P, C);
    MG_CHECK_LAUNCH("mg_gru_gate2_bwd");
    return MG_OK;
}

extern "C" int mg_gru_gate1_bwd(const void* dcat2, const void* rz, const void* cat1, const void* dh_direct, void* drz, void* dpart,
                                size_t P, int C, void* stream) {
    MG_REQUIRE(dcat2 && rz && cat1 && dh_direct && drz && dpart && C % 8 == 0, "mg_gru_gate1_bwd: null pointer or C %% 8 != 0");
    if (P == 0) return MG_OK;
    MG_LAUNCH(gru_gate1_bwd_kernel, grid_for(P * (C / 8)), 256, 0, stream, static_cast<const __half*>(dcat2),
              static_cast<const __half*>(rz), static_cast<const __half*>(cat1), static_cast<const __half*>(dh_direct),
              static_cast<__half*>(drz), static_cast<__half*>(dpart), P, C);
    MG_CHECK_LAUNCH("mg_gru_gate1_bwd");
    return MG_OK;
}

extern "C" int mg_temporal_fuse_fwd(const float* fd, const float* bd, const float* preds, float* fused, int B, int F, int n_i,
                                    size_t HW, void* stream) {
    MG_REQUIRE(fd && bd && preds && fused, "mg_temporal_fuse_fwd: null pointer");
    MG_REQUIRE(F >= 2 && F <= MAX_F && n_i >= 1, "mg_temporal_fuse_fwd: 2 <= frames <= %d (got %d)", MAX_F, F);
    if ((size_t)B * HW == 0) return MG_OK;
    MG_LAUNCH(temporal_fuse_fwd_kernel, grid_for((size_t)B * HW), 256, 0, stream, fd, bd, preds, fused, B, F, n_i, HW);
    MG_CHECK_LAUNCH("mg_temporal_fuse_fwd");
    return MG_OK;
}

extern "C" int mg_temporal_fuse_bwd(const float* fd, const float* bd, const float* preds, const float* dfused, float* dpreds,
                                    float* dfd, float* dbd, int B, int F, int n_i, size_t HW, void* stream) {
    MG_REQUIRE(fd && bd && preds && dfused && dpreds && dfd && dbd, "mg_temporal_fuse_bwd: null pointer");
    MG_REQUIRE(F >= 2 && F <= MAX_F && n_i >= 1, "mg_temporal_fuse_bwd: 2 <= frames <= %d (got %d)", MAX_F, F);
    if ((size_t)B * HW == 0) return MG_OK;
    MG_LAUNCH(temporal_fuse_bwd_kernel, grid_for((size_t)B * HW), 256, 0, stream, fd, bd, preds, dfused, dpreds, dfd, dbd, B, F, n_i, HW);
    MG_CHECK_LAUNCH("mg_temporal_fuse_bwd");
    return MG_OK;
}
