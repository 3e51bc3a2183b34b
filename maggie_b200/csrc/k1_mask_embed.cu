// K1: mask-id embedding front-end fused with the NCHW fp32 -> NHWC fp16 pack of the encoder input.
// One thread per pixel: 3 image planes + M mask planes are read coalesced along x, the 11x3 table sits
// in shared memory, the packed pixel (C halfs) is written with one vector store.  HBM-bound:
// (3 + M) * 4 B read + 2*C B written per pixel; the reference materialises [B,10,H,W,3] fp32 instead.
#include "common.cuh"

namespace {

constexpr int MAX_M = 16;
struct SlotIds {
    int v[MAX_M];
};

// out32 (optional, evaluation at fp32-level accuracy): the packed pixel is written as C floats instead of C halfs
template <int C>
__global__ void __launch_bounds__(256)
mask_embed_fwd_kernel(const float* __restrict__ image, const float* __restrict__ masks, const int32_t* __restrict__ slot_ids,
                      int M, const float* __restrict__ table, __half* __restrict__ out, float* __restrict__ out32, int B, int HW) {
    mg::pdl_prologue();
    __shared__ float s_tab[33];
    __shared__ SlotIds ids;
    if (threadIdx.x < 33) s_tab[threadIdx.x] = table[threadIdx.x];
    if (threadIdx.x < MAX_M) ids.v[threadIdx.x] = threadIdx.x < M ? min(max(slot_ids[threadIdx.x], 0), 9) : 0;
    __syncthreads();
    const size_t total = (size_t)B * HW;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(p / HW);
        const int q = (int)(p - (size_t)b * HW);
        const float* img = image + (size_t)b * 3 * HW + q;
        const float r = __ldg(img), g = __ldg(img + HW), bl = __ldg(img + 2 * HW);
        float e0 = 0.f, e1 = 0.f, e2 = 0.f, cnt = 0.f;
        const float* mk = masks + (size_t)b * M * HW + q;
        for (int m = 0; m < M; ++m) {
            // reference: id = long(mask * (slot+1)); a slot participates iff id > 0
            const int id = (int)(__ldg(mk + (size_t)m * HW) * (float)(ids.v[m] + 1));
            if (id > 0) {
                const int row = min(id, 10) * 3;
                e0 += s_tab[row], e1 += s_tab[row + 1], e2 += s_tab[row + 2];
                cnt += 1.f;
            }
        }
        const float inv = 1.f / (cnt + 1e-6f);
        if (out32) {
            float* o = out32 + p * C;
            if (C % 4 == 0) {
                reinterpret_cast<float4*>(o)[0] = make_float4(r, g, bl, e0 * inv);
                reinterpret_cast<float4*>(o)[1] = make_float4(e1 * inv, e2 * inv, 0.f, 0.f);
#pragma unroll
                for (int i = 2; i < C / 4; ++i) reinterpret_cast<float4*>(o)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                o[0] = r, o[1] = g, o[2] = bl, o[3] = e0 * inv, o[4] = e1 * inv, o[5] = e2 * inv;
            }
            continue;
        }
        __align__(16) __half2 v[C / 2 < 4 ? 4 : C / 2];
        v[0] = __floats2half2_rn(r, g);
        v[1] = __floats2half2_rn(bl, e0 * inv);
        v[2] = __floats2half2_rn(e1 * inv, e2 * inv);
#pragma unroll
        for (int i = 3; i < C / 2; ++i) v[i] = __floats2half2_rn(0.f, 0.f);
        __half2* o = reinterpret_cast<__half2*>(out + p * C);
        if (C == 8) {
            *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(v);
        } else if (C == 16) {
            reinterpret_cast<uint4*>(o)[0] = reinterpret_cast<uint4*>(v)[0];
            reinterpret_cast<uint4*>(o)[1] = reinterpret_cast<uint4*>(v)[1];
        } else if (C == 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(o)[i] = reinterpret_cast<uint4*>(v)[i];
        } else {
#pragma unroll
            for (int i = 0; i < C / 2; ++i) o[i] = v[i];
        }
    }
}

template <int C>
__global__ void __launch_bounds__(256)
mask_embed_bwd_kernel(const __half* __restrict__ gout, const float* __restrict__ masks, const int32_t* __restrict__ slot_ids,
                      int M, float* __restrict__ gtable, int B, int HW) {
    mg::pdl_prologue();
    __shared__ float s_acc[33];
    __shared__ SlotIds ids;
    if (threadIdx.x < 33) s_acc[threadIdx.x] = 0.f;
    if (threadIdx.x < MAX_M) ids.v[threadIdx.x] = threadIdx.x < M ? min(max(slot_ids[threadIdx.x], 0), 9) : 0;
    __syncthreads();
    float acc[MAX_M][3];
#pragma unroll
    for (int m = 0; m < MAX_M; ++m) acc[m][0] = acc[m][1] = acc[m][2] = 0.f;
    const size_t total = (size_t)B * HW;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(p / HW);
        const int q = (int)(p - (size_t)b * HW);
        const float* mk = masks + (size_t)b * M * HW + q;
        float cnt = 0.f;
        unsigned on = 0u;
#pragma unroll
        for (int m = 0; m < MAX_M; ++m)
            if (m < M && (int)(__ldg(mk + (size_t)m * HW) * (float)(ids.v[m] + 1)) > 0) on |= 1u << m, cnt += 1.f;
        if (!on) continue;
        const float inv = 1.f / (cnt + 1e-6f);
        const __half* gp = gout + p * C + 3;
        const float g0 = __half2float(gp[0]) * inv, g1 = __half2float(gp[1]) * inv, g2 = __half2float(gp[2]) * inv;
#pragma unroll
        for (int m = 0; m < MAX_M; ++m)
            if (on & (1u << m)) acc[m][0] += g0, acc[m][1] += g1, acc[m][2] += g2;
    }
#pragma unroll
    for (int m = 0; m < MAX_M; ++m) {
        if (m >= M) break;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = acc[m][c];
#pragma unroll
            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&s_acc[min(ids.v[m] + 1, 10) * 3 + c], v);
        }
    }
    __syncthreads();
    if (threadIdx.x < 33 && s_acc[threadIdx.x] != 0.f) atomicAdd(gtable + threadIdx.x, s_acc[threadIdx.x]);
}

int check(const int32_t* slot_ids, int M, int C, const char* who) {
    MG_REQUIRE((slot_ids || M == 0) && M >= 0 && M <= MAX_M, "%s: M must be 0..%d (got %d) with slot ids", who, MAX_M, M);
    MG_REQUIRE(C == 6 || C == 8 || C == 16 || C == 32, "%s: C must be 6, 8, 16 or 32 (got %d)", who, C);
    return MG_OK;
}

}  // namespace

static int mask_embed_fwd(const float* image, const float* masks, const int32_t* slot_ids, int M, const float* table,
                          void* out_f16, float* out32, int B, int H, int W, int C, void* stream) {
    MG_REQUIRE(image && table && (out_f16 || out32) && (masks || M == 0), "mg_mask_embed_fwd: null pointer");
    if (int e = check(slot_ids, M, C, "mg_mask_embed_fwd")) return e;
    if (B <= 0) return MG_OK;
    const int HW = H * W;
    const int grid = (int)std::min<size_t>(((size_t)B * HW + 255) / 256, (size_t)mg::kNumSMs * 16);
    __half* out = static_cast<__half*>(out_f16);
    if (C == 6) MG_LAUNCH(mask_embed_fwd_kernel<6>, grid, 256, 0, stream, image, masks, slot_ids, M, table, out, out32, B, HW);
    else if (C == 8) MG_LAUNCH(mask_embed_fwd_kernel<8>, grid, 256, 0, stream, image, masks, slot_ids, M, table, out, out32, B, HW);
    else if (C == 16) MG_LAUNCH(mask_embed_fwd_kernel<16>, grid, 256, 0, stream, image, masks, slot_ids, M, table, out, out32, B, HW);
    else MG_LAUNCH(mask_embed_fwd_kernel<32>, grid, 256, 0, stream, image, masks, slot_ids, M, table, out, out32, B, HW);
    MG_CHECK_LAUNCH("mg_mask_embed_fwd");
    return MG_OK;
}

extern "C" int mg_mask_embed_fwd(const float* image, const float* masks, const int32_t* slot_ids, int M,
                                 const float* table, void* out_f16, int B, int H, int W, int C, void* stream) {
    MG_REQUIRE(out_f16, "mg_mask_embed_fwd: null pointer");
    return mask_embed_fwd(image, masks, slot_ids, M, table, out_f16, nullptr, B, H, W, C, stream);
}

extern "C" int mg_mask_embed_fwd_f32(const float* image, const float* masks, const int32_t* slot_ids, int M,
                                     const float* table, float* out_f32, int B, int H, int W, int C, void* stream) {
    MG_REQUIRE(out_f32, "mg_mask_embed_fwd_f32: null pointer");
    return mask_embed_fwd(image, masks, slot_ids, M, table, nullptr, out_f32, B, H, W, C, stream);
}

extern "C" int mg_mask_embed_bwd(const void* grad_out_f16, const float* masks, const int32_t* slot_ids, int M,
                                 float* grad_table, int B, int H, int W, int C, void* stream) {
    MG_REQUIRE(grad_out_f16 && grad_table && (masks || M == 0), "mg_mask_embed_bwd: null pointer");
    if (int e = check(slot_ids, M, C, "mg_mask_embed_bwd")) return e;
    if (B <= 0 || M == 0) return MG_OK;
    const int HW = H * W;
    const int grid = (int)std::min<size_t>(((size_t)B * HW + 255) / 256, (size_t)mg::kNumSMs * 8);
    const __half* g = static_cast<const __half*>(grad_out_f16);
    if (C == 6) MG_LAUNCH(mask_embed_bwd_kernel<6>, grid, 256, 0, stream, g, masks, slot_ids, M, grad_table, B, HW);
    else if (C == 8) MG_LAUNCH(mask_embed_bwd_kernel<8>, grid, 256, 0, stream, g, masks, slot_ids, M, grad_table, B, HW);
    else if (C == 16) MG_LAUNCH(mask_embed_bwd_kernel<16>, grid, 256, 0, stream, g, masks, slot_ids, M, grad_table, B, HW);
    else MG_LAUNCH(mask_embed_bwd_kernel<32>, grid, 256, 0, stream, g, masks, slot_ids, M, grad_table, B, HW);
    MG_CHECK_LAUNCH("mg_mask_embed_bwd");
    return MG_OK;
}
