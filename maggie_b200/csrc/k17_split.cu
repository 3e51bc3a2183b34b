// K17: operand split for the fp32-accurate ("x3") evaluation mode of the tcgen05 kernels.
//   x (fp32) -> hi = fp16(x), lo = fp16(x - hi):  x = hi + lo up to 2^-22 |x| (or 2^-24 absolute: fp16 subnormals).
// The tensor cores then evaluate x*w as hi*hi + lo*hi + hi*lo in one fp32 TMEM accumulator (k2_conv.cu, HP = true).
// HBM-bound streaming kernel: 4 B read + 4 B written per element.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
split_f32_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, size_t n4) {
    mg::pdl_prologue();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
        uint2 a, b;
        a.x = *reinterpret_cast<const uint32_t*>(&h0), a.y = *reinterpret_cast<const uint32_t*>(&h1);
        b.x = *reinterpret_cast<const uint32_t*>(&l0), b.y = *reinterpret_cast<const uint32_t*>(&l1);
        hi[i] = a, lo[i] = b;
    }
}

}  // namespace

extern "C" int mg_split_f32(const float* x, void* hi_f16, void* lo_f16, size_t n, void* stream) {
    if (n == 0) return MG_OK;
    MG_REQUIRE(x && hi_f16 && lo_f16, "mg_split_f32: null pointer");
    MG_REQUIRE(n % 4 == 0, "mg_split_f32: element count must be a multiple of 4 (got %zu)", n);
    const size_t n4 = n / 4;
    const int grid = (int)std::min<size_t>((n4 + 255) / 256, (size_t)mg::kNumSMs * 8);
    MG_LAUNCH(split_f32_kernel, grid, 256, 0, stream, reinterpret_cast<const float4*>(x), static_cast<uint2*>(hi_f16),
              static_cast<uint2*>(lo_f16), n4);
    MG_CHECK_LAUNCH("mg_split_f32");
    return MG_OK;
}
