// K8a: uncertainty mask = threshold + binary dilation by cv2's MORPH_ELLIPSE(width), bit-packed.
//
// One CTA owns one slice and a band of TH output rows.  Phase 1 thresholds the alpha rows the band needs
// (TH + k - 1) into a bit image in shared memory with warp ballots (one coalesced 128-byte read per
// ballot).  Phase 2: each thread produces one 32-pixel output word; for every structuring-element row it
// builds a 64-bit window of the source row, forms the running OR over the row's span by shift doubling,
// and ORs the aligned 32 bits into its accumulator.  Output is the {0,1} byte image (and/or the bit image).
// HBM-bound: reads 4 B/px once (+ halo rows, mostly L2 hits), writes 1 B/px.
#include "common.cuh"

#include <cstdlib>

namespace {

constexpr int TH_DEFAULT = 64;  // output rows per CTA (run-time `th`: see unknown_mask())
constexpr int MAXK = 29;        // largest ellipse the reference can draw (utils.py:27)
constexpr int THREADS = 256;

// Row span of cv2.getStructuringElement(MORPH_ELLIPSE,(k,k)), row i: columns [j1, j2).
// OpenCV: r = c = k/2; dx = saturate_cast<int>(c*sqrt((r*r-dy*dy)/r^2)) (round half to even).
__device__ __forceinline__ void ellipse_span(int k, int i, int& j1, int& j2) {
    const int r = k / 2, c = k / 2;
    const double inv_r2 = r ? __ddiv_rn(1.0, (double)(r * r)) : 0.0;
    const int dy = i - r;
    const int dx = (int)rint(__dmul_rn((double)c, __dsqrt_rn(__dmul_rn((double)(r * r - dy * dy), inv_r2))));
    j1 = max(c - dx, 0);
    j2 = min(c + dx + 1, k);
}

__device__ __forceinline__ uint64_t run_or(uint64_t v, int w) {
    // d[b] = OR_{t=0..w-1} v[b+t]
    uint64_t d = v;
    int p = 1;
    while (2 * p <= w) {
        d |= d >> p;
        p *= 2;
    }
    if (w > p) d |= d >> (w - p);
    return d;
}

__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t n) {
    return (n & 1u) | ((n & 2u) << 7) | ((n & 4u) << 14) | ((n & 8u) << 21);
}

// alt / use_alt (optional): when *use_alt != 0 the slices are read from `alt` instead of `alpha` - the device-side form of
// the reference's warm-up switch "guide the detail stage with the ground truth when the predicted alpha is all zero"
// (decoder/resnet_inst_matt_spconv.py:311-316) without a host read of the sum.
__global__ void __launch_bounds__(THREADS)
unknown_mask_kernel(const float* __restrict__ alpha, const float* __restrict__ alt, const int32_t* __restrict__ use_alt,
                    int H, int W, const int32_t* __restrict__ widths, const uint8_t* __restrict__ and_mask,
                    uint8_t* __restrict__ out_u8, uint32_t* __restrict__ out_bits, const float* __restrict__ blend_x,
                    const float* __restrict__ blend_y, float* __restrict__ blend_out, int TH) {
    mg::pdl_prologue();
    if (alt && use_alt && *use_alt != 0) alpha = alt;
    extern __shared__ uint32_t sbits[];  // [TH + MAXK - 1][Wd + 2]
    __shared__ int s_lo[MAXK], s_w[MAXK];

    const int s = blockIdx.y, y0 = blockIdx.x * TH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int k = widths[s];
    k = min(max(k, 1), MAXK);
    const int a = k / 2;
    const int Wd = (W + 31) >> 5, RW = Wd + 2;
    const int R = min(TH, H - y0) + k - 1;

    if (tid < k) {
        int j1, j2;
        ellipse_span(k, tid, j1, j2);
        s_lo[tid] = j1 - a;
        s_w[tid] = j2 - j1;
    }
    const float lower = 1.0f / 255.0f, upper = 254.0f / 255.0f;  // fp32-rounded, as torch compares
    const float* aslice = alpha + (size_t)s * H * W;
    if ((W & 31) == 0) {
        // vectorised threshold: a thread owns 8 consecutive pixels (two 16-byte loads in flight per item, items of
        // several rows interleaved), four neighbouring lanes assemble one 32-pixel word with two shuffles
        const int per_row = Wd * 4, total = R * per_row;
        for (int base = warp * 32; base < total; base += THREADS) {
            const int idx = base + lane;
            uint32_t m = 0u;
            int r = 0, q = 0;
            if (idx < total) {
                r = idx / per_row, q = idx - r * per_row;
                const int gy = y0 - a + r;
                if (gy >= 0 && gy < H) {
                    const float4* p = reinterpret_cast<const float4*>(aslice + (size_t)gy * W + q * 8);
                    const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
                    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) m |= (uint32_t)((v[i] > lower) && (v[i] < upper)) << i;
                }
                m <<= 8 * (q & 3);
            }
            m |= __shfl_xor_sync(0xffffffffu, m, 1);
            m |= __shfl_xor_sync(0xffffffffu, m, 2);
            if (idx < total) {
                uint32_t* row = sbits + r * RW;
                if ((q & 3) == 0) row[1 + (q >> 2)] = m;
                if (q == 0) row[0] = 0u, row[Wd + 1] = 0u;
            }
        }
    } else
    for (int r = warp; r < R; r += THREADS / 32) {
        const int gy = y0 - a + r;
        uint32_t* row = sbits + r * RW;
        if (lane == 0) row[0] = 0u, row[Wd + 1] = 0u;
        const bool inb = (gy >= 0) && (gy < H);
        const float* arow = aslice + (size_t)(inb ? gy : 0) * W;
#pragma unroll 4
        for (int j = 0; j < Wd; ++j) {
            const int x = (j << 5) + lane;
            float v = 0.f;
            if (inb && x < W) v = __ldg(arow + x);
            const uint32_t b = __ballot_sync(0xffffffffu, (v > lower) && (v < upper));
            if (lane == 0) row[1 + j] = b;
        }
    }
    __syncthreads();

    const int rows_out = min(TH, H - y0);
    const bool vec = (W & 31) == 0;
    for (int item = tid; item < rows_out * Wd; item += THREADS) {
        const int r = item / Wd, j = item - r * Wd;
        uint32_t acc = 0u;
        for (int i = 0; i < k; ++i) {
            const uint32_t* row = sbits + (r + i) * RW + j;
            const uint64_t v = (uint64_t)(row[0] >> 16) | ((uint64_t)row[1] << 16) | ((uint64_t)row[2] << 48);
            acc |= (uint32_t)(run_or(v, s_w[i]) >> (16 + s_lo[i]));
        }
        const int x0 = j << 5;
        if (x0 + 32 > W) acc &= (1u << (W - x0)) - 1u;
        const size_t off = ((size_t)s * H + (y0 + r)) * W + x0;
        if (and_mask) {
            uint32_t m = 0u;
            if (vec) {
                const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(and_mask + off));
                const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(and_mask + off) + 1);
                const uint32_t wds[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t t = wds[q];
                    m |= (((t & 0xffu) != 0) | (((t & 0xff00u) != 0) << 1) | (((t & 0xff0000u) != 0) << 2) |
                          (((t & 0xff000000u) != 0) << 3)) << (4 * q);
                }
            } else {
                for (int b = 0; b < 32 && x0 + b < W; ++b) m |= (uint32_t)(and_mask[off + b] != 0) << b;
            }
            acc &= m;
        }
        if (out_bits) out_bits[((size_t)s * H + (y0 + r)) * Wd + j] = acc;
        if (blend_out) {
            // K10 (progressive fusion, decoder/resnet_inst_matt_spconv.py:272-290): the {0,1} mask selects between the finer
            // and the coarser alpha - `x * w + y * (1 - w)` of the reference - in the kernel that produces the mask
            if (vec) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(blend_x + off) + q);
                    const float4 yv = __ldg(reinterpret_cast<const float4*>(blend_y + off) + q);
                    const uint32_t b = acc >> (4 * q);
                    reinterpret_cast<float4*>(blend_out + off)[q] =
                        make_float4((b & 1u) ? xv.x : yv.x, (b & 2u) ? xv.y : yv.y, (b & 4u) ? xv.z : yv.z, (b & 8u) ? xv.w : yv.w);
                }
            } else {
                for (int b = 0; b < 32 && x0 + b < W; ++b) blend_out[off + b] = ((acc >> b) & 1u) ? blend_x[off + b] : blend_y[off + b];
            }
        }
        if (out_u8) {
            if (vec) {
                uint4 o0, o1;
                o0.x = nibble_to_bytes(acc), o0.y = nibble_to_bytes(acc >> 4), o0.z = nibble_to_bytes(acc >> 8),
                o0.w = nibble_to_bytes(acc >> 12), o1.x = nibble_to_bytes(acc >> 16), o1.y = nibble_to_bytes(acc >> 20),
                o1.z = nibble_to_bytes(acc >> 24), o1.w = nibble_to_bytes(acc >> 28);
                reinterpret_cast<uint4*>(out_u8 + off)[0] = o0;
                reinterpret_cast<uint4*>(out_u8 + off)[1] = o1;
            } else {
                for (int b = 0; b < 32 && x0 + b < W; ++b) out_u8[off + b] = (acc >> b) & 1u;
            }
        }
    }
}

}  // namespace

static int unknown_mask(const float* alpha, const float* alt, const int32_t* use_alt, int slices, int H, int W,
                        const int32_t* widths, const uint8_t* and_mask, uint8_t* out_u8, uint32_t* out_bits, void* stream,
                        const float* blend_x = nullptr, const float* blend_y = nullptr, float* blend_out = nullptr) {
    MG_REQUIRE(alpha && widths && (out_u8 || out_bits), "mg_unknown_mask: null pointer");
    MG_REQUIRE(slices >= 0 && H > 0 && W > 0 && W <= 4096, "mg_unknown_mask: bad shape %d x %d x %d", slices, H, W);
    if (slices == 0) return MG_OK;
    MG_REQUIRE(slices <= 65535, "mg_unknown_mask: too many slices (%d)", slices);
    const int Wd = (W + 31) / 32;
    // rows per CTA: enough CTAs to keep several resident per SM (the threshold phase and the dilation phase of ONE CTA do
    // not overlap; co-resident CTAs make them overlap), at the price of re-reading the k - 1 halo rows more often (L2 hits)
    static const int th_env = [] { const char* e = std::getenv("MAGGIE_B200_UNKNOWN_TH"); return e ? std::atoi(e) : 0; }();
    int TH = th_env > 0 ? th_env : TH_DEFAULT;
    if (th_env <= 0)
        while (TH > 16 && (long long)mg::ceil_div(H, TH) * slices < 5LL * mg::kNumSMs) TH >>= 1;   // measured: 24 planes 36.9 -> 18.4 us, 80 planes 55.9 -> 47.4 us
    const size_t smem = (size_t)(TH + MAXK - 1) * (Wd + 2) * sizeof(uint32_t);
    dim3 grid(mg::ceil_div(H, TH), slices);
    MG_LAUNCH(unknown_mask_kernel, grid, THREADS, smem, stream, alpha, alt, use_alt, H, W, widths, and_mask, out_u8, out_bits,
              blend_x, blend_y, blend_out, TH);
    MG_CHECK_LAUNCH("mg_unknown_mask");
    return MG_OK;
}

extern "C" int mg_unknown_mask(const float* alpha, int slices, int H, int W, const int32_t* widths,
                               const uint8_t* and_mask, uint8_t* out_u8, uint32_t* out_bits, void* stream) {
    return unknown_mask(alpha, nullptr, nullptr, slices, H, W, widths, and_mask, out_u8, out_bits, stream);
}

extern "C" int mg_unknown_mask_select(const float* alpha, const float* alt, const int32_t* use_alt, int slices, int H, int W,
                                      const int32_t* widths, const uint8_t* and_mask, uint8_t* out_u8, uint32_t* out_bits,
                                      void* stream) {
    MG_REQUIRE(alt && use_alt, "mg_unknown_mask_select: null pointer");
    return unknown_mask(alpha, alt, use_alt, slices, H, W, widths, and_mask, out_u8, out_bits, stream);
}

extern "C" int mg_fuse_stage(const float* src, const float* finer, const float* coarser, int slices, int H, int W,
                             const int32_t* widths, const uint8_t* and_mask, uint8_t* out_w_u8, float* out_alpha, void* stream) {
    MG_REQUIRE(finer && coarser && out_alpha && out_w_u8, "mg_fuse_stage: null pointer");
    return unknown_mask(src, nullptr, nullptr, slices, H, W, widths, and_mask, out_w_u8, nullptr, stream, finer, coarser,
                        out_alpha);
}
