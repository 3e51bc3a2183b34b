// K0: grouped weight preparation for every dense conv layer of the hot path in a handful of launches.
//
// Forward  (3 kernels over ALL layers): spectral-norm power iteration (module/spectral_norm.py:22-35:
//     v <- normalize(W^T u), u <- normalize(W v), sigma = u . W v), then W / sigma is written as fp16 in the two
//     operand layouts the conv kernels consume:  P [Co][tap][Ci_pad] (fprop / wgrad layout) and
//     D [Ci_pad][tap][Co] (data-gradient layout).  Plain (un-normalised) convs take sigma = 1.  The AvgPool2d(2)+1x1
//     skip convs (encoder/resnet.py:111-116) are emitted as 2x2 stride-2 taps of 0.25 * W.
// Backward (2 kernels over ALL layers): the conv weight gradients G (fp32, P layout, written by K4) are taken back
//     through W = W_bar / sigma (u, v constants):  dW_bar = (G - <G, W_bar>/sigma * u v^T) / sigma, and re-laid-out to
//     the torch layout of the master weights.
// All kernels are HBM streams over the 30 M weights (~120 MB fp32): work items are (layer, tile) pairs from a
// host-built table, so one launch covers the whole model.
#include "common.cuh"

namespace {

constexpr int TA = 16, TB = 32, MAX_TAPS = 16;     // pack tile: TA indices of dim0 x TB indices of dim1 (x taps)
constexpr int SROW = TB * MAX_TAPS + 1;
constexpr float kEps = 1e-12f;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// ---- A: v_raw = W^T u   (64 columns per CTA, 4 row groups, fixed-order reduction: no atomics) ----------------------
// Deterministic on purpose: sigma feeds the fp16 rounding of every weight, and an order-dependent last bit here shows up
// as one-ulp flips of deep activations that the network amplifies to ~1e-2 in alpha from one run to the next.
__global__ void __launch_bounds__(256)
wprep_vt_kernel(const mg_wprep_layer* __restrict__ layers, const int4* __restrict__ items, float* __restrict__ vec) {
    mg::pdl_prologue();
    __shared__ float s_part[4][64];
    const int4 it = items[blockIdx.x];
    const mg_wprep_layer L = layers[it.x];
    const int d0 = L.transposed ? L.Ci : L.Co, d1 = L.transposed ? L.Co : L.Ci;
    const int width = d1 * L.taps;
    if ((width & 3) == 0 && (reinterpret_cast<uintptr_t>(L.w) & 15) == 0 && (reinterpret_cast<uintptr_t>(vec + L.vec_off) & 15) == 0) {
        // 16 threads x 4 columns, 16 row groups: 16-byte loads, the same fixed summation order on every run
        __shared__ float4 s_p4[16][16];
        const int c4 = threadIdx.x & 15, rg = threadIdx.x >> 4, col = it.z + 4 * c4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < width) {
#pragma unroll 4
            for (int r = rg; r < d0; r += 16) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(L.w + (size_t)r * width + col));
                const float u = __ldg(L.u + r);
                acc.x += w.x * u, acc.y += w.y * u, acc.z += w.z * u, acc.w += w.w * u;
            }
        }
        s_p4[rg][c4] = acc;
        __syncthreads();
        if (rg == 0 && col < width) {
            float4 t = s_p4[0][c4];
#pragma unroll
            for (int g = 1; g < 16; ++g) {
                const float4 q = s_p4[g][c4];
                t.x += q.x, t.y += q.y, t.z += q.z, t.w += q.w;
            }
            *reinterpret_cast<float4*>(vec + L.vec_off + col) = t;
        }
        return;
    }
    const int cx = threadIdx.x & 63, rg = threadIdx.x >> 6, col = it.z + cx;
    float acc = 0.f;
    if (col < width) {
#pragma unroll 4
        for (int r = rg; r < d0; r += 4) acc += L.w[(size_t)r * width + col] * __ldg(L.u + r);
    }
    s_part[rg][cx] = acc;
    __syncthreads();
    if (rg == 0 && col < width) vec[L.vec_off + col] = (s_part[0][cx] + s_part[1][cx]) + (s_part[2][cx] + s_part[3][cx]);
}

// ---- B: t = W v_raw / (|v_raw| + eps)   (one warp per row) --------------------------------------------------------
__global__ void __launch_bounds__(256)
wprep_u_kernel(const mg_wprep_layer* __restrict__ layers, const int4* __restrict__ items, float* __restrict__ vec) {
    mg::pdl_prologue();
    const int4 it = items[blockIdx.x];
    const mg_wprep_layer L = layers[it.x];
    const int d0 = L.transposed ? L.Ci : L.Co, d1 = L.transposed ? L.Co : L.Ci;
    const int width = d1 * L.taps, row = it.y + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= d0) return;
    const float* vr = vec + L.vec_off;
    const float* wr = L.w + (size_t)row * width;
    float acc = 0.f, nn = 0.f;
    if ((width & 3) == 0 && ((reinterpret_cast<uintptr_t>(wr) | reinterpret_cast<uintptr_t>(vr)) & 15) == 0) {
#pragma unroll 2
        for (int j = lane; j < (width >> 2); j += 32) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(wr) + j);
            const float4 x = *(reinterpret_cast<const float4*>(vr) + j);
            acc += (w.x * x.x + w.y * x.y) + (w.z * x.z + w.w * x.w);
            nn += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
        }
    } else
    for (int j = lane; j < width; j += 32) {
        const float x = vr[j];
        acc += wr[j] * x;
        nn += x * x;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, d);
        nn += __shfl_xor_sync(0xffffffffu, nn, d);
    }
    if (lane == 0) vec[L.vec_off + width + row] = acc / (sqrtf(nn) + kEps);
}

// The tile bodies are instantiated for the tap counts of the model (1, 9, 16; 0 = run-time value): every index of the
// re-layout loops is a quotient / remainder by `taps`-derived sizes, and with run-time divisors the kernels were bound by
// integer-division instructions (wprep_pack 219 us for a 240 MB stream), not by memory.
// A full tile (no fold) of a layer whose rows are 16-byte aligned takes the vector forms below: 16-byte loads of the master
// weights / packed gradients and 16-byte stores of eight fp16 pack entries (or four fp32 gradients) per thread.  With one
// scalar access and its index arithmetic per element the kernels were bound by instruction issue, not by memory.
__device__ __forceinline__ bool full_tile(const mg_wprep_layer& L, const int4 it, int taps) {
    const int d0 = L.transposed ? L.Ci : L.Co, d1 = L.transposed ? L.Co : L.Ci;
    return !L.fold && it.y + TA <= d0 && it.z + TB <= d1 && ((d1 * taps) & 3) == 0 && (L.ci_pad & 7) == 0 && (L.Co & 7) == 0 &&
           (reinterpret_cast<uintptr_t>(L.w) & 15) == 0;
}

// s[a][e] <- W[a0 + a][b0 * taps + e], e < TB * taps, 16 bytes per load
__device__ __forceinline__ void load_w_tile_v4(const mg_wprep_layer& L, int a0, int b0, int taps, int width, float (*s)[SROW]) {
    const int seg4 = (TB * taps) >> 2;
    for (int i = threadIdx.x; i < TA * seg4; i += 256) {
        const int a = i / seg4, q = i - a * seg4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(L.w + (size_t)(a0 + a) * width + (size_t)b0 * taps) + q);
        float* d = &s[a][4 * q];
        d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
    }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8], float mul) {
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[2 * j] * mul, f[2 * j + 1] * mul);
    return o;
}

// dst[((b0 + b) * to + tp) * ld + a0 + 8h .. +8] = s[8h + j][b * taps + tp]      ("a" is the contiguous index of dst)
template <int TAPS>
__device__ __forceinline__ void store_a_fastest(__half* __restrict__ dst, int ld, int a0, int b0, int taps, float mul,
                                                float (*s)[SROW]) {
    constexpr int HG = TA / 8;
    for (int i = threadIdx.x; i < TB * taps * HG; i += 256) {
        const int h = i % HG, tp = (i / HG) % taps, b = i / (HG * taps);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = s[8 * h + j][b * taps + tp];
        *reinterpret_cast<uint4*>(dst + ((size_t)(b0 + b) * taps + tp) * ld + a0 + 8 * h) = pack8(f, mul);
    }
}

// dst[((a0 + a) * to + tp) * ld + b0 + 8g .. +8] = s[a][(8g + j) * taps + tp]    ("b" is the contiguous index of dst)
template <int TAPS>
__device__ __forceinline__ void store_b_fastest(__half* __restrict__ dst, int ld, int a0, int b0, int taps, float mul,
                                                float (*s)[SROW]) {
    constexpr int BG = TB / 8;
    for (int i = threadIdx.x; i < TA * taps * BG; i += 256) {
        const int g = i % BG, tp = (i / BG) % taps, a = i / (BG * taps);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = s[a][(8 * g + j) * taps + tp];
        *reinterpret_cast<uint4*>(dst + ((size_t)(a0 + a) * taps + tp) * ld + b0 + 8 * g) = pack8(f, mul);
    }
}

template <int TAPS>
__device__ __forceinline__ void pack_tile(const mg_wprep_layer& L, const int4 it, const float inv_sigma, float (*s)[SROW],
                                          __half* __restrict__ P, __half* __restrict__ D) {
    const int d0 = L.transposed ? L.Ci : L.Co, d1 = L.transposed ? L.Co : L.Ci;
    const int taps = TAPS ? TAPS : L.taps, width = d1 * taps, tid = threadIdx.x;
    const int a0 = it.y, b0 = it.z, seg = TB * taps;
    if (full_tile(L, it, taps)) {
        load_w_tile_v4(L, a0, b0, taps, width, s);
        __syncthreads();
        __half* Pl = P + L.p_off;
        __half* Dl = D ? D + L.d_off : nullptr;
        if (!L.transposed) {   // a = co, b = ci:  P [co][tap][ci_pad],  D [ci][tap][Co]
            store_b_fastest<TAPS>(Pl, L.ci_pad, a0, b0, taps, inv_sigma, s);
            if (Dl) store_a_fastest<TAPS>(Dl, L.Co, a0, b0, taps, inv_sigma, s);
        } else {               // a = ci, b = co
            store_a_fastest<TAPS>(Pl, L.ci_pad, a0, b0, taps, inv_sigma, s);
            if (Dl) store_b_fastest<TAPS>(Dl, L.Co, a0, b0, taps, inv_sigma, s);
        }
        return;
    }
    for (int i = tid; i < TA * seg; i += 256) {
        const int a = i / seg, e = i - a * seg;
        const int col = b0 * taps + e;
        s[a][e] = (a0 + a < d0 && col < width) ? L.w[(size_t)(a0 + a) * width + col] : 0.f;
    }
    __syncthreads();
    const float mul = inv_sigma * (L.fold ? 0.25f : 1.f);
    const int to = TAPS > 1 ? TAPS : (L.fold ? 4 : taps), Co = L.Co, cip = L.ci_pad;   // (only 1x1 layers fold)
    __half* Pl = P + L.p_off;
    __half* Dl = D ? D + L.d_off : nullptr;
    if (!L.transposed) {   // a = co, b = ci
        for (int i = tid; i < TA * to * TB; i += 256) {             // P: ci fastest
            const int b = i % TB, tp = (i / TB) % to, a = i / (TB * to);
            const int co = a0 + a, ci = b0 + b;
            if (co < Co && ci < cip) Pl[((size_t)co * to + tp) * cip + ci] = __float2half(s[a][b * taps + (L.fold ? 0 : tp)] * mul);
        }
        if (Dl)
            for (int i = tid; i < TA * to * TB; i += 256) {         // D: co fastest
                const int a = i % TA, tp = (i / TA) % to, b = i / (TA * to);
                const int co = a0 + a, ci = b0 + b;
                if (co < Co && ci < cip) Dl[((size_t)ci * to + tp) * Co + co] = __float2half(s[a][b * taps + (L.fold ? 0 : tp)] * mul);
            }
    } else {               // a = ci, b = co
        for (int i = tid; i < TA * to * TB; i += 256) {             // P: ci fastest
            const int a = i % TA, tp = (i / TA) % to, b = i / (TA * to);
            const int ci = a0 + a, co = b0 + b;
            if (co < Co && ci < cip) Pl[((size_t)co * to + tp) * cip + ci] = __float2half(s[a][b * taps + tp] * mul);
        }
        if (Dl)
            for (int i = tid; i < TA * to * TB; i += 256) {         // D: co fastest
                const int b = i % TB, tp = (i / TB) % to, a = i / (TB * to);
                const int ci = a0 + a, co = b0 + b;
                if (co < Co && ci < cip) Dl[((size_t)ci * to + tp) * Co + co] = __float2half(s[a][b * taps + tp] * mul);
            }
    }
}

// ---- C: sigma, in-place u / v update, fp16 packs -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
wprep_pack_kernel(const mg_wprep_layer* __restrict__ layers, const int4* __restrict__ items, const float* __restrict__ vec,
                  float* __restrict__ scal, __half* __restrict__ P, __half* __restrict__ D) {
    mg::pdl_prologue();
    __shared__ float s[TA][SROW];
    __shared__ float red[8];
    const int4 it = items[blockIdx.x];
    const mg_wprep_layer L = layers[it.x];
    const int d0 = L.transposed ? L.Ci : L.Co, d1 = L.transposed ? L.Co : L.Ci;
    const int taps = L.taps, width = d1 * taps, tid = threadIdx.x;
    float inv_sigma = 1.f;
    if (L.u) {
        const float* vr = vec + L.vec_off;
        const float* t = vr + width;
        float tt = 0.f;
        for (int i = tid; i < d0; i += 256) tt += t[i] * t[i];
        tt = block_sum(tt, red);
        const float tn = sqrtf(tt);
        inv_sigma = (tn + kEps) / tt;                     // sigma = u . t = |t|^2 / (|t| + eps)
        if (it.w) {                                       // first tile of the layer: publish u, v and the scalars
            float vv = 0.f;
            for (int j = tid; j < width; j += 256) vv += vr[j] * vr[j];
            vv = block_sum(vv, red);
            const float ivn = 1.f / (sqrtf(vv) + kEps), itn = 1.f / (tn + kEps);
            for (int j = tid; j < width; j += 256) L.v[j] = vr[j] * ivn;
            for (int i = tid; i < d0; i += 256) L.u[i] = t[i] * itn;
            if (tid == 0) scal[it.x * 4 + 0] = tt * itn, scal[it.x * 4 + 1] = ivn, scal[it.x * 4 + 2] = itn, scal[it.x * 4 + 3] = 0.f;
        }
    } else if (it.w && tid == 0) {
        scal[it.x * 4 + 0] = 1.f, scal[it.x * 4 + 1] = 0.f, scal[it.x * 4 + 2] = 0.f, scal[it.x * 4 + 3] = 0.f;
    }
    switch (L.taps) {
        case 1: pack_tile<1>(L, it, inv_sigma, s, P, D); break;
        case 9: pack_tile<9>(L, it, inv_sigma, s, P, D); break;
        case 16: pack_tile<16>(L, it, inv_sigma, s, P, D); break;
        default: pack_tile<0>(L, it, inv_sigma, s, P, D); break;
    }
}

// Loads the tile of dL/dW (torch layout order) from the packed gradient G [Co][taps_out][ci_pad] into s[a][b*taps+tap].
template <int TAPS>
__device__ __forceinline__ void load_grad_tile(const mg_wprep_layer& L, const float* __restrict__ G, int a0, int b0,
                                               float (*s)[SROW]) {
    const int taps = TAPS ? TAPS : L.taps, to = TAPS > 1 ? TAPS : (L.fold ? 4 : taps), Co = L.Co, Ci = L.Ci, cip = L.ci_pad, tid = threadIdx.x;
    const float* Gl = G + L.g_off;
    if (full_tile(L, make_int4(0, a0, b0, 0), taps) && (reinterpret_cast<uintptr_t>(Gl) & 15) == 0) {
        if (!L.transposed) {   // G [co = a0 + a][tap][ci = b0 + b]: b contiguous
            constexpr int Q = TB / 4;
            for (int i = tid; i < TA * taps * Q; i += 256) {
                const int q = i % Q, tp = (i / Q) % taps, a = i / (Q * taps);
                const float4 v = __ldg(reinterpret_cast<const float4*>(Gl + ((size_t)(a0 + a) * taps + tp) * cip + b0) + q);
                float* d = &s[a][(4 * q) * taps + tp];
                d[0] = v.x, d[taps] = v.y, d[2 * taps] = v.z, d[3 * taps] = v.w;
            }
        } else {               // G [co = b0 + b][tap][ci = a0 + a]: a contiguous
            constexpr int Q = TA / 4;
            for (int i = tid; i < TB * taps * Q; i += 256) {
                const int q = i % Q, tp = (i / Q) % taps, b = i / (Q * taps);
                const float4 v = __ldg(reinterpret_cast<const float4*>(Gl + ((size_t)(b0 + b) * taps + tp) * cip + a0) + q);
                const int e = b * taps + tp;
                s[4 * q][e] = v.x, s[4 * q + 1][e] = v.y, s[4 * q + 2][e] = v.z, s[4 * q + 3][e] = v.w;
            }
        }
        return;
    }
    if (!L.transposed) {
        for (int i = tid; i < TA * taps * TB; i += 256) {
            const int b = i % TB, tp = (i / TB) % taps, a = i / (TB * taps);
            const int co = a0 + a, ci = b0 + b;
            float v = 0.f;
            if (co < Co && ci < Ci) {
                if (L.fold) {
                    const float* g = Gl + (size_t)co * 4 * cip + ci;
                    v = 0.25f * (g[0] + g[cip] + g[2 * cip] + g[3 * cip]);
                } else
                    v = Gl[((size_t)co * to + tp) * cip + ci];
            }
            s[a][b * taps + tp] = v;
        }
    } else {
        for (int i = tid; i < TA * taps * TB; i += 256) {
            const int a = i % TA, tp = (i / TA) % taps, b = i / (TA * taps);
            const int ci = a0 + a, co = b0 + b;
            s[a][b * taps + tp] = (co < Co && ci < Ci) ? Gl[((size_t)co * to + tp) * cip + ci] : 0.f;
        }
    }
}

template <int TAPS>
__device__ __forceinline__ float inner_tile(const mg_wprep_layer& L, const int4 it, const float* __restrict__ G, float (*s)[SROW]) {
    const int d0 = L.transposed ? L.Ci : L.Co, d1 = L.transposed ? L.Co : L.Ci;
    const int taps = TAPS ? TAPS : L.taps, width = d1 * taps, seg = TB * taps;
    load_grad_tile<TAPS>(L, G, it.y, it.z, s);
    __syncthreads();
    float acc = 0.f;
    if (full_tile(L, it, taps)) {
        const int seg4 = seg >> 2;
        for (int i = threadIdx.x; i < TA * seg4; i += 256) {
            const int a = i / seg4, q = i - a * seg4;
            const float4 w = __ldg(reinterpret_cast<const float4*>(L.w + (size_t)(it.y + a) * width + (size_t)it.z * taps) + q);
            const float* g = &s[a][4 * q];
            acc += g[0] * w.x + g[1] * w.y + g[2] * w.z + g[3] * w.w;
        }
        return acc;
    }
    for (int i = threadIdx.x; i < TA * seg; i += 256) {
        const int a = i / seg, e = i - a * seg, col = it.z * taps + e;
        if (it.y + a < d0 && col < width) acc += s[a][e] * L.w[(size_t)(it.y + a) * width + col];
    }
    return acc;
}

template <int TAPS>
__device__ __forceinline__ void final_tile(const mg_wprep_layer& L, const int4 it, const float* __restrict__ G,
                                           const float* __restrict__ vec, const float* __restrict__ scal, float* __restrict__ grad,
                                           float (*s)[SROW]) {
    const int d0 = L.transposed ? L.Ci : L.Co, d1 = L.transposed ? L.Co : L.Ci;
    const int taps = TAPS ? TAPS : L.taps, width = d1 * taps, seg = TB * taps;
    load_grad_tile<TAPS>(L, G, it.y, it.z, s);
    __syncthreads();
    float* gl = grad + L.grad_off;
    if (full_tile(L, it, taps) && (reinterpret_cast<uintptr_t>(gl) & 15) == 0) {
        float inv_sigma = 1.f, k = 0.f;
        const float* vr = vec + L.vec_off;
        const float* t = vr + width;
        if (L.u) {
            const float sigma = scal[it.x * 4 + 0], ivn = scal[it.x * 4 + 1], itn = scal[it.x * 4 + 2], inner = scal[it.x * 4 + 3];
            inv_sigma = 1.f / sigma, k = inner * inv_sigma * ivn * itn;
        }
        const int seg4 = seg >> 2;
        for (int i = threadIdx.x; i < TA * seg4; i += 256) {
            const int a = i / seg4, q = i - a * seg4, col = it.z * taps + 4 * q;
            const float* g = &s[a][4 * q];
            float4 o = make_float4(g[0], g[1], g[2], g[3]);
            if (L.u) {
                const float kt = k * t[it.y + a];
                o.x = (o.x - kt * vr[col]) * inv_sigma, o.y = (o.y - kt * vr[col + 1]) * inv_sigma;
                o.z = (o.z - kt * vr[col + 2]) * inv_sigma, o.w = (o.w - kt * vr[col + 3]) * inv_sigma;
            }
            *reinterpret_cast<float4*>(gl + (size_t)(it.y + a) * width + col) = o;
        }
        return;
    }
    if (L.u) {
        const float sigma = scal[it.x * 4 + 0], ivn = scal[it.x * 4 + 1], itn = scal[it.x * 4 + 2], inner = scal[it.x * 4 + 3];
        const float inv_sigma = 1.f / sigma, k = inner * inv_sigma * ivn * itn;
        const float* vr = vec + L.vec_off;
        const float* t = vr + width;
        for (int i = threadIdx.x; i < TA * seg; i += 256) {
            const int a = i / seg, e = i - a * seg, col = it.z * taps + e;
            if (it.y + a < d0 && col < width)
                gl[(size_t)(it.y + a) * width + col] = (s[a][e] - k * t[it.y + a] * vr[col]) * inv_sigma;
        }
    } else {
        for (int i = threadIdx.x; i < TA * seg; i += 256) {
            const int a = i / seg, e = i - a * seg, col = it.z * taps + e;
            if (it.y + a < d0 && col < width) gl[(size_t)(it.y + a) * width + col] = s[a][e];
        }
    }
}

// ---- D: inner[layer] = <dL/dW, W_bar>  (spectral-norm layers only) -------------------------------------------------
__global__ void __launch_bounds__(256)
wprep_bwd_inner_kernel(const mg_wprep_layer* __restrict__ layers, const int4* __restrict__ items, const float* __restrict__ G,
                       float* __restrict__ scal) {
    mg::pdl_prologue();
    __shared__ float s[TA][SROW];
    __shared__ float red[8];
    const int4 it = items[blockIdx.x];
    const mg_wprep_layer L = layers[it.x];
    if (!L.u) return;
    float acc;
    switch (L.taps) {
        case 1: acc = inner_tile<1>(L, it, G, s); break;
        case 9: acc = inner_tile<9>(L, it, G, s); break;
        case 16: acc = inner_tile<16>(L, it, G, s); break;
        default: acc = inner_tile<0>(L, it, G, s); break;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && acc != 0.f) atomicAdd(scal + it.x * 4 + 3, acc);
}

// ---- E: dW_bar in the torch layout ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
wprep_bwd_final_kernel(const mg_wprep_layer* __restrict__ layers, const int4* __restrict__ items, const float* __restrict__ G,
                       const float* __restrict__ vec, const float* __restrict__ scal, float* __restrict__ grad) {
    mg::pdl_prologue();
    __shared__ float s[TA][SROW];
    const int4 it = items[blockIdx.x];
    const mg_wprep_layer L = layers[it.x];
    switch (L.taps) {
        case 1: final_tile<1>(L, it, G, vec, scal, grad, s); break;
        case 9: final_tile<9>(L, it, G, vec, scal, grad, s); break;
        case 16: final_tile<16>(L, it, G, vec, scal, grad, s); break;
        default: final_tile<0>(L, it, G, vec, scal, grad, s); break;
    }
}

}  // namespace

extern "C" int mg_wprep_fwd(const mg_wprep_layer* layers, const int32_t* items_vt, int n_vt, const int32_t* items_u, int n_u,
                            const int32_t* items_tile, int n_tile, float* vec, float* scal, void* P, void* D, void* stream) {
    MG_REQUIRE(layers && items_tile && vec && scal && P, "mg_wprep_fwd: null pointer");
    MG_REQUIRE(n_tile > 0, "mg_wprep_fwd: no work");
    if (n_vt > 0) MG_LAUNCH(wprep_vt_kernel, n_vt, 256, 0, stream, layers, reinterpret_cast<const int4*>(items_vt), vec);
    if (n_u > 0) MG_LAUNCH(wprep_u_kernel, n_u, 256, 0, stream, layers, reinterpret_cast<const int4*>(items_u), vec);
    MG_LAUNCH(wprep_pack_kernel, n_tile, 256, 0, stream, layers, reinterpret_cast<const int4*>(items_tile), vec, scal,
              static_cast<__half*>(P), static_cast<__half*>(D));
    MG_CHECK_LAUNCH("mg_wprep_fwd");
    return MG_OK;
}

extern "C" int mg_wprep_bwd(const mg_wprep_layer* layers, const int32_t* items_tile, int n_tile, const float* G, const float* vec,
                            float* scal, float* grad, void* stream) {
    MG_REQUIRE(layers && items_tile && G && vec && scal && grad, "mg_wprep_bwd: null pointer");
    MG_REQUIRE(n_tile > 0, "mg_wprep_bwd: no work");
    MG_LAUNCH(wprep_bwd_inner_kernel, n_tile, 256, 0, stream, layers, reinterpret_cast<const int4*>(items_tile), G, scal);
    MG_LAUNCH(wprep_bwd_final_kernel, n_tile, 256, 0, stream, layers, reinterpret_cast<const int4*>(items_tile), G, vec,
              const_cast<const float*>(scal), grad);
    MG_CHECK_LAUNCH("mg_wprep_bwd");
    return MG_OK;
}
