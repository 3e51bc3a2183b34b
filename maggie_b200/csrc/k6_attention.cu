// K6: mask-guided attention core of the InstanceMatteDecoder - single head, E = 128, between a FEW side
// (<= 16 instance tokens) and a MANY side (h*w*n_f pixel features, 4096 per 512x512 frame).
//
//   tok <- feat ("tq"): queries = tokens, keys/values = pixels; softmax over the pixels; also emits the attention
//                       statistic  stat[b,q] = sum_k guidance[b,q,k] * A[b,q,k]  that the attention-max loss needs
//                       (the attention matrix itself is never written).  Keys are split over CTAs (flash-style
//                       partial max / sum / output) and merged by a second tiny kernel.  Also used for the 10x10
//                       token self-attention (with key padding).
//   feat <- tok ("fq"): queries = pixels, keys/values = tokens (with key padding); softmax over <= 16 tokens is
//                       local to each pixel.
// The 128x128 projections around these cores dominate the FLOPs and run on tcgen05 (K9 rows GEMM); the cores
// themselves have N = 10 and are HBM/latency bound, so they are CUDA-core kernels with the few side in shared
// memory and the many side streamed once through a 128-row tile.  Softmax, statistics and gradients in fp32.
#include "common.cuh"

namespace {

constexpr int E = 128;        // attention width (atten_dim of both live configs)
constexpr int TM = 128;       // many-side rows per CTA
constexpr int MAXF = 16;      // few-side rows
constexpr int ROWP = E + 8;   // padded tile row (halfs): 272 B = 17 x 16 B, conflict-free for 16-byte loads by consecutive threads
constexpr int PF = 20;        // pitch (floats) of the key-major probability / score-gradient rows of the backward kernels

__device__ __forceinline__ void load_tile(__half (*dst)[ROWP], const __half* __restrict__ src, int rows_valid) {
    // 128 rows x 128 halfs, coalesced 16-byte pieces
    for (int i = threadIdx.x; i < TM * (E / 8); i += blockDim.x) {
        const int r = i / (E / 8), c = i - r * (E / 8);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows_valid) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * E) + c);
        *reinterpret_cast<uint4*>(&dst[r][c * 8]) = v;
    }
}

// Row . q over E = 128 dims.  16-byte shared-memory loads: the padded row pitch (272 B = 17 x 16 B) makes the 128-bit row loads
// of consecutive threads conflict-free (the 4-byte loads of the first version were 4-way conflicted: pitch 68 words), and the
// broadcast reads of q go four floats at a time.  Two accumulators break the dependent FMA chain.
__device__ __forceinline__ float dot_row(const __half* row, const float* q) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
    for (int e = 0; e < E; e += 8) {
        const uint4 kk = *reinterpret_cast<const uint4*>(row + e);
        const float4 q0 = *reinterpret_cast<const float4*>(q + e), q1 = *reinterpret_cast<const float4*>(q + e + 4);
        const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&kk.x)), k1 = __half22float2(*reinterpret_cast<const __half2*>(&kk.y));
        const float2 k2 = __half22float2(*reinterpret_cast<const __half2*>(&kk.z)), k3 = __half22float2(*reinterpret_cast<const __half2*>(&kk.w));
        a0 = fmaf(k0.x, q0.x, a0), a1 = fmaf(k0.y, q0.y, a1);
        a0 = fmaf(k1.x, q0.z, a0), a1 = fmaf(k1.y, q0.w, a1);
        a0 = fmaf(k2.x, q1.x, a0), a1 = fmaf(k2.y, q1.y, a1);
        a0 = fmaf(k3.x, q1.z, a0), a1 = fmaf(k3.y, q1.w, a1);
    }
    return a0 + a1;
}

// ---- fp32 many side (evaluation at fp32-level accuracy, forward only): rows of E floats, padded to E + 4 -------------
constexpr int ROWF = E + 4;

__device__ __forceinline__ void load_tile(float (*dst)[ROWF], const float* __restrict__ src, int rows_valid) {
    for (int i = threadIdx.x; i < TM * (E / 4); i += blockDim.x) {
        const int r = i / (E / 4), c = i - r * (E / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_valid) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * E) + c);
        *reinterpret_cast<float4*>(&dst[r][c * 4]) = v;
    }
}

__device__ __forceinline__ float dot_row(const float* row, const float* q) {
    float acc = 0.f;
#pragma unroll 8
    for (int e = 0; e < E; e += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(row + e);
        acc = fmaf(k4.x, q[e], acc);
        acc = fmaf(k4.y, q[e + 1], acc);
        acc = fmaf(k4.z, q[e + 2], acc);
        acc = fmaf(k4.w, q[e + 3], acc);
    }
    return acc;
}

template <typename T> struct Tile;
template <> struct Tile<__half> { static constexpr int ROW = ROWP; };
template <> struct Tile<float> { static constexpr int ROW = ROWF; };
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ void from_f(__half& d, float v) { d = __float2half(v); }
__device__ __forceinline__ void from_f(float& d, float v) { d = v; }
// fp32 mode uses the accurate exponential (the fast one carries ~2 ulp + range-reduction error)
template <typename T> __device__ __forceinline__ float exp_t(float x) { return __expf(x); }
template <> __device__ __forceinline__ float exp_t<float>(float x) { return expf(x); }

// ================================================================================================ tq forward
// grid (nsplit, B), 128 threads.  partial outputs: pm/pl [B][nsplit][F], po [B][nsplit][F][E], pstat [B][nsplit][F]
template <typename T>
__global__ void __launch_bounds__(128)
attn_tq_partial_kernel(const float* __restrict__ Q, const T* __restrict__ K, const T* __restrict__ V,
                       const uint8_t* __restrict__ key_pad, const uint8_t* __restrict__ guid, int F, int S, float scale,
                       float* __restrict__ pm, float* __restrict__ pl, float* __restrict__ po, float* __restrict__ pstat) {
    mg::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    constexpr int ROW = Tile<T>::ROW;
    T (*sK)[ROW] = reinterpret_cast<T (*)[ROW]>(smem_raw);
    T (*sV)[ROW] = sK + TM;
    float* sQ = reinterpret_cast<float*>(sV + TM);   // [MAXF][E]
    float* sP = sQ + MAXF * E;                        // [MAXF][TM]
    float* sRed = sP + MAXF * TM;                     // [4]
    const int b = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x, t = threadIdx.x;
    const int s0 = split * TM, rows = min(TM, S - s0);
    load_tile(sK, K + ((size_t)b * S + s0) * E, rows);
    load_tile(sV, V + ((size_t)b * S + s0) * E, rows);
    for (int i = t; i < F * E; i += 128) sQ[i] = Q[(size_t)b * F * E + i];
    __syncthreads();
    const bool live = t < rows && !(key_pad && key_pad[(size_t)b * S + s0 + t]);
    for (int f = 0; f < F; ++f) {
        const float s = live ? dot_row(sK[t], sQ + f * E) * scale : -INFINITY;
        // block max
        float m = s;
#pragma unroll
        for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        if ((t & 31) == 0) sRed[t >> 5] = m;
        __syncthreads();
        m = fmaxf(fmaxf(sRed[0], sRed[1]), fmaxf(sRed[2], sRed[3]));
        __syncthreads();
        const float p = (live && m > -INFINITY) ? exp_t<T>(s - m) : 0.f;
        sP[f * TM + t] = p;
        float l = p, st = (guid && live && guid[((size_t)b * F + f) * S + s0 + t]) ? p : 0.f;
#pragma unroll
        for (int d = 16; d; d >>= 1) l += __shfl_xor_sync(0xffffffffu, l, d), st += __shfl_xor_sync(0xffffffffu, st, d);
        if ((t & 31) == 0) sRed[t >> 5] = l, sRed[4 + (t >> 5)] = st;
        __syncthreads();
        if (t == 0) {
            const size_t o = ((size_t)b * nsplit + split) * F + f;
            pm[o] = m, pl[o] = sRed[0] + sRed[1] + sRed[2] + sRed[3], pstat[o] = sRed[4] + sRed[5] + sRed[6] + sRed[7];
        }
        __syncthreads();
    }
    // partial output: thread = channel e
    for (int f = 0; f < F; ++f) {
        float acc = 0.f;
        for (int k = 0; k < rows; ++k) acc = fmaf(sP[f * TM + k], to_f(sV[k][t]), acc);
        po[(((size_t)b * nsplit + split) * F + f) * E + t] = acc;
    }
}

// grid (F, B), 128 threads: merge the key splits.  The per-split scalars are staged in shared memory with ONE round of
// loads (the first version walked the splits serially, two dependent global-load latencies per split: 25 us for a kernel
// that moves 160 KB), and the partial outputs are read eight at a time.
__global__ void __launch_bounds__(128)
attn_tq_merge_kernel(const float* __restrict__ pm, const float* __restrict__ pl, const float* __restrict__ po,
                     const float* __restrict__ pstat, int F, int nsplit, float* __restrict__ O, float* __restrict__ stat,
                     float* __restrict__ M, float* __restrict__ L, int accurate) {
    mg::pdl_prologue();
    constexpr int MAXS = 256;                        // splits staged per pass (S <= 32768 keys per sample)
    __shared__ float s_m[MAXS], s_w[MAXS];
    __shared__ float s_red[4];
    const int f = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    float m = -INFINITY;
    for (int s = t; s < nsplit; s += 128) m = fmaxf(m, pm[((size_t)b * nsplit + s) * F + f]);
#pragma unroll
    for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((t & 31) == 0) s_red[t >> 5] = m;
    __syncthreads();
    m = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    float l = 0.f, st = 0.f, acc = 0.f;
    for (int s0 = 0; s0 < nsplit; s0 += MAXS) {
        const int ns = min(MAXS, nsplit - s0);
        __syncthreads();
        for (int s = t; s < ns; s += 128) {
            const float v = pm[((size_t)b * nsplit + s0 + s) * F + f];
            s_m[s] = v;
            s_w[s] = v > -INFINITY ? (accurate ? expf(v - m) : __expf(v - m)) : 0.f;
        }
        __syncthreads();
        // l and stat: every thread accumulates the same sums from shared weights (the loads of pl / pstat are independent)
        for (int s = 0; s < ns; ++s) {
            const size_t o = ((size_t)b * nsplit + s0 + s) * F + f;
            l = fmaf(s_w[s], __ldg(pl + o), l), st = fmaf(s_w[s], __ldg(pstat + o), st);
        }
        int s = 0;
        for (; s + 8 <= ns; s += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(po + (((size_t)b * nsplit + s0 + s + j) * F + f) * E + t);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(s_w[s + j], v[j], acc);
        }
        for (; s < ns; ++s) acc = fmaf(s_w[s], __ldg(po + (((size_t)b * nsplit + s0 + s) * F + f) * E + t), acc);
    }
    // a fully masked row (l == 0) yields NaN exactly like softmax over an all -inf row in the reference
    O[((size_t)b * F + f) * E + t] = acc / l;
    if (t == 0) stat[(size_t)b * F + f] = st / l, M[(size_t)b * F + f] = m, L[(size_t)b * F + f] = l;
}

// ================================================================================================ tq backward
// grid (nsplit, B), 128 threads.  dQ accumulated with atomics (fp32, caller zeroes); dK, dV fp16 rows.
__global__ void __launch_bounds__(128)
attn_tq_bwd_kernel(const float* __restrict__ Q, const __half* __restrict__ K, const __half* __restrict__ V,
                   const uint8_t* __restrict__ key_pad, const uint8_t* __restrict__ guid, const float* __restrict__ O,
                   const float* __restrict__ stat, const float* __restrict__ M, const float* __restrict__ L,
                   const float* __restrict__ dO, const float* __restrict__ dstat, int F, int S, float scale,
                   float* __restrict__ dQ, __half* __restrict__ dK, __half* __restrict__ dV) {
    mg::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    __half (*sK)[ROWP] = reinterpret_cast<__half (*)[ROWP]>(smem_raw);
    __half (*sV)[ROWP] = sK + TM;
    float* sQ = reinterpret_cast<float*>(sV + TM);   // [MAXF][E]
    float* sdO = sQ + MAXF * E;                       // [MAXF][E]
    float* sP = sdO + MAXF * E;                       // [TM][PF]
    float* sdS = sP + TM * PF;                        // [TM][PF]
    float* sD = sdS + TM * PF;                        // [MAXF]: sum_k P dP
    const int b = blockIdx.y, split = blockIdx.x, t = threadIdx.x;
    const int s0 = split * TM, rows = min(TM, S - s0);
    load_tile(sK, K + ((size_t)b * S + s0) * E, rows);
    load_tile(sV, V + ((size_t)b * S + s0) * E, rows);
    for (int i = t; i < F * E; i += 128) sQ[i] = Q[(size_t)b * F * E + i], sdO[i] = dO[(size_t)b * F * E + i];
    __syncthreads();
    // D[f] = dO[f] . O[f] (+ the statistic's term): one warp per query, coalesced reads of O (ten threads walking O with 128
    // dependent global loads each held the whole CTA back)
    for (int f = t >> 5; f < F; f += 4) {
        const int lane = t & 31;
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < E / 32; ++j) d = fmaf(sdO[f * E + lane + 32 * j], __ldg(O + ((size_t)b * F + f) * E + lane + 32 * j), d);
#pragma unroll
        for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) sD[f] = d + (dstat ? dstat[(size_t)b * F + f] * stat[(size_t)b * F + f] : 0.f);
    }
    __syncthreads();
    // sP / sdS are KEY-major ([TM][PF], PF = 20 floats: 16-byte aligned rows): the channel phase below reads all queries of
    // a key with broadcast 16-byte loads
    const bool live = t < rows && !(key_pad && key_pad[(size_t)b * S + s0 + t]);
    for (int f = 0; f < F; ++f) {
        float p = 0.f, ds = 0.f;
        if (live) {
            const float s = dot_row(sK[t], sQ + f * E) * scale;
            p = __expf(s - M[(size_t)b * F + f]) / L[(size_t)b * F + f];
            float dp = dot_row(sV[t], sdO + f * E);
            if (dstat && guid && guid[((size_t)b * F + f) * S + s0 + t]) dp += dstat[(size_t)b * F + f];
            ds = p * (dp - sD[f]);
        }
        sP[t * PF + f] = p, sdS[t * PF + f] = ds;
    }
    __syncthreads();
    // thread = channel e: dV[k][e], dK[k][e] rows and dQ[f][e] in ONE pass over the keys; the query-side factors of this
    // channel live in registers
    float rdO[MAXF], rQ[MAXF], aq[MAXF];
#pragma unroll
    for (int f = 0; f < MAXF; ++f) rdO[f] = f < F ? sdO[f * E + t] : 0.f, rQ[f] = f < F ? sQ[f * E + t] : 0.f, aq[f] = 0.f;
    for (int k = 0; k < rows; ++k) {
        float pv[MAXF], dv[MAXF];
#pragma unroll
        for (int j = 0; j < MAXF / 4; ++j) {
            if (4 * j < F) {
                const float4 a4 = *reinterpret_cast<const float4*>(sP + k * PF + 4 * j), b4 = *reinterpret_cast<const float4*>(sdS + k * PF + 4 * j);
                pv[4 * j] = a4.x, pv[4 * j + 1] = a4.y, pv[4 * j + 2] = a4.z, pv[4 * j + 3] = a4.w;
                dv[4 * j] = b4.x, dv[4 * j + 1] = b4.y, dv[4 * j + 2] = b4.z, dv[4 * j + 3] = b4.w;
            }
        }
        const float kv = __half2float(sK[k][t]);
        float av = 0.f, ak = 0.f;
#pragma unroll
        for (int f = 0; f < MAXF; ++f) {
            if (f < F) {
                av = fmaf(pv[f], rdO[f], av);
                ak = fmaf(dv[f], rQ[f], ak);
                aq[f] = fmaf(dv[f], kv, aq[f]);
            }
        }
        dV[((size_t)b * S + s0 + k) * E + t] = __float2half(av);
        dK[((size_t)b * S + s0 + k) * E + t] = __float2half(ak * scale);
    }
#pragma unroll
    for (int f = 0; f < MAXF; ++f)
        if (f < F) atomicAdd(dQ + ((size_t)b * F + f) * E + t, aq[f] * scale);
}

// ================================================================================================ fq forward / backward
// queries = many side.  grid (ceil(S/128), B), 128 threads.
template <typename T>
__global__ void __launch_bounds__(128)
attn_fq_fwd_kernel(const T* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                   const uint8_t* __restrict__ key_pad, int F, int S, float scale, T* __restrict__ O) {
    mg::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    constexpr int ROW = Tile<T>::ROW;
    T (*sQ)[ROW] = reinterpret_cast<T (*)[ROW]>(smem_raw);
    float* sK = reinterpret_cast<float*>(sQ + TM);   // [MAXF][E]
    float* sV = sK + MAXF * E;
    float* sP = sV + MAXF * E;                        // [TM][MAXF+1]
    const int b = blockIdx.y, t = threadIdx.x, s0 = blockIdx.x * TM, rows = min(TM, S - s0);
    load_tile(sQ, Q + ((size_t)b * S + s0) * E, rows);
    for (int i = t; i < F * E; i += 128) sK[i] = K[(size_t)b * F * E + i], sV[i] = V[(size_t)b * F * E + i];
    __syncthreads();
    {
        float s[MAXF], m = -INFINITY;
#pragma unroll
        for (int f = 0; f < MAXF; ++f) {
            s[f] = -INFINITY;
            if (f < F && !(key_pad && key_pad[(size_t)b * F + f])) s[f] = dot_row(sQ[t], sK + f * E) * scale;
            m = fmaxf(m, s[f]);
        }
        float l = 0.f;
#pragma unroll
        for (int f = 0; f < MAXF; ++f) s[f] = (s[f] > -INFINITY) ? exp_t<T>(s[f] - m) : 0.f, l += s[f];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) sP[t * (MAXF + 1) + f] = s[f] / l;   // l == 0 (all keys padded) -> NaN as in the reference
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
        float acc = 0.f;
        for (int f = 0; f < F; ++f) acc = fmaf(sP[r * (MAXF + 1) + f], sV[f * E + t], acc);
        from_f(O[((size_t)b * S + s0 + r) * E + t], acc);
    }
}

// dQ fp16 rows; dK, dV [B][F][E] fp32 accumulated with atomics (caller zeroes).
__global__ void __launch_bounds__(128)
attn_fq_bwd_kernel(const __half* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                   const uint8_t* __restrict__ key_pad, const __half* __restrict__ dO, int F, int S, float scale,
                   __half* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dV) {
    mg::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    __half (*sQ)[ROWP] = reinterpret_cast<__half (*)[ROWP]>(smem_raw);
    __half (*sdO)[ROWP] = sQ + TM;
    float* sK = reinterpret_cast<float*>(sdO + TM);
    float* sV = sK + MAXF * E;
    float* sP = sV + MAXF * E;                        // [TM][MAXF+1]
    float* sdS = sP + TM * (MAXF + 1);                // [TM][MAXF+1]
    const int b = blockIdx.y, t = threadIdx.x, s0 = blockIdx.x * TM, rows = min(TM, S - s0);
    load_tile(sQ, Q + ((size_t)b * S + s0) * E, rows);
    load_tile(sdO, dO + ((size_t)b * S + s0) * E, rows);
    for (int i = t; i < F * E; i += 128) sK[i] = K[(size_t)b * F * E + i], sV[i] = V[(size_t)b * F * E + i];
    __syncthreads();
    {
        float s[MAXF], dp[MAXF], m = -INFINITY;
#pragma unroll
        for (int f = 0; f < MAXF; ++f) {
            s[f] = -INFINITY, dp[f] = 0.f;
            if (f < F && !(key_pad && key_pad[(size_t)b * F + f])) {
                s[f] = dot_row(sQ[t], sK + f * E) * scale;
                dp[f] = dot_row(sdO[t], sV + f * E);
            }
            m = fmaxf(m, s[f]);
        }
        float l = 0.f, d = 0.f;
#pragma unroll
        for (int f = 0; f < MAXF; ++f) s[f] = (s[f] > -INFINITY) ? __expf(s[f] - m) : 0.f, l += s[f];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) s[f] = (t < rows) ? s[f] / l : 0.f, d += s[f] * dp[f];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) sP[t * (MAXF + 1) + f] = s[f], sdS[t * (MAXF + 1) + f] = s[f] * (dp[f] - d);
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
        float acc = 0.f;
        for (int f = 0; f < F; ++f) acc = fmaf(sdS[r * (MAXF + 1) + f], sK[f * E + t], acc);
        dQ[((size_t)b * S + s0 + r) * E + t] = __float2half(acc * scale);
    }
    for (int f = 0; f < F; ++f) {
        float ak = 0.f, av = 0.f;
        for (int r = 0; r < rows; ++r) {
            ak = fmaf(sdS[r * (MAXF + 1) + f], __half2float(sQ[r][t]), ak);
            av = fmaf(sP[r * (MAXF + 1) + f], __half2float(sdO[r][t]), av);
        }
        atomicAdd(dK + ((size_t)b * F + f) * E + t, ak * scale);
        atomicAdd(dV + ((size_t)b * F + f) * E + t, av);
    }
}

size_t tq_smem() { return 2 * TM * ROWP * 2 + (MAXF * E + MAXF * TM + 8) * 4; }
size_t tq_smem_f32() { return 2 * TM * ROWF * 4 + (MAXF * E + MAXF * TM + 8) * 4; }
size_t fq_smem_f32() { return TM * ROWF * 4 + (2 * MAXF * E + TM * (MAXF + 1)) * 4; }
size_t tq_bwd_smem() { return 2 * TM * ROWP * 2 + (2 * MAXF * E + 2 * TM * PF + MAXF) * 4; }
size_t fq_smem() { return TM * ROWP * 2 + (2 * MAXF * E + TM * (MAXF + 1)) * 4; }
size_t fq_bwd_smem() { return 2 * TM * ROWP * 2 + (2 * MAXF * E + 2 * TM * (MAXF + 1)) * 4; }

template <typename Kern>
int raise_smem(Kern k, size_t bytes, const char* who, bool& done) {
    if (done) return MG_OK;
    if (bytes > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        mg::set_error("%s: cannot raise dynamic shared memory limit", who);
        return MG_ERR_CUDA;
    }
    done = true;
    return MG_OK;
}

}  // namespace

extern "C" size_t mg_attn_tq_workspace_floats(int B, int F, int S) {
    const size_t ns = (size_t)mg::ceil_div(S, TM);
    return (size_t)B * ns * F * (3 + E);
}

static int attn_tq_fwd(const float* q, const void* k, const void* v, bool kv_f32, const uint8_t* key_pad, const uint8_t* guidance,
                       int B, int F, int S, int Edim, float* out, float* stat, float* row_max, float* row_sum,
                       float* ws, void* stream) {
    MG_REQUIRE(q && k && v && out && stat && row_max && row_sum && ws, "mg_attn_tq_fwd: null pointer");
    MG_REQUIRE(Edim == E && F >= 1 && F <= MAXF && S >= 1 && B >= 1 && B <= 65535, "mg_attn_tq_fwd: unsupported shape E=%d F=%d S=%d", Edim, F, S);
    const int ns = mg::ceil_div(S, TM);
    float* pm = ws;
    float* pl = pm + (size_t)B * ns * F;
    float* pstat = pl + (size_t)B * ns * F;
    float* po = pstat + (size_t)B * ns * F;
    static bool done = false, done32 = false;
    if (int e = raise_smem(attn_tq_partial_kernel<__half>, tq_smem(), "mg_attn_tq_fwd", done)) return e;
    if (int e = raise_smem(attn_tq_partial_kernel<float>, tq_smem_f32(), "mg_attn_tq_fwd", done32)) return e;
    const float scale = 1.0f / sqrtf((float)E);
    if (kv_f32)
        MG_LAUNCH(attn_tq_partial_kernel<float>, dim3(ns, B), 128, tq_smem_f32(), stream, q, static_cast<const float*>(k),
                  static_cast<const float*>(v), key_pad, guidance, F, S, scale, pm, pl, po, pstat);
    else
        MG_LAUNCH(attn_tq_partial_kernel<__half>, dim3(ns, B), 128, tq_smem(), stream, q, static_cast<const __half*>(k),
                  static_cast<const __half*>(v), key_pad, guidance, F, S, scale, pm, pl, po, pstat);
    MG_LAUNCH(attn_tq_merge_kernel, dim3(F, B), 128, 0, stream, pm, pl, po, pstat, F, ns, out, stat, row_max, row_sum,
              kv_f32 ? 1 : 0);
    MG_CHECK_LAUNCH("mg_attn_tq_fwd");
    return MG_OK;
}

extern "C" int mg_attn_tq_fwd(const float* q, const void* k, const void* v, const uint8_t* key_pad, const uint8_t* guidance,
                              int B, int F, int S, int Edim, float* out, float* stat, float* row_max, float* row_sum,
                              float* ws, void* stream) {
    return attn_tq_fwd(q, k, v, false, key_pad, guidance, B, F, S, Edim, out, stat, row_max, row_sum, ws, stream);
}

extern "C" int mg_attn_tq_fwd_f32(const float* q, const float* k, const float* v, const uint8_t* key_pad,
                                  const uint8_t* guidance, int B, int F, int S, int Edim, float* out, float* stat,
                                  float* row_max, float* row_sum, float* ws, void* stream) {
    return attn_tq_fwd(q, k, v, true, key_pad, guidance, B, F, S, Edim, out, stat, row_max, row_sum, ws, stream);
}

extern "C" int mg_attn_tq_bwd(const float* q, const void* k, const void* v, const uint8_t* key_pad, const uint8_t* guidance,
                              const float* out, const float* stat, const float* row_max, const float* row_sum,
                              const float* d_out, const float* d_stat, int B, int F, int S, int Edim, float* dq, void* dk,
                              void* dv, void* stream) {
    MG_REQUIRE(q && k && v && out && stat && row_max && row_sum && d_out && dq && dk && dv, "mg_attn_tq_bwd: null pointer");
    MG_REQUIRE(Edim == E && F >= 1 && F <= MAXF && S >= 1 && B >= 1 && B <= 65535, "mg_attn_tq_bwd: unsupported shape");
    static bool done = false;
    if (int e = raise_smem(attn_tq_bwd_kernel, tq_bwd_smem(), "mg_attn_tq_bwd", done)) return e;
    const float scale = 1.0f / sqrtf((float)E);
    MG_LAUNCH(attn_tq_bwd_kernel, dim3(mg::ceil_div(S, TM), B), 128, tq_bwd_smem(), stream, q, static_cast<const __half*>(k),
              static_cast<const __half*>(v), key_pad, guidance, out, stat, row_max, row_sum, d_out, d_stat, F, S, scale, dq,
              static_cast<__half*>(dk), static_cast<__half*>(dv));
    MG_CHECK_LAUNCH("mg_attn_tq_bwd");
    return MG_OK;
}

extern "C" int mg_attn_fq_fwd(const void* q, const float* k, const float* v, const uint8_t* key_pad, int B, int F, int S,
                              int Edim, void* out, void* stream) {
    MG_REQUIRE(q && k && v && out, "mg_attn_fq_fwd: null pointer");
    MG_REQUIRE(Edim == E && F >= 1 && F <= MAXF && S >= 1 && B >= 1 && B <= 65535, "mg_attn_fq_fwd: unsupported shape");
    static bool done = false;
    if (int e = raise_smem(attn_fq_fwd_kernel<__half>, fq_smem(), "mg_attn_fq_fwd", done)) return e;
    MG_LAUNCH(attn_fq_fwd_kernel<__half>, dim3(mg::ceil_div(S, TM), B), 128, fq_smem(), stream, static_cast<const __half*>(q), k, v,
              key_pad, F, S, 1.0f / sqrtf((float)E), static_cast<__half*>(out));
    MG_CHECK_LAUNCH("mg_attn_fq_fwd");
    return MG_OK;
}

extern "C" int mg_attn_fq_fwd_f32(const float* q, const float* k, const float* v, const uint8_t* key_pad, int B, int F, int S,
                                  int Edim, float* out, void* stream) {
    MG_REQUIRE(q && k && v && out, "mg_attn_fq_fwd_f32: null pointer");
    MG_REQUIRE(Edim == E && F >= 1 && F <= MAXF && S >= 1 && B >= 1 && B <= 65535, "mg_attn_fq_fwd_f32: unsupported shape");
    static bool done = false;
    if (int e = raise_smem(attn_fq_fwd_kernel<float>, fq_smem_f32(), "mg_attn_fq_fwd_f32", done)) return e;
    MG_LAUNCH(attn_fq_fwd_kernel<float>, dim3(mg::ceil_div(S, TM), B), 128, fq_smem_f32(), stream, q, k, v, key_pad, F, S,
              1.0f / sqrtf((float)E), out);
    MG_CHECK_LAUNCH("mg_attn_fq_fwd_f32");
    return MG_OK;
}

extern "C" int mg_attn_fq_bwd(const void* q, const float* k, const float* v, const uint8_t* key_pad, const void* d_out, int B,
                              int F, int S, int Edim, void* dq, float* dk, float* dv, void* stream) {
    MG_REQUIRE(q && k && v && d_out && dq && dk && dv, "mg_attn_fq_bwd: null pointer");
    MG_REQUIRE(Edim == E && F >= 1 && F <= MAXF && S >= 1 && B >= 1 && B <= 65535, "mg_attn_fq_bwd: unsupported shape");
    static bool done = false;
    if (int e = raise_smem(attn_fq_bwd_kernel, fq_bwd_smem(), "mg_attn_fq_bwd", done)) return e;
    MG_LAUNCH(attn_fq_bwd_kernel, dim3(mg::ceil_div(S, TM), B), 128, fq_bwd_smem(), stream, static_cast<const __half*>(q), k, v,
              key_pad, static_cast<const __half*>(d_out), F, S, 1.0f / sqrtf((float)E), static_cast<__half*>(dq), dk, dv);
    MG_CHECK_LAUNCH("mg_attn_fq_bwd");
    return MG_OK;
}
