// K13: row-wise helpers of the mask-guided attention block and of the sparse stage (HBM-bound streaming kernels).
//   mg_layer_norm_fwd / _bwd : y = LayerNorm(a + b) * gamma + beta over rows of E in {64, 128} fp16 elements, fp32
//                              statistics; the residual add of the post-norm attention layers (`tgt + o`) is fused in.
//                              One warp per row, 8-byte (E = 128) / 4-byte (E = 64) accesses, rows in a grid-stride loop.
//   mg_col_sum               : out[c] += sum_r x[r][c] for fp16 rows (bias gradients of the rulebook convolutions;
//                              replaces a fp32 copy of the rows plus a torch reduction).
#include "common.cuh"

namespace {

template <int PER>  // elements per lane: 4 (E = 128) or 2 (E = 64)
struct Vec;
template <>
struct Vec<4> {
    static __device__ __forceinline__ void load(const __half* p, float (&f)[4]) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y;
    }
    static __device__ __forceinline__ void store(__half* p, const float (&f)[4]) {
        uint2 u;
        __half2 h = __floats2half2_rn(f[0], f[1]);
        u.x = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2half2_rn(f[2], f[3]);
        u.y = *reinterpret_cast<uint32_t*>(&h);
        *reinterpret_cast<uint2*>(p) = u;
    }
};
template <>
struct Vec<2> {
    static __device__ __forceinline__ void load(const __half* p, float (&f)[2]) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u));
        f[0] = a.x, f[1] = a.y;
    }
    static __device__ __forceinline__ void store(__half* p, const float (&f)[2]) {
        const __half2 h = __floats2half2_rn(f[0], f[1]);
        *reinterpret_cast<__half2*>(p) = h;
    }
};

// fp32 rows (evaluation at fp32-level accuracy, forward only)
template <int PER>
struct VecF {
    static __device__ __forceinline__ void load(const float* p, float (&f)[PER]) {
#pragma unroll
        for (int i = 0; i < PER; ++i) f[i] = __ldg(p + i);
    }
    static __device__ __forceinline__ void store(float* p, const float (&f)[PER]) {
#pragma unroll
        for (int i = 0; i < PER; ++i) p[i] = f[i];
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// y = LN(a + b); optionally stores s = a + b (fp16, the tensor the backward normalises again) and (mean, rstd).
template <int PER>
__global__ void __launch_bounds__(256)
layer_norm_fwd_kernel(const __half* __restrict__ a, const __half* __restrict__ b, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, __half* __restrict__ s_out, __half* __restrict__ y,
                      float2* __restrict__ stat, int rows) {
    mg::pdl_prologue();
    constexpr int E = PER * 32;
    const int lane = threadIdx.x & 31, c0 = lane * PER;
    float g[PER], bt[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) g[i] = gamma[c0 + i], bt[i] = beta[c0 + i];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        float v[PER];
        Vec<PER>::load(a + (size_t)r * E + c0, v);
        if (b) {
            float w[PER];
            Vec<PER>::load(b + (size_t)r * E + c0, w);
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] += w[i];
            // the backward sees exactly the fp16 sum a torch `tgt + o` would have produced
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] = __half2float(__float2half_rn(v[i]));
            if (s_out) Vec<PER>::store(s_out + (size_t)r * E + c0, v);
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) sum += v[i];
        const float mean = warp_sum(sum) * (1.f / E);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) sq += (v[i] - mean) * (v[i] - mean);
        const float rstd = rsqrtf(warp_sum(sq) * (1.f / E) + eps);
        float o[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) o[i] = (v[i] - mean) * rstd * g[i] + bt[i];
        Vec<PER>::store(y + (size_t)r * E + c0, o);
        if (stat && lane == 0) stat[r] = make_float2(mean, rstd);
    }
}

// y = LN(a + b) on fp32 rows (no rounding of the sum, sqrt + divide instead of rsqrt)
template <int PER>
__global__ void __launch_bounds__(256)
layer_norm_fwd_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, float* __restrict__ y, int rows) {
    mg::pdl_prologue();
    constexpr int E = PER * 32;
    const int lane = threadIdx.x & 31, c0 = lane * PER;
    float g[PER], bt[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) g[i] = gamma[c0 + i], bt[i] = beta[c0 + i];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        float v[PER];
        VecF<PER>::load(a + (size_t)r * E + c0, v);
        if (b) {
            float w[PER];
            VecF<PER>::load(b + (size_t)r * E + c0, w);
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] += w[i];
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) sum += v[i];
        const float mean = warp_sum(sum) * (1.f / E);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) sq += (v[i] - mean) * (v[i] - mean);
        const float rstd = 1.f / sqrtf(warp_sum(sq) * (1.f / E) + eps);
        float o[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) o[i] = (v[i] - mean) * rstd * g[i] + bt[i];
        VecF<PER>::store(y + (size_t)r * E + c0, o);
    }
}

// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat));  dgamma += sum g*xhat;  dbeta += sum g
template <int PER>
__global__ void __launch_bounds__(256)
layer_norm_bwd_kernel(const __half* __restrict__ s, const __half* __restrict__ gy, const float* __restrict__ gamma,
                      const float2* __restrict__ stat, __half* __restrict__ dx, float* __restrict__ dgb, int rows) {
    mg::pdl_prologue();
    constexpr int E = PER * 32;
    __shared__ float s_acc[8][2 * E];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = lane * PER;
    float g[PER], dg[PER], db[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) g[i] = gamma[c0 + i], dg[i] = 0.f, db[i] = 0.f;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        float v[PER], go[PER];
        Vec<PER>::load(s + (size_t)r * E + c0, v);
        Vec<PER>::load(gy + (size_t)r * E + c0, go);
        const float2 st = stat[r];
        float m1 = 0.f, m2 = 0.f, xh[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            xh[i] = (v[i] - st.x) * st.y;
            dg[i] += go[i] * xh[i], db[i] += go[i];
            const float t = go[i] * g[i];
            m1 += t, m2 += t * xh[i];
        }
        m1 = warp_sum(m1) * (1.f / E), m2 = warp_sum(m2) * (1.f / E);
        float o[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) o[i] = st.y * (go[i] * g[i] - m1 - xh[i] * m2);
        Vec<PER>::store(dx + (size_t)r * E + c0, o);
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) s_acc[warp][c0 + i] = dg[i], s_acc[warp][E + c0 + i] = db[i];
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * E; t += 256) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_acc[w][t];
        if (acc != 0.f) atomicAdd(dgb + t, acc);
    }
}

// out[c] += sum over rows of x[r][c]; 8 channels per thread, C/8 threads per row group.
__global__ void __launch_bounds__(256)
col_sum_kernel(const __half* __restrict__ x, int stride, int rows, int C, float* __restrict__ out) {
    mg::pdl_prologue();
    __shared__ float s_red[256 * 8];
    const int G = C >> 3, per = 256 / G;
    const int g = threadIdx.x % G, sub = threadIdx.x / G;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (sub < per) {
        for (int r = blockIdx.x * per + sub; r < rows; r += gridDim.x * per) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (size_t)r * stride + (g << 3)));
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 t = __half22float2(h[i]);
                acc[2 * i] += t.x, acc[2 * i + 1] += t.y;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[threadIdx.x * 8 + i] = sub < per ? acc[i] : 0.f;
    __syncthreads();
    for (int t = threadIdx.x; t < C; t += 256) {
        const int gg = t >> 3, k = t & 7;
        float a = 0.f;
        for (int s2 = 0; s2 < per; ++s2) a += s_red[(s2 * G + gg) * 8 + k];
        if (a != 0.f) atomicAdd(out + t, a);
    }
}

int row_grid(int rows) { return std::max(1, std::min((rows + 7) / 8, mg::kNumSMs * 8)); }

}  // namespace

extern "C" int mg_layer_norm_fwd(const void* a, const void* b, const float* gamma, const float* beta, float eps, void* sum_out,
                                 void* y, float* stat, int rows, int E, void* stream) {
    MG_REQUIRE(E == 64 || E == 128, "mg_layer_norm_fwd: E must be 64 or 128 (got %d)", E);
    if (rows <= 0) return MG_OK;
    MG_REQUIRE(a && gamma && beta && y, "mg_layer_norm_fwd: null pointer");
    const __half *pa = static_cast<const __half*>(a), *pb = static_cast<const __half*>(b);
    __half *ps = static_cast<__half*>(sum_out), *py = static_cast<__half*>(y);
    float2* st = reinterpret_cast<float2*>(stat);
    if (E == 128)
        MG_LAUNCH(layer_norm_fwd_kernel<4>, row_grid(rows), 256, 0, stream, pa, pb, gamma, beta, eps, ps, py, st, rows);
    else
        MG_LAUNCH(layer_norm_fwd_kernel<2>, row_grid(rows), 256, 0, stream, pa, pb, gamma, beta, eps, ps, py, st, rows);
    MG_CHECK_LAUNCH("mg_layer_norm_fwd");
    return MG_OK;
}

extern "C" int mg_layer_norm_fwd_f32(const float* a, const float* b, const float* gamma, const float* beta, float eps, float* y,
                                     int rows, int E, void* stream) {
    MG_REQUIRE(E == 64 || E == 128, "mg_layer_norm_fwd_f32: E must be 64 or 128 (got %d)", E);
    if (rows <= 0) return MG_OK;
    MG_REQUIRE(a && gamma && beta && y, "mg_layer_norm_fwd_f32: null pointer");
    if (E == 128)
        MG_LAUNCH(layer_norm_fwd_f32_kernel<4>, row_grid(rows), 256, 0, stream, a, b, gamma, beta, eps, y, rows);
    else
        MG_LAUNCH(layer_norm_fwd_f32_kernel<2>, row_grid(rows), 256, 0, stream, a, b, gamma, beta, eps, y, rows);
    MG_CHECK_LAUNCH("mg_layer_norm_fwd_f32");
    return MG_OK;
}

extern "C" int mg_layer_norm_bwd(const void* s, const void* gy, const float* gamma, const float* stat, void* dx, float* dgb,
                                 int rows, int E, void* stream) {
    MG_REQUIRE(E == 64 || E == 128, "mg_layer_norm_bwd: E must be 64 or 128 (got %d)", E);
    if (rows <= 0) return MG_OK;
    MG_REQUIRE(s && gy && gamma && stat && dx && dgb, "mg_layer_norm_bwd: null pointer");
    const __half *ps = static_cast<const __half*>(s), *pg = static_cast<const __half*>(gy);
    const float2* st = reinterpret_cast<const float2*>(stat);
    const int grid = std::max(1, std::min((rows + 7) / 8, mg::kNumSMs * 2));  // few blocks: one atomic flush per block
    if (E == 128)
        MG_LAUNCH(layer_norm_bwd_kernel<4>, grid, 256, 0, stream, ps, pg, gamma, st, static_cast<__half*>(dx), dgb, rows);
    else
        MG_LAUNCH(layer_norm_bwd_kernel<2>, grid, 256, 0, stream, ps, pg, gamma, st, static_cast<__half*>(dx), dgb, rows);
    MG_CHECK_LAUNCH("mg_layer_norm_bwd");
    return MG_OK;
}

extern "C" int mg_col_sum(const void* x, int stride, int rows, int C, float* out, void* stream) {
    MG_REQUIRE(C % 8 == 0 && C >= 8 && C <= 2048 && stride % 8 == 0, "mg_col_sum: C and stride must be multiples of 8 (C=%d)", C);
    MG_REQUIRE(256 % (C / 8) == 0, "mg_col_sum: C/8 must divide 256 (C=%d)", C);
    if (rows <= 0) return MG_OK;
    MG_REQUIRE(x && out, "mg_col_sum: null pointer");
    const int per = 256 / (C / 8);
    const int grid = std::max(1, std::min((rows + per - 1) / per, mg::kNumSMs * 4));
    MG_LAUNCH(col_sum_kernel, grid, 256, 0, stream, static_cast<const __half*>(x), stride, rows, C, out);
    MG_CHECK_LAUNCH("mg_col_sum");
    return MG_OK;
}

// ---- token logits: logits[bt][q][p] = sum_c tok[bt / n_f][q][c] * x[bt][p][c]  (the einsum of the OS8 head) -----------
// x is the NHWC fp16 feature map (rows of C = 64 channels), tok the 10 instance tokens per sample in fp32.  Replaces an
// fp32 copy of x plus three cuBLAS batched GEMMs with K = 10 / 64 / 4096 shapes it handles badly (147 us for the token
// gradient alone).
namespace {

constexpr int TL_C = 64, TL_QMAX = 16;

__global__ void __launch_bounds__(256)
token_logits_fwd_kernel(const float* __restrict__ tok, const __half* __restrict__ x, float* __restrict__ logits, int BT, int n_f,
                        int Q, int HW) {
    mg::pdl_prologue();
    __shared__ float s_tok[TL_QMAX][TL_C];
    const int bt = blockIdx.y, b = bt / n_f;
    for (int i = threadIdx.x; i < Q * TL_C; i += 256) s_tok[i / TL_C][i % TL_C] = tok[(size_t)b * Q * TL_C + i];
    __syncthreads();
    for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + ((size_t)bt * HW + p) * TL_C);
        float xv[TL_C];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 u = __ldg(xr + j);
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(h[i]);
                xv[j * 8 + 2 * i] = f.x, xv[j * 8 + 2 * i + 1] = f.y;
            }
        }
        for (int q = 0; q < Q; ++q) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < TL_C; ++c) acc += s_tok[q][c] * xv[c];
            logits[((size_t)bt * Q + q) * HW + p] = acc;
        }
    }
}

// the same with fp32 NHWC features (evaluation at fp32-level accuracy)
__global__ void __launch_bounds__(256)
token_logits_fwd_f32_kernel(const float* __restrict__ tok, const float* __restrict__ x, float* __restrict__ logits, int BT,
                            int n_f, int Q, int HW) {
    mg::pdl_prologue();
    __shared__ float s_tok[TL_QMAX][TL_C];
    const int bt = blockIdx.y, b = bt / n_f;
    for (int i = threadIdx.x; i < Q * TL_C; i += 256) s_tok[i / TL_C][i % TL_C] = tok[(size_t)b * Q * TL_C + i];
    __syncthreads();
    for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
        const float4* xr = reinterpret_cast<const float4*>(x + ((size_t)bt * HW + p) * TL_C);
        float xv[TL_C];
#pragma unroll
        for (int j = 0; j < TL_C / 4; ++j) {
            const float4 u = __ldg(xr + j);
            xv[4 * j] = u.x, xv[4 * j + 1] = u.y, xv[4 * j + 2] = u.z, xv[4 * j + 3] = u.w;
        }
        for (int q = 0; q < Q; ++q) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < TL_C; ++c) acc = fmaf(s_tok[q][c], xv[c], acc);
            logits[((size_t)bt * Q + q) * HW + p] = acc;
        }
    }
}

// dx[bt][p][c] = sum_q g[bt][q][p] * tok[b][q][c]   (fp16 NHWC)
__global__ void __launch_bounds__(256)
token_logits_bwd_x_kernel(const float* __restrict__ tok, const float* __restrict__ g, __half* __restrict__ dx, int BT, int n_f,
                          int Q, int HW) {
    mg::pdl_prologue();
    __shared__ float s_tok[TL_QMAX][TL_C];
    const int bt = blockIdx.y, b = bt / n_f;
    for (int i = threadIdx.x; i < Q * TL_C; i += 256) s_tok[i / TL_C][i % TL_C] = tok[(size_t)b * Q * TL_C + i];
    __syncthreads();
    for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
        float gv[TL_QMAX];
        for (int q = 0; q < Q; ++q) gv[q] = __ldg(g + ((size_t)bt * Q + q) * HW + p);
        uint4* orow = reinterpret_cast<uint4*>(dx + ((size_t)bt * HW + p) * TL_C);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int q = 0; q < Q; ++q) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += gv[q] * s_tok[q][j * 8 + i];
            }
            uint4 u;
            __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
            orow[j] = u;
        }
    }
}

// dtok[b][q][c] += sum over the block's 128 pixels  g[bt][q][p] * x[bt][p][c]
constexpr int TL_TILE = 128;
__global__ void __launch_bounds__(256)
token_logits_bwd_tok_kernel(const __half* __restrict__ x, const float* __restrict__ g, float* __restrict__ dtok, int BT, int n_f,
                            int Q, int HW) {
    mg::pdl_prologue();
    __shared__ __half s_x[TL_TILE][TL_C + 2];
    __shared__ float s_g[TL_QMAX][TL_TILE];
    const int bt = blockIdx.y, b = bt / n_f, p0 = blockIdx.x * TL_TILE;
    for (int i = threadIdx.x; i < TL_TILE * TL_C / 2; i += 256) {
        const int r = i / (TL_C / 2), c2 = i - r * (TL_C / 2);
        const __half2 v = p0 + r < HW ? *reinterpret_cast<const __half2*>(x + ((size_t)bt * HW + p0 + r) * TL_C + 2 * c2) : __floats2half2_rn(0.f, 0.f);
        *reinterpret_cast<__half2*>(&s_x[r][2 * c2]) = v;
    }
    for (int i = threadIdx.x; i < Q * TL_TILE; i += 256) {
        const int q = i / TL_TILE, r = i - q * TL_TILE;
        s_g[q][r] = p0 + r < HW ? __ldg(g + ((size_t)bt * Q + q) * HW + p0 + r) : 0.f;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < Q * TL_C; o += 256) {
        const int q = o / TL_C, c = o - q * TL_C;
        float acc = 0.f;
#pragma unroll 8
        for (int r = 0; r < TL_TILE; ++r) acc += s_g[q][r] * __half2float(s_x[r][c]);
        atomicAdd(dtok + ((size_t)b * Q + q) * TL_C + c, acc);
    }
}

}  // namespace

extern "C" int mg_token_logits_fwd(const float* tok, const void* x, float* logits, int BT, int n_f, int Q, int HW, int C,
                                   void* stream) {
    MG_REQUIRE(C == TL_C && Q >= 1 && Q <= TL_QMAX && n_f >= 1, "mg_token_logits_fwd: C must be 64 and Q <= 16 (C=%d Q=%d)", C, Q);
    if (BT <= 0 || HW <= 0) return MG_OK;
    MG_REQUIRE(tok && x && logits, "mg_token_logits_fwd: null pointer");
    MG_LAUNCH(token_logits_fwd_kernel, dim3(std::min(mg::ceil_div(HW, 256), 64), BT), 256, 0, stream, tok,
              static_cast<const __half*>(x), logits, BT, n_f, Q, HW);
    MG_CHECK_LAUNCH("mg_token_logits_fwd");
    return MG_OK;
}

extern "C" int mg_token_logits_fwd_f32(const float* tok, const float* x, float* logits, int BT, int n_f, int Q, int HW, int C,
                                       void* stream) {
    MG_REQUIRE(C == TL_C && Q >= 1 && Q <= TL_QMAX && n_f >= 1, "mg_token_logits_fwd_f32: C must be 64 and Q <= 16 (C=%d Q=%d)", C, Q);
    if (BT <= 0 || HW <= 0) return MG_OK;
    MG_REQUIRE(tok && x && logits, "mg_token_logits_fwd_f32: null pointer");
    MG_LAUNCH(token_logits_fwd_f32_kernel, dim3(std::min(mg::ceil_div(HW, 256), 64), BT), 256, 0, stream, tok, x, logits, BT,
              n_f, Q, HW);
    MG_CHECK_LAUNCH("mg_token_logits_fwd_f32");
    return MG_OK;
}

extern "C" int mg_token_logits_bwd(const float* tok, const void* x, const float* g, void* dx, float* dtok, int BT, int n_f, int Q,
                                   int HW, int C, void* stream) {
    MG_REQUIRE(C == TL_C && Q >= 1 && Q <= TL_QMAX && n_f >= 1, "mg_token_logits_bwd: C must be 64 and Q <= 16 (C=%d Q=%d)", C, Q);
    if (BT <= 0 || HW <= 0) return MG_OK;
    MG_REQUIRE(tok && x && g && (dx || dtok), "mg_token_logits_bwd: null pointer");
    if (dx)
        MG_LAUNCH(token_logits_bwd_x_kernel, dim3(std::min(mg::ceil_div(HW, 256), 64), BT), 256, 0, stream, tok, g,
                  static_cast<__half*>(dx), BT, n_f, Q, HW);
    if (dtok)   // caller zeroes dtok [B][Q][C]
        MG_LAUNCH(token_logits_bwd_tok_kernel, dim3(mg::ceil_div(HW, TL_TILE), BT), 256, 0, stream, static_cast<const __half*>(x), g, dtok,
                  BT, n_f, Q, HW);
    MG_CHECK_LAUNCH("mg_token_logits_bwd");
    return MG_OK;
}
