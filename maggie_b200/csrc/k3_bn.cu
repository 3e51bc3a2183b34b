// K3: BatchNorm around the conv kernel, NHWC fp16 activations, fp32 statistics.  All HBM-bound streaming
// kernels (one 16-byte vector = 8 channels per thread access).
//   forward (train): conv epilogue -> per-channel sum/sumsq copies -> mg_bn_train_apply (finalize + apply in one launch;
//                    mg_bn_finalize -> mg_bn_apply when the statistics are exchanged across ranks first)
//   forward (eval) : mg_bn_finalize (running stats) -> scale/shift folded into the conv epilogue (no extra pass)
//   backward       : mg_bn_bwd_reduce (sum dz, sum dz*xhat) -> mg_bn_bwd_apply (dx, optional d-residual)
#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>

namespace {

constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;

// block = 8 copy-groups x 32 channels: the statistic copies are summed by 8 threads per channel (coalesced over the
// channels) and combined through shared memory - the kernel is pure latency, so the serial chain is kept short.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ stats, float count, const float* __restrict__ count_dev,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ rmean,
                   float* __restrict__ rvar, float momentum, float eps, float* __restrict__ scale, float* __restrict__ shift,
                   float* __restrict__ save_mean, float* __restrict__ save_invstd, int C) {
    mg::pdl_prologue();
    // statistics exchanged across ranks (K15): ONE copy of global sums + the global element count on the device
    const int n_copies = count_dev ? 1 : STAT_COPIES;
    if (count_dev) count = fmaxf(__ldg(count_dev), 1.f);
    __shared__ float s_s[8][32], s_q[8][32];
    const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    if (stats) {
        float s = 0.f, q = 0.f;
        if (c < C) {
#pragma unroll
            for (int k = grp; k < n_copies; k += 8) {
                s += __ldg(stats + (size_t)k * 2 * C + c);
                q += __ldg(stats + (size_t)k * 2 * C + C + c);
            }
        }
        s_s[grp][cl] = s, s_q[grp][cl] = q;
        __syncthreads();
    }
    if (grp != 0 || c >= C) return;
    float mean, var;
    if (stats) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_s[k][cl], q += s_q[k][cl];
        mean = s / count;
        var = fmaxf(q / count - mean * mean, 0.f);
        if (rmean) {
            rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
            rvar[c] = (1.f - momentum) * rvar[c] + momentum * var * (count > 1.f ? count / (count - 1.f) : 1.f);
        }
    } else {
        mean = rmean[c], var = rvar[c];
    }
    const float invstd = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale[c] = g * invstd;
    shift[c] = b - mean * g * invstd;
    if (save_mean) save_mean[c] = mean, save_invstd[c] = invstd;
}

__device__ __forceinline__ float act_fwd(float v, int act) {
    return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v > 0.f ? v : 0.2f * v) : v);
}
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
    return act == 1 ? (y > 0.f ? 1.f : 0.f) : (act == 2 ? (y > 0.f ? 1.f : 0.2f) : 1.f);
}

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

// y = act(x*scale + shift (+ res)), vectorised over 8 channels; s_ss = [scale | shift] in shared memory.
__device__ __forceinline__ void bn_apply_body(const float* s_ss, const __half* __restrict__ x, const __half* __restrict__ res,
                                              int res_up, __half* __restrict__ y, size_t npix, int C, int H, int W, int act);

__global__ void __launch_bounds__(256)
bn_apply_kernel(const __half* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                const __half* __restrict__ res, int res_up, __half* __restrict__ y, size_t npix, int C, int H, int W, int act) {
    mg::pdl_prologue();
    extern __shared__ float s_ss[];  // [2][C]
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_ss[i] = scale[i], s_ss[C + i] = shift[i];
    __syncthreads();
    bn_apply_body(s_ss, x, res, res_up, y, npix, C, H, W, act);
}

// Training forward in ONE launch after the conv: every CTA adds up the statistic copies of the conv epilogue itself
// (2 C sums of STAT_COPIES values: a few loads per thread from L2) and derives scale / shift, then streams its share of
// the tensor; CTA 0 also publishes scale, shift, mean, 1/std for the backward and updates the running statistics.  Same
// arithmetic, same summation order as bn_finalize_kernel + bn_apply_kernel (which remain for exchanged statistics and
// for evaluation): a separate finalize launch cost more than the work it did, 71 times per step.
__global__ void __launch_bounds__(256)
bn_train_apply_kernel(const float* __restrict__ stats, float count, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps, float* __restrict__ out4,
                      const __half* __restrict__ x, const __half* __restrict__ res, int res_up, __half* __restrict__ y, size_t npix,
                      int C, int H, int W, int act) {
    mg::pdl_prologue();
    extern __shared__ float s_ss[];  // [2][C]: first the sums (sum | sum of squares), then scale | shift
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float part[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float p = 0.f;
#pragma unroll
            for (int k = g; k < STAT_COPIES; k += 8) p += __ldg(stats + (size_t)k * 2 * C + i);
            part[g] = p;
        }
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += part[g];
        s_ss[i] = t;
    }
    __syncthreads();
    float sc[2], sh[2];   // C <= 512: at most two channels per thread
    for (int c = threadIdx.x, j = 0; c < C; c += blockDim.x, ++j) {
        const float mean = s_ss[c] / count;
        const float var = fmaxf(s_ss[C + c] / count - mean * mean, 0.f);
        const float invstd = rsqrtf(var + eps);
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        sc[j] = g * invstd, sh[j] = b - mean * g * invstd;
        if (blockIdx.x == 0) {
            if (rmean) {
                rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
                rvar[c] = (1.f - momentum) * rvar[c] + momentum * var * (count > 1.f ? count / (count - 1.f) : 1.f);
            }
            out4[c] = sc[j], out4[C + c] = sh[j], out4[2 * C + c] = mean, out4[3 * C + c] = invstd;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x, j = 0; c < C; c += blockDim.x, ++j) s_ss[c] = sc[j], s_ss[C + c] = sh[j];
    __syncthreads();
    bn_apply_body(s_ss, x, res, res_up, y, npix, C, H, W, act);
}

__device__ __forceinline__ void bn_apply_body(const float* s_ss, const __half* __restrict__ x, const __half* __restrict__ res,
                                              int res_up, __half* __restrict__ y, size_t npix, int C, int H, int W, int act) {
    const int G = C >> 3, gs = (G & (G - 1)) == 0 ? __ffs(G) - 1 : -1;   // (a 64-bit division per 16-byte item otherwise)
    const size_t total = npix * G;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const size_t p = gs >= 0 ? v >> gs : v / G;
        const int g = (int)(v - p * G), c0 = g << 3;
        H8 in;
        in.u = __ldg(reinterpret_cast<const uint4*>(x + p * C + c0));
        float f[8];
        in.to_float(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], s_ss[c0 + i], s_ss[C + c0 + i]);
        if (res) {
            size_t rp = p;
            if (res_up) {
                const int xx = (int)(p % W), yy = (int)((p / W) % H);
                const size_t n = p / ((size_t)W * H);
                rp = (n * (H >> 1) + (yy >> 1)) * (W >> 1) + (xx >> 1);
            }
            H8 r;
            r.u = __ldg(reinterpret_cast<const uint4*>(res + rp * C + c0));
            float rf[8];
            r.to_float(rf);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += rf[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = act_fwd(f[i], act);
        H8 o;
        o.from_float(f);
        *reinterpret_cast<uint4*>(y + p * C + c0) = o.u;
    }
}

// sums[0][c] = sum dz, sums[1][c] = sum dz * xhat   with dz = dy * act'(y) (act after BN) or dy (act_first)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const __half* __restrict__ dy, const __half* __restrict__ yout, const __half* __restrict__ r,
                     const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ sums,
                     size_t npix, int C, int act) {
    mg::pdl_prologue();
    __shared__ __align__(16) float s_red[256 * 16];
    const int G = C >> 3, lanes_per_g = 256 / G;  // G in {4..64} divides 256
    const int g = threadIdx.x % G, sub = threadIdx.x / G, c0 = g << 3;
    float m[8], is[8], a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = mean[c0 + i], is[i] = invstd[c0 + i], a0[i] = 0.f, a1[i] = 0.f;
    // two pixels per iteration: 4..6 independent 16-byte loads in flight per thread (the kernel is a pure HBM stream)
    const size_t step = (size_t)gridDim.x * lanes_per_g;
    for (size_t p = (size_t)blockIdx.x * lanes_per_g + sub; p < npix; p += 2 * step) {
        const size_t p2 = p + step;
        const bool two = p2 < npix;
        H8 d[2], rr[2], yy[2];
        d[0].u = __ldg(reinterpret_cast<const uint4*>(dy + p * C + c0));
        rr[0].u = __ldg(reinterpret_cast<const uint4*>(r + p * C + c0));
        if (act) yy[0].u = __ldg(reinterpret_cast<const uint4*>(yout + p * C + c0));
        if (two) {
            d[1].u = __ldg(reinterpret_cast<const uint4*>(dy + p2 * C + c0));
            rr[1].u = __ldg(reinterpret_cast<const uint4*>(r + p2 * C + c0));
            if (act) yy[1].u = __ldg(reinterpret_cast<const uint4*>(yout + p2 * C + c0));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            float df[8], rf[8];
            d[h].to_float(df), rr[h].to_float(rf);
            if (act) {
                float yf[8];
                yy[h].to_float(yf);
#pragma unroll
                for (int i = 0; i < 8; ++i) df[i] *= act_grad_from_out(yf[i], act);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) a0[i] += df[i], a1[i] += df[i] * (rf[i] - m[i]) * is[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[threadIdx.x * 16 + i] = a0[i], s_red[threadIdx.x * 16 + 8 + i] = a1[i];
    __syncthreads();
    // thread t < 4*G: channel group t/4, four consecutive quantities -> sum over the lanes_per_g sub-rows, ONE 16-byte
    // vector reduction (the kernel was bound by same-address atomics: 888 blocks x 2C scalar atomicAdds on 2C addresses
    // took longer than streaming the three tensors out of L2)
    for (int t = threadIdx.x; t < 4 * G; t += 256) {
        const int gg = t >> 2, k4 = t & 3;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < lanes_per_g; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(&s_red[(s * G + gg) * 16 + 4 * k4]);
            acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        mg::ptx::red_add_v4(sums + (k4 >> 1) * C + (gg << 3) + 4 * (k4 & 1), acc.x, acc.y, acc.z, acc.w);
    }
}

// dx = scale * (dz - mean_dz - xhat * mean_dzx) [* relu'(r) when act_first]; dres = dz (optional)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const __half* __restrict__ dy, const __half* __restrict__ yout, const __half* __restrict__ r,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ sums, float inv_count, const float* __restrict__ count_dev,
                    __half* __restrict__ dx, __half* __restrict__ dres, size_t npix, int C, int act, int pre_act) {
    mg::pdl_prologue();
    if (count_dev) inv_count = 1.f / fmaxf(__ldg(count_dev), 1.f);
    extern __shared__ float s_p[];  // [4][C]: mean, scale=gamma*invstd, mean_dz, invstd*mean_dzx
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        const float is = invstd[i];
        s_p[i] = mean[i];
        s_p[C + i] = (gamma ? gamma[i] : 1.f) * is;
        s_p[2 * C + i] = sums[i] * inv_count;
        s_p[3 * C + i] = sums[C + i] * inv_count * is;
    }
    __syncthreads();
    const int G = C >> 3, gs = (G & (G - 1)) == 0 ? __ffs(G) - 1 : -1;
    const size_t total = npix * G;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
        const size_t p = gs >= 0 ? v >> gs : v / G;
        const int c0 = (int)(v - p * G) << 3;
        H8 d, rr;
        d.u = __ldg(reinterpret_cast<const uint4*>(dy + p * C + c0));
        rr.u = __ldg(reinterpret_cast<const uint4*>(r + p * C + c0));
        float df[8], rf[8], o[8];
        d.to_float(df), rr.to_float(rf);
        if (act) {
            H8 yy;
            yy.u = __ldg(reinterpret_cast<const uint4*>(yout + p * C + c0));
            float yf[8];
            yy.to_float(yf);
#pragma unroll
            for (int i = 0; i < 8; ++i) df[i] *= act_grad_from_out(yf[i], act);
        }
        if (dres) {
            H8 z;
            z.from_float(df);
            *reinterpret_cast<uint4*>(dres + p * C + c0) = z.u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + i;
            o[i] = s_p[C + c] * (df[i] - s_p[2 * C + c] - (rf[i] - s_p[c]) * s_p[3 * C + c]);
            if (pre_act) o[i] *= act_grad_from_out(rf[i], pre_act);
        }
        H8 w;
        w.from_float(o);
        *reinterpret_cast<uint4*>(dx + p * C + c0) = w.u;
    }
}

int stream_grid(size_t work_items) {
    return (int)std::min<size_t>((work_items + 255) / 256, (size_t)mg::kNumSMs * 8);
}

}  // namespace

extern "C" int mg_bn_finalize(const float* stats, float count, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, float momentum, float eps, float* scale, float* shift,
                              float* save_mean, float* save_invstd, int C, const float* count_dev, void* stream) {
    MG_REQUIRE(scale && shift && C > 0, "mg_bn_finalize: null pointer");
    MG_REQUIRE(stats || (running_mean && running_var), "mg_bn_finalize: need batch statistics or running statistics");
    MG_REQUIRE((save_mean == nullptr) == (save_invstd == nullptr), "mg_bn_finalize: save_mean/save_invstd go together");
    MG_LAUNCH(bn_finalize_kernel, mg::ceil_div(C, 32), 256, 0, stream, stats, count, count_dev, gamma, beta, running_mean, running_var,
              momentum, eps, scale, shift, save_mean, save_invstd, C);
    MG_CHECK_LAUNCH("mg_bn_finalize");
    return MG_OK;
}

extern "C" int mg_bn_apply(const void* x, const float* scale, const float* shift, const void* res, int res_up, void* y,
                           int N, int H, int W, int C, int act, void* stream) {
    MG_REQUIRE(x && scale && shift && y, "mg_bn_apply: null pointer");
    MG_REQUIRE(C % 8 == 0 && C <= 4096, "mg_bn_apply: C must be a multiple of 8 (got %d)", C);
    const size_t npix = (size_t)N * H * W;
    if (npix == 0) return MG_OK;
    MG_LAUNCH(bn_apply_kernel, stream_grid(npix * (C / 8)), 256, 2 * C * sizeof(float), stream,
              static_cast<const __half*>(x), scale, shift, static_cast<const __half*>(res), res_up,
              static_cast<__half*>(y), npix, C, H, W, act);
    MG_CHECK_LAUNCH("mg_bn_apply");
    return MG_OK;
}

extern "C" int mg_bn_train_apply(const float* stats, float count, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, float momentum, float eps, float* out4, const void* x, const void* res,
                                 int res_up, void* y, int N, int H, int W, int C, int act, void* stream) {
    MG_REQUIRE(stats && out4 && x && y, "mg_bn_train_apply: null pointer");
    MG_REQUIRE(C % 8 == 0 && C <= 512, "mg_bn_train_apply: C must be a multiple of 8, at most 512 (got %d)", C);
    MG_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "mg_bn_train_apply: running_mean/running_var go together");
    const size_t npix = (size_t)N * H * W;
    MG_REQUIRE(npix > 0, "mg_bn_train_apply: empty tensor");
    MG_LAUNCH(bn_train_apply_kernel, stream_grid(npix * (C / 8)), 256, 2 * C * sizeof(float), stream, stats, count, gamma, beta,
              running_mean, running_var, momentum, eps, out4, static_cast<const __half*>(x), static_cast<const __half*>(res), res_up,
              static_cast<__half*>(y), npix, C, H, W, act);
    MG_CHECK_LAUNCH("mg_bn_train_apply");
    return MG_OK;
}

extern "C" int mg_bn_bwd_reduce(const void* dy, const void* y, const void* conv_out, const float* mean, const float* invstd,
                                float* sums, int N, int H, int W, int C, int act, void* stream) {
    MG_REQUIRE(dy && conv_out && mean && invstd && sums && (y || !act), "mg_bn_bwd_reduce: null pointer");
    MG_REQUIRE(C >= 32 && C <= 512 && (C & (C - 1)) == 0, "mg_bn_bwd_reduce: C must be a power of two in 32..512 (got %d)", C);
    const size_t npix = (size_t)N * H * W;
    if (npix == 0) return MG_OK;
    const int rows_per_block = 256 / (C / 8);
    static const int waves = [] { const char* e = std::getenv("MAGGIE_B200_BN_REDUCE_WAVES"); return e ? std::atoi(e) : 2; }();
    const int grid = (int)std::min<size_t>((npix + 2 * rows_per_block - 1) / (2 * rows_per_block), (size_t)mg::kNumSMs * waves);
    MG_LAUNCH(bn_bwd_reduce_kernel, grid, 256, 0, stream, static_cast<const __half*>(dy), static_cast<const __half*>(y),
              static_cast<const __half*>(conv_out), mean, invstd, sums, npix, C, act);
    MG_CHECK_LAUNCH("mg_bn_bwd_reduce");
    return MG_OK;
}

extern "C" int mg_bn_bwd_apply(const void* dy, const void* y, const void* conv_out, const float* mean, const float* invstd,
                               const float* gamma, const float* sums, void* dx, void* dres, int N, int H, int W, int C,
                               int act, int pre_act, const float* count_dev, void* stream) {
    MG_REQUIRE(dy && conv_out && mean && invstd && sums && dx && (y || !act), "mg_bn_bwd_apply: null pointer");
    MG_REQUIRE(C % 8 == 0 && C <= 2048, "mg_bn_bwd_apply: C must be a multiple of 8 (got %d)", C);
    const size_t npix = (size_t)N * H * W;
    if (npix == 0) return MG_OK;
    MG_LAUNCH(bn_bwd_apply_kernel, stream_grid(npix * (C / 8)), 256, 4 * C * sizeof(float), stream,
              static_cast<const __half*>(dy), static_cast<const __half*>(y), static_cast<const __half*>(conv_out), mean,
              invstd, gamma, sums, 1.0f / (float)npix, count_dev, static_cast<__half*>(dx), static_cast<__half*>(dres), npix, C, act,
              pre_act);
    MG_CHECK_LAUNCH("mg_bn_bwd_apply");
    return MG_OK;
}
