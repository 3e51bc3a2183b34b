// K12: the training losses evaluated inside forward() - weighted L1, 3-level Laplacian pyramid, Sobel gradient -
// for the three alpha scales (OS1, OS4, OS8) in a handful of fused stencil-reduction kernels, forward and backward.
//
// The pyramid is linear, so it is built ONCE per scale on d = pred - target:
//     d_{k+1} = D G d_k  (5x5 binomial, reflect pad, keep even pixels),   L_k = d_k - U d_{k+1}  (zero insert, 4G)
//     lap_k   = sum |L_k| * w_k,   w_k = w[::2^k, ::2^k]
// Forward stores s_k = sign(L_k) * w_k (fp16); backward runs the adjoint pyramid coarse -> fine with gathers only:
//     g_k = c_k s_k + (DG)^T g_{k+1} - U^T (c_{k-1} s_{k-1})
// and the full-resolution kernel adds the weighted-L1 and Sobel-magnitude terms.  Everything is HBM-bound:
// ~ (3 scales) x (pred, weight: 8 B/px read) + target 4 B/px, s_0 2 B/px written.
#include "common.cuh"

namespace {

constexpr int NCOPY = MG_LOSS_COPIES;  // partial-sum copies to keep the atomics uncontended
constexpr int NQ = 8;                  // rec, lap0, lap1, lap2, sobel, wsum0, wsum1, wsum2

__device__ __forceinline__ int refl(int p, int n) { return p < 0 ? -p : (p >= n ? 2 * (n - 1) - p : p); }
__device__ __forceinline__ int clampi(int p, int n) { return p < 0 ? 0 : (p >= n ? n - 1 : p); }
__device__ __forceinline__ float gk(int k) { return k == 0 || k == 4 ? 0.0625f : (k == 2 ? 0.375f : 0.25f); }  // [1,4,6,4,1]/16
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

struct Ptr3 {
    const float* p[3];
};
struct MPtr3 {
    float* p[3];
};

__device__ __forceinline__ void block_accumulate(float* vals, int n, float* dst) {
    // warp shuffle + one atomic per warp per quantity
#pragma unroll 1
    for (int q = 0; q < n; ++q) {
        float v = vals[q];
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        const int lin = threadIdx.y * blockDim.x + threadIdx.x;
        if ((lin & 31) == 0 && v != 0.f) atomicAdd(dst + q, v);
    }
}

// ---- F1: full-resolution pass: weighted L1, Sobel term, weight sum; writes d1 = D G (p - t) -------------------
// grid (W/32, H/32, 3*S), block 16x16; each thread owns a 2x2 pixel quad and one d1 output.
__global__ void __launch_bounds__(256)
loss_level0_kernel(Ptr3 P, const float* __restrict__ T, Ptr3 Wt, const float* __restrict__ plane_scale, float* __restrict__ D1,
                   float* __restrict__ sums, uint8_t* __restrict__ tile_raw, int S, int H, int W) {
    mg::pdl_prologue();
    __shared__ float s_d[36][37], s_pw[34][35], s_tw[34][35];
    const int scale = blockIdx.z / S, sl = blockIdx.z - scale * S;
    const float* p = P.p[scale] + (size_t)sl * H * W;
    const float* t = T + (size_t)sl * H * W;
    const float* w = Wt.p[scale] + (size_t)sl * H * W;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32, tid = threadIdx.y * 16 + threadIdx.x;
    const float ps = plane_scale ? __ldg(plane_scale + sl) : 1.f;   // per-plane factor on the prediction (`pred * valid`)
    for (int i = tid; i < 36 * 36; i += 256) {
        const int ly = i / 36, lx = i - ly * 36;
        const int gy = refl(y0 + ly - 2, H), gx = refl(x0 + lx - 2, W);
        s_d[ly][lx] = p[(size_t)gy * W + gx] * ps - t[(size_t)gy * W + gx];
    }
    int any_w = 0;
    for (int i = tid; i < 34 * 34; i += 256) {
        const int ly = i / 34, lx = i - ly * 34;
        const int gy = clampi(y0 + ly - 1, H), gx = clampi(x0 + lx - 1, W);
        const float ww = w[(size_t)gy * W + gx];
        any_w |= ww != 0.f;
        s_pw[ly][lx] = ww != 0.f ? p[(size_t)gy * W + gx] * ps * ww : 0.f;
        s_tw[ly][lx] = ww != 0.f ? t[(size_t)gy * W + gx] * ww : 0.f;
    }
    // tiles without any weight (most of the OS1 / OS4 scales) contribute nothing to the weighted-L1 / Sobel / weight sums
    const bool weighted = __syncthreads_or(any_w) != 0;
    // tile map for the later passes: 1 = some weight inside this 32 x 32 tile (+ 1 pixel of halo)
    if (tid == 0) tile_raw[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = weighted ? 1 : 0;
    float acc[3] = {0.f, 0.f, 0.f};  // rec, sobel, wsum
#pragma unroll
    for (int qy = 0; qy < 2; ++qy)
#pragma unroll
        for (int qx = 0; qx < 2; ++qx) {
            const int ly = threadIdx.y * 2 + qy, lx = threadIdx.x * 2 + qx;  // tile-local pixel
            if (!weighted || y0 + ly >= H || x0 + lx >= W) continue;
            const int cy = ly + 1, cx = lx + 1;                              // index into the halo-1 arrays
            acc[0] += fabsf(s_pw[cy][cx] - s_tw[cy][cx]);
            acc[2] += w[(size_t)(y0 + ly) * W + x0 + lx];
            float m[2];
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                float (*a)[35] = which ? s_tw : s_pw;
                const float gx = (-a[cy - 1][cx - 1] + a[cy - 1][cx + 1] - 2.f * a[cy][cx - 1] + 2.f * a[cy][cx + 1] -
                                  a[cy + 1][cx - 1] + a[cy + 1][cx + 1]) * 0.125f;
                const float gy = (-a[cy - 1][cx - 1] - 2.f * a[cy - 1][cx] - a[cy - 1][cx + 1] + a[cy + 1][cx - 1] +
                                  2.f * a[cy + 1][cx] + a[cy + 1][cx + 1]) * 0.125f;
                m[which] = sqrtf(gx * gx + gy * gy + 1e-6f);
            }
            acc[1] += fabsf(m[0] - m[1]);
        }
    // d1 output (i, j) of this thread: centre at tile-local (2*ty, 2*tx) -> s_d index +2
    {
        const int oy = (y0 >> 1) + threadIdx.y, ox = (x0 >> 1) + threadIdx.x;
        if (oy < (H >> 1) && ox < (W >> 1)) {
            float v = 0.f;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
                float r = 0.f;
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) r += gk(kx) * s_d[threadIdx.y * 2 + ky][threadIdx.x * 2 + kx];
                v += gk(ky) * r;
            }
            D1[((size_t)(scale * S + sl) * (H >> 1) + oy) * (W >> 1) + ox] = v;
        }
    }
    float* dst = sums + ((size_t)((blockIdx.x + blockIdx.y * 7 + blockIdx.z * 13) % NCOPY) * 3 + scale) * NQ;
    float out3[3] = {acc[0], acc[1], acc[2]};
    block_accumulate(&out3[0], 1, dst + 0);
    block_accumulate(&out3[1], 1, dst + 4);
    block_accumulate(&out3[2], 1, dst + 5);
}

// ---- tile map: near[tile] = some weight within the 3 x 3 tile neighbourhood (+-32 pixels).  The Laplacian terms, their
// stored signs and every level of the adjoint pyramid are EXACTLY zero at pixels of tiles that are not `near` (the reach
// of the three pyramid levels is 2 + 4 + 8 + 12 = 26 fine pixels), and the OS1 / OS4 weights are ~5 % wide bands: the
// passes below skip such tiles (forward: nothing to add, nothing stored; backward: zeros written, nothing read).
__global__ void __launch_bounds__(256)
loss_tiles_near_kernel(const uint8_t* __restrict__ raw, uint8_t* __restrict__ near, int planes, int TY, int TX) {
    mg::pdl_prologue();
    const int total = planes * TY * TX;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tx = i % TX, ty = (i / TX) % TY;
        const uint8_t* base = raw + (size_t)(i / (TX * TY)) * TY * TX;
        uint8_t v = 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = ty + dy, xx = tx + dx;
                if (yy >= 0 && yy < TY && xx >= 0 && xx < TX) v |= base[yy * TX + xx];
            }
        near[i] = v;
    }
}

// near-flag of the tile holding level-`level` pixel (y, x) of plane `plane` (= scale * S + slice)
__device__ __forceinline__ bool tile_near(const uint8_t* __restrict__ near, size_t plane, int y, int x, int level, int TY, int TX) {
    return near[(plane * TY + ((y << level) >> 5)) * TX + ((x << level) >> 5)] != 0;
}

// ---- F2: d_{k+1} = D G d_k on the small levels (one thread per output) ----------------------------------------
__global__ void __launch_bounds__(256)
loss_down_kernel(const float* __restrict__ din, float* __restrict__ dout, int n_img, int h, int w) {
    mg::pdl_prologue();
    const int ho = h >> 1, wo = w >> 1;
    // (32-bit index arithmetic: every plane set of the path is < 2^31 elements (checked by the launcher), and the 64-bit
    //  quotients / remainders by run-time sizes were a large share of these per-pixel kernels' instructions)
    const unsigned total = (unsigned)n_img * ho * wo, uwo = wo, uho = ho;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned q = i / uwo;
        const int ox = (int)(i - q * uwo), oy = (int)(q % uho);
        const float* src = din + (size_t)(q / uho) * h * w;
        float v = 0.f;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
            const int yy = refl(2 * oy + ky - 2, h);
            float r = 0.f;
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) r += gk(kx) * __ldg(src + (size_t)yy * w + refl(2 * ox + kx - 2, w));
            v += gk(ky) * r;
        }
        dout[i] = v;
    }
}

// U d_{k+1} at fine pixel (y,x): zero-inserted upsampling followed by 4*G with reflect padding.
__device__ __forceinline__ float upsample_at(const float* __restrict__ dc, int hc, int wc, int y, int x, int h, int w) {
    // Only taps that land on an even (= non-zero-inserted) fine pixel contribute, and reflection keeps parity (h, w are
    // even): k = y & 1, y & 1 + 2, ... -> 3 x 3 taps on even / 2 x 2 on odd coordinates instead of 25 predicated ones.
    float v = 0.f;
    const int ky0 = y & 1, kx0 = x & 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int ky = ky0 + 2 * a;
        if (ky > 4) break;
        const float* row = dc + (size_t)(refl(y + ky - 2, h) >> 1) * wc;
        float r = 0.f;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int kx = kx0 + 2 * b;
            if (kx > 4) break;
            r += gk(kx) * __ldg(row + (refl(x + kx - 2, w) >> 1));
        }
        v += gk(ky) * r;
    }
    return 4.f * v;
}

// ---- F3: Laplacian level k: L = d_k - U d_{k+1}; lap_k += |L| w_k; s_k = sign(L) w_k --------------------------
// level 0: d_0 = p - t on the fly (dk == nullptr).  wstep = 2^k (sub-sampling of the full-resolution weight).
__global__ void __launch_bounds__(256)
loss_lap_kernel(Ptr3 P, const float* __restrict__ T, const float* __restrict__ dk, const float* __restrict__ dk1, Ptr3 Wt,
                const float* __restrict__ plane_scale, __half* __restrict__ sg, float* __restrict__ sums,
                const uint8_t* __restrict__ near, int TY, int TX, int S, int H, int W, int level) {
    mg::pdl_prologue();
    const int h = H >> level, w = W >> level, hc = h >> 1, wc = w >> 1, wstep = 1 << level;
    const int scale = blockIdx.y;                                   // one scale per grid row: partial sums never mix
    const size_t per_scale = (size_t)S * h * w, base = scale * per_scale;
    float acc[2] = {0.f, 0.f};
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < (unsigned)per_scale; r += gridDim.x * blockDim.x) {
        const unsigned qy = r / (unsigned)w, qs = qy / (unsigned)h;
        const int x = (int)(r - qy * (unsigned)w), y = (int)(qy - qs * (unsigned)h), sl = (int)qs;
        if (!tile_near(near, (size_t)scale * S + sl, y, x, level, TY, TX)) {
            sg[base + r] = __float2half(0.f);   // nothing to add, nothing read; (the adjoint's taps may reach into this tile)
            continue;
        }
        // the Laplacian value only matters under a non-zero weight (the OS1 / OS4 weights are the ~5 % wide refinement
        // bands): everywhere else the term and its stored sign are exactly zero, and neither the difference image nor the
        // 3x3 coarse neighbourhood is read
        const float ww = Wt.p[scale][((size_t)sl * H + (size_t)y * wstep) * W + (size_t)x * wstep];
        float sv = 0.f;
        if (ww != 0.f) {
            float d;
            if (dk) d = dk[base + r];
            else {
                const size_t o = ((size_t)sl * H + y) * W + x;
                d = P.p[scale][o] * (plane_scale ? __ldg(plane_scale + sl) : 1.f) - T[o];
            }
            const float L = d - upsample_at(dk1 + (size_t)(scale * S + sl) * hc * wc, hc, wc, y, x, h, w);
            acc[0] += fabsf(L) * ww;
            acc[1] += ww;
            sv = sgn(L) * ww;
        }
        sg[base + r] = __float2half(sv);
    }
    float* dst = sums + ((size_t)(blockIdx.x % NCOPY) * 3 + scale) * NQ;
    block_accumulate(&acc[0], 1, dst + 1 + level);
    if (level > 0) block_accumulate(&acc[1], 1, dst + 5 + level);
}

// ---- backward ---------------------------------------------------------------------------------------------
// 1-D adjoint weights.  A(i,y)  = sum_k g[k]  [refl(2i+k-2, n) == y]   ((DG)^T : coarse i -> fine y)
//                       B(y,i)  = sum_k 2g[k] [refl(y+k-2, n) == 2i]   (U^T     : fine y   -> coarse i; 2 per axis = the 4x of U)
__device__ __forceinline__ float adjA(int i, int y, int n) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (refl(2 * i + k - 2, n) == y) a += gk(k);
    return a;
}
__device__ __forceinline__ float adjB(int y, int i, int n) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (refl(y + k - 2, n) == 2 * i) a += 2.f * gk(k);
    return a;
}

// Non-zero taps of the two 1-D adjoints for output index o.  Interior indices use the closed form (2-3 taps for
// (DG)^T, 5 taps for U^T); the few indices whose 5-tap window crosses the image border fall back to the generic
// "which taps reflect onto me" search.
struct Taps {
    int n, idx[5];
    float wt[5];
};
__device__ __forceinline__ void taps_DGt(int y, int n_fine, Taps& t) {   // coarse i feeding fine y
    t.n = 0;
    const int nc = n_fine >> 1;
    if (y >= 3 && y + 3 < n_fine) {   // no reflected tap lands on y (reflections reach 1, 2 and n-2 only)
        if (y & 1) {
            t.idx[0] = (y - 1) >> 1, t.wt[0] = 0.25f, t.idx[1] = (y + 1) >> 1, t.wt[1] = 0.25f, t.n = 2;
        } else {
            t.idx[0] = (y >> 1) - 1, t.wt[0] = 0.0625f, t.idx[1] = y >> 1, t.wt[1] = 0.375f, t.idx[2] = (y >> 1) + 1,
            t.wt[2] = 0.0625f, t.n = 3;
        }
        return;
    }
    const int i0 = max(0, ((y - 2) >> 1) - 1), i1 = min(nc - 1, ((y + 2) >> 1) + 1);
    for (int i = i0; i <= i1 && t.n < 5; ++i) {
        const float a = adjA(i, y, n_fine);
        if (a != 0.f) t.idx[t.n] = i, t.wt[t.n] = a, ++t.n;
    }
}
__device__ __forceinline__ void taps_Ut(int i, int n_fine, Taps& t) {    // fine y feeding coarse i
    t.n = 0;
    if (i >= 2 && 2 * i + 2 < n_fine - 1) {   // reflections reach coarse 1 and n/2-1 only
#pragma unroll
        for (int k = 0; k < 5; ++k) t.idx[k] = 2 * i + 2 - k, t.wt[k] = 2.f * gk(k);
        t.n = 5;
        return;
    }
    const int y0 = max(0, 2 * i - 3), y1 = min(n_fine - 1, 2 * i + 3);
    for (int y = y0; y <= y1 && t.n < 5; ++y) {
        const float b = adjB(y, i, n_fine);
        if (b != 0.f) t.idx[t.n] = y, t.wt[t.n] = b, ++t.n;
    }
}

// g_k[y,x] = c_k s_k + sum_{i,j} A(i,y) A(j,x) g_{k+1}[i,j] - sum_{yy,xx} B(yy,y) B(xx,x) c_{k-1} s_{k-1}[yy,xx]
// (level k has size h x w; k+1 is h/2 x w/2; k-1 is 2h x 2w).  Any of the three terms may be absent.
__device__ __forceinline__ float bwd_level_value(const __half* __restrict__ sg_k, float c_k, const float* __restrict__ g_k1,
                                                 const __half* __restrict__ sg_km1, float c_km1, size_t img, int y, int x, int h,
                                                 int w) {
    float v = 0.f;
    if (sg_k) v = c_k * __half2float(sg_k[(img * h + y) * w + x]);
    // Interior pixels (all but a 3-pixel frame) take the closed forms of taps_DGt / taps_Ut unrolled in registers, in the
    // same summation order (a zero third tap for odd indices adds exactly 0): the generic tap lists are run-time sized
    // arrays, i.e. local memory and divergent loops in every thread of these per-pixel kernels.
    if (g_k1) {
        const int hc = h >> 1, wc = w >> 1;
        if (y >= 3 && y + 3 < h && x >= 3 && x + 3 < w) {
            const int oy = y & 1, ox = x & 1;
            const float* gp = g_k1 + (img * hc + (y >> 1) - 1 + oy) * wc + (x >> 1) - 1 + ox;
            const float wy[3] = {oy ? 0.25f : 0.0625f, oy ? 0.25f : 0.375f, oy ? 0.f : 0.0625f};
            const float wx[3] = {ox ? 0.25f : 0.0625f, ox ? 0.25f : 0.375f, ox ? 0.f : 0.0625f};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float r = 0.f;
#pragma unroll
                for (int b = 0; b < 3; ++b) r += wx[b] * __ldg(gp + a * wc + b);
                v += wy[a] * r;
            }
        } else {
            Taps ty, tx;
            taps_DGt(y, h, ty);
            taps_DGt(x, w, tx);
            for (int a = 0; a < ty.n; ++a) {
                float r = 0.f;
                for (int b = 0; b < tx.n; ++b) r += tx.wt[b] * __ldg(g_k1 + (img * hc + ty.idx[a]) * wc + tx.idx[b]);
                v += ty.wt[a] * r;
            }
        }
    }
    if (sg_km1) {
        const int hf = h << 1, wf = w << 1;
        float u = 0.f;
        if (y >= 2 && 2 * y + 2 < hf - 1 && x >= 2 && 2 * x + 2 < wf - 1) {
            const __half* sp = sg_km1 + (img * hf + 2 * y + 2) * wf + 2 * x + 2;
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                float r = 0.f;
#pragma unroll
                for (int b = 0; b < 5; ++b) r += 2.f * gk(b) * __half2float(sp[-(a * wf) - b]);
                u += 2.f * gk(a) * r;
            }
        } else {
            Taps ty, tx;
            taps_Ut(y, hf, ty);
            taps_Ut(x, wf, tx);
            for (int a = 0; a < ty.n; ++a) {
                float r = 0.f;
                for (int b = 0; b < tx.n; ++b) r += tx.wt[b] * __half2float(sg_km1[(img * hf + ty.idx[a]) * wf + tx.idx[b]]);
                u += ty.wt[a] * r;
            }
        }
        v -= c_km1 * u;
    }
    return v;
}

// coefficient layout: coef[scale][5] = upstream gradient of (rec, lap0, lap1, lap2, sobel) numerators
__global__ void __launch_bounds__(256)
loss_bwd_small_kernel(const __half* __restrict__ sg_k, const float* __restrict__ g_k1, const __half* __restrict__ sg_km1,
                      const float* __restrict__ coef, float* __restrict__ g_out, const uint8_t* __restrict__ near, int TY, int TX,
                      int S, int h, int w, int level) {
    mg::pdl_prologue();
    const size_t per_scale = (size_t)S * h * w, total = 3 * per_scale;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x * blockDim.x) {
        const int scale = (int)(i / (unsigned)per_scale);
        const unsigned qy = i / (unsigned)w, img = qy / (unsigned)h;
        const int x = (int)(i - qy * (unsigned)w), y = (int)(qy - img * (unsigned)h);
        if (!tile_near(near, img, y, x, level, TY, TX)) {     // exactly zero there (see loss_tiles_near_kernel)
            g_out[i] = 0.f;
            continue;
        }
        const float c_k = level <= 2 ? coef[scale * 5 + 1 + level] : 0.f;
        const float c_km1 = coef[scale * 5 + level];  // level >= 1 here: lap_{level-1}
        g_out[i] = bwd_level_value(sg_k, c_k, g_k1, sg_km1, c_km1, img, y, x, h, w);
    }
}

// full-resolution backward: lap (level 0) + weighted L1 + Sobel; one gradient tensor per scale.
// grid (W/32, H/32, 3*S), block 32x8 (each thread 4 rows).
__global__ void __launch_bounds__(256)
loss_bwd_level0_kernel(Ptr3 P, const float* __restrict__ T, Ptr3 Wt, const float* __restrict__ plane_scale,
                       const __half* __restrict__ sg0, const float* __restrict__ g1, const float* __restrict__ coef, MPtr3 G,
                       const uint8_t* __restrict__ near, int S, int H, int W) {
    mg::pdl_prologue();
    __shared__ float s_pw[36][37], s_tw[36][37], s_gx[34][35], s_gy[34][35];
    const int scale = blockIdx.z / S, sl = blockIdx.z - scale * S;
    if (!near[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x]) {
        // no weight within 32 pixels: every term of the gradient is exactly zero here
        float* gp = G.p[scale] + (size_t)sl * H * W;
        for (int row = threadIdx.y; row < 32; row += 8) {
            const int y = blockIdx.y * 32 + row, x = blockIdx.x * 32 + threadIdx.x;
            if (y < H && x < W) gp[(size_t)y * W + x] = 0.f;
        }
        return;
    }
    const float* p = P.p[scale] + (size_t)sl * H * W;
    const float* t = T + (size_t)sl * H * W;
    const float* w = Wt.p[scale] + (size_t)sl * H * W;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32, tid = threadIdx.y * 32 + threadIdx.x;
    const float c_rec = coef[scale * 5 + 0], c_lap0 = coef[scale * 5 + 1], c_sob = coef[scale * 5 + 4];
    const float ps = plane_scale ? __ldg(plane_scale + sl) : 1.f;   // prediction = p * ps; the gradient w.r.t. p gets ps too
    int any_w = 0;
    for (int i = tid; i < 36 * 36; i += 256) {
        const int ly = i / 36, lx = i - ly * 36;
        const int gy = clampi(y0 + ly - 2, H), gx = clampi(x0 + lx - 2, W);
        const float ww = w[(size_t)gy * W + gx];
        any_w |= ww != 0.f;
        s_pw[ly][lx] = ww != 0.f ? p[(size_t)gy * W + gx] * ps * ww : 0.f;
        s_tw[ly][lx] = ww != 0.f ? t[(size_t)gy * W + gx] * ww : 0.f;
    }
    // Tiles whose weights are zero over the whole halo (most tiles of the OS1 / OS4 scales: their weights are the narrow
    // refinement bands) have no weighted-L1 and no Sobel contribution at all: only the Laplacian adjoint remains.
    const bool weighted = __syncthreads_or(any_w) != 0;
    // d(sum |m_p - m_t|) / d(gx_p, gy_p) at the tile + halo 1 (only for output pixels inside the image)
    for (int i = tid; weighted && i < 34 * 34; i += 256) {
        const int ly = i / 34, lx = i - ly * 34;
        const int oy = y0 + ly - 1, ox = x0 + lx - 1;
        float ggx = 0.f, ggy = 0.f;
        if (oy >= 0 && oy < H && ox >= 0 && ox < W) {
            const int cy = ly + 1, cx = lx + 1;  // position in the halo-2 arrays
            float gx[2], gy[2], m[2];
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                float (*a)[37] = which ? s_tw : s_pw;
                // replicate padding: neighbours outside the image take the clamped pixel, which the halo load already did
                gx[which] = (-a[cy - 1][cx - 1] + a[cy - 1][cx + 1] - 2.f * a[cy][cx - 1] + 2.f * a[cy][cx + 1] - a[cy + 1][cx - 1] +
                             a[cy + 1][cx + 1]) * 0.125f;
                gy[which] = (-a[cy - 1][cx - 1] - 2.f * a[cy - 1][cx] - a[cy - 1][cx + 1] + a[cy + 1][cx - 1] + 2.f * a[cy + 1][cx] +
                             a[cy + 1][cx + 1]) * 0.125f;
                m[which] = sqrtf(gx[which] * gx[which] + gy[which] * gy[which] + 1e-6f);
            }
            const float gm = c_sob * sgn(m[0] - m[1]) / m[0];
            ggx = gm * gx[0], ggy = gm * gy[0];
        }
        s_gx[ly][lx] = ggx, s_gy[ly][lx] = ggy;
    }
    __syncthreads();
    const size_t img = (size_t)scale * S + sl;
    for (int row = threadIdx.y; row < 32; row += 8) {
        const int y = y0 + row, x = x0 + threadIdx.x;
        if (y >= H || x >= W) continue;
        const size_t o = (size_t)y * W + x;
        // Laplacian level 0 + (DG)^T g_1
        float g = bwd_level_value(sg0, c_lap0, g1, nullptr, 0.f, img, y, x, H, W);
        if (!weighted) {
            G.p[scale][(size_t)sl * H * W + o] = g * ps;
            continue;
        }
        const float ww = w[o];
        // weighted L1
        g += c_rec * sgn(s_pw[row + 2][threadIdx.x + 2] - s_tw[row + 2][threadIdx.x + 2]) * ww;
        // Sobel adjoint (replicate padding: taps that were clamped onto this pixel come back to it)
        float sa = 0.f;
        if (y >= 1 && y + 1 < H && x >= 1 && x + 1 < W) {
            // interior: output (y+dy, x+dx) reaches (y,x) through exactly one tap (ky,kx) = (1-dy, 1-dx)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const float ggx = s_gx[row + 1 + dy][threadIdx.x + 1 + dx], ggy = s_gy[row + 1 + dy][threadIdx.x + 1 + dx];
                    const float a_y = dy == 0 ? 2.f : 1.f, a_x = dx == 0 ? 2.f : 1.f;
                    sa += (a_y * (float)(-dx) * ggx + (float)(-dy) * a_x * ggy) * 0.125f;
                }
        } else
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int oy = y + dy, ox = x + dx;  // an output pixel whose 3x3 window may touch (y,x)
                if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;
                const float ggx = s_gx[row + 1 + dy][threadIdx.x + 1 + dx], ggy = s_gy[row + 1 + dy][threadIdx.x + 1 + dx];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    if (clampi(oy + ky - 1, H) != y) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        if (clampi(ox + kx - 1, W) != x) continue;
                        const float a_y = ky == 1 ? 2.f : 1.f, a_x = kx == 1 ? 2.f : 1.f;
                        const float b_y = (float)(ky - 1), b_x = (float)(kx - 1);
                        sa += (a_y * b_x * ggx + b_y * a_x * ggy) * 0.125f;
                    }
                }
            }
        g += sa * ww;
        G.p[scale][(size_t)sl * H * W + o] = g * ps;
    }
}

int small_grid(size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)mg::kNumSMs * 16); }

}  // namespace

extern "C" size_t mg_loss_workspace_floats(int S, int H, int W) {
    // d1, d2, d3 (fp32) + g1, g2, g3 (fp32) for three scales
    const size_t l1 = (size_t)3 * S * (H / 2) * (W / 2), l2 = l1 / 4, l3 = l2 / 4;
    // + the two tile maps (raw / near), one byte per 32 x 32 tile and plane
    const size_t tiles = (size_t)3 * S * ((H + 31) / 32) * ((W + 31) / 32);
    return 2 * (l1 + l2 + l3) + 2 * ((tiles + 3) / 4) + 8;
}

static inline uint8_t* tile_maps(float* ws, int S, int H, int W, size_t* n_tiles) {
    const size_t l1 = (size_t)3 * S * (H / 2) * (W / 2), l2 = l1 / 4, l3 = l2 / 4;
    *n_tiles = (size_t)3 * S * ((H + 31) / 32) * ((W + 31) / 32);
    return reinterpret_cast<uint8_t*>(ws + 2 * (l1 + l2 + l3));
}

extern "C" int mg_loss_fwd(const float* a1, const float* a4, const float* a8, const float* target, const float* w1,
                           const float* w4, const float* w8, const float* plane_scale, int S, int H, int W, float* ws,
                           void* sg_f16, float* sums, void* stream) {
    MG_REQUIRE(a1 && a4 && a8 && target && w1 && w4 && w8 && ws && sg_f16 && sums, "mg_loss_fwd: null pointer");
    MG_REQUIRE(S > 0 && H % 8 == 0 && W % 8 == 0 && H >= 16 && W >= 16, "mg_loss_fwd: H, W must be multiples of 8, >= 16");
    MG_REQUIRE(3 * S <= 65535, "mg_loss_fwd: too many slices");
    MG_REQUIRE((size_t)3 * S * H * W < ((size_t)1 << 31), "mg_loss_fwd: 3 * S * H * W must stay below 2^31 (32-bit pixel indices)");
    Ptr3 P{{a1, a4, a8}}, Wt{{w1, w4, w8}};
    const size_t n0 = (size_t)3 * S * H * W, n1 = n0 / 4, n2 = n1 / 4, n3 = n2 / 4;
    float *d1 = ws, *d2 = d1 + n1, *d3 = d2 + n2;
    __half* sg0 = static_cast<__half*>(sg_f16);
    __half *sg1 = sg0 + n0, *sg2 = sg1 + n1;
    dim3 grid0(mg::ceil_div(W, 32), mg::ceil_div(H, 32), 3 * S);
    size_t n_tiles;
    uint8_t* t_raw = tile_maps(ws, S, H, W, &n_tiles);
    uint8_t* t_near = t_raw + ((n_tiles + 3) / 4) * 4;
    const int TY = (int)grid0.y, TX = (int)grid0.x;
    MG_LAUNCH(loss_level0_kernel, grid0, dim3(16, 16), 0, stream, P, target, Wt, plane_scale, d1, sums, t_raw, S, H, W);
    MG_LAUNCH(loss_tiles_near_kernel, small_grid(n_tiles), 256, 0, stream, (const uint8_t*)t_raw, t_near, 3 * S, TY, TX);
    MG_LAUNCH(loss_down_kernel, small_grid(n2), 256, 0, stream, d1, d2, 3 * S, H / 2, W / 2);
    MG_LAUNCH(loss_down_kernel, small_grid(n3), 256, 0, stream, d2, d3, 3 * S, H / 4, W / 4);
    const uint8_t* nr = t_near;
    MG_LAUNCH(loss_lap_kernel, dim3(small_grid(n0 / 3), 3), 256, 0, stream, P, target, (const float*)nullptr, d1, Wt, plane_scale, sg0, sums, nr, TY, TX, S, H, W, 0);
    MG_LAUNCH(loss_lap_kernel, dim3(small_grid(n1 / 3), 3), 256, 0, stream, P, target, d1, d2, Wt, plane_scale, sg1, sums, nr, TY, TX, S, H, W, 1);
    MG_LAUNCH(loss_lap_kernel, dim3(small_grid(n2 / 3), 3), 256, 0, stream, P, target, d2, d3, Wt, plane_scale, sg2, sums, nr, TY, TX, S, H, W, 2);
    MG_CHECK_LAUNCH("mg_loss_fwd");
    return MG_OK;
}

extern "C" int mg_loss_bwd(const float* a1, const float* a4, const float* a8, const float* target, const float* w1,
                           const float* w4, const float* w8, const float* plane_scale, int S, int H, int W, float* ws,
                           const void* sg_f16,
                           const float* coef, float* g1_out, float* g4_out, float* g8_out, void* stream) {
    MG_REQUIRE(a1 && a4 && a8 && target && w1 && w4 && w8 && ws && sg_f16 && coef && g1_out && g4_out && g8_out,
               "mg_loss_bwd: null pointer");
    MG_REQUIRE((size_t)3 * S * H * W < ((size_t)1 << 31), "mg_loss_bwd: 3 * S * H * W must stay below 2^31 (32-bit pixel indices)");
    Ptr3 P{{a1, a4, a8}}, Wt{{w1, w4, w8}};
    MPtr3 G{{g1_out, g4_out, g8_out}};
    const size_t n0 = (size_t)3 * S * H * W, n1 = n0 / 4, n2 = n1 / 4, n3 = n2 / 4;
    float* gbase = ws + (n1 + n2 + n3);
    float *g1 = gbase, *g2 = g1 + n1, *g3 = g2 + n2;
    const __half* sg0 = static_cast<const __half*>(sg_f16);
    const __half *sg1 = sg0 + n0, *sg2 = sg1 + n1;
    // g_3 = -U^T (c_2 s_2);  g_2 = c_2 s_2 + (DG)^T g_3 - U^T (c_1 s_1);  g_1 = c_1 s_1 + (DG)^T g_2 - U^T (c_0 s_0)
    dim3 grid0(mg::ceil_div(W, 32), mg::ceil_div(H, 32), 3 * S);
    size_t n_tiles;
    uint8_t* t_raw = tile_maps(ws, S, H, W, &n_tiles);
    const uint8_t* nr = t_raw + ((n_tiles + 3) / 4) * 4;     // the `near` map left by the forward
    const int TY = (int)grid0.y, TX = (int)grid0.x;
    MG_LAUNCH(loss_bwd_small_kernel, small_grid(n3), 256, 0, stream, (const __half*)nullptr, (const float*)nullptr, sg2, coef, g3, nr,
              TY, TX, S, H / 8, W / 8, 3);
    MG_LAUNCH(loss_bwd_small_kernel, small_grid(n2), 256, 0, stream, sg2, (const float*)g3, sg1, coef, g2, nr, TY, TX, S, H / 4, W / 4, 2);
    MG_LAUNCH(loss_bwd_small_kernel, small_grid(n1), 256, 0, stream, sg1, (const float*)g2, sg0, coef, g1, nr, TY, TX, S, H / 2, W / 2, 1);
    MG_LAUNCH(loss_bwd_level0_kernel, grid0, dim3(32, 8), 0, stream, P, target, Wt, plane_scale, sg0, (const float*)g1, coef, G, nr, S, H, W);
    MG_CHECK_LAUNCH("mg_loss_bwd");
    return MG_OK;
}
