// K2: dense convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in
// TMEM), operands staged by TMA.
//
//   D[128 pixels x BN channels] = sum over (tap, Cin chunk)  A_tap[128 x BK] * W_tap[BN x BK]^T
//
// * A tile: a th x tw rectangle of output pixels (th*tw = 128).  For tap (dy,dx) the operand is simply the input
//   box shifted by (dy,dx): ONE 4-D TMA load {BK channels, tw, th, 1 image} of the NHWC tensor, with the
//   hardware zero-filling out-of-bounds coordinates (= the conv's zero padding) and `elementStrides` doing the
//   stride-2 subsampling.  The box lands in shared memory as 128 rows of BK*2 bytes with 32/64/128-byte
//   swizzle = the canonical K-major UMMA operand layout.  No im2col buffer ever exists.
// * W tile: 2-D TMA box {BK, BN} of the packed weight matrix [Cout][taps*Cin] (K contiguous).
// * 4-stage mbarrier pipeline: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2..5 =
//   epilogue (tcgen05.ld -> optional ReLU -> fp16 NHWC store, + per-channel sum / sum-of-squares partials for
//   training-mode BatchNorm).
// The same kernel serves 3x3/1x1/dilated/strided forward convs, the four sub-pixel phases of the 4x4 stride-2
// transposed conv, the 2x2 stride-2 "avg-pool + 1x1" skip conv, and every data-gradient (dgrad) by passing the
// transposed/flipped weight pack and tap table.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

#include <cstdlib>
#include <mutex>

namespace {

using namespace mg::ptx;

constexpr int BM = 128;            // pixels per tile (UMMA M)
constexpr int THREADS = 192;       // 6 warps
constexpr int MAX_TAPS = MG_CONV_MAX_TAPS;
constexpr int STAT_COPIES = MG_CONV_STAT_COPIES;

struct KArgs {
    int n_taps;
    int tap_dy[MAX_TAPS], tap_dx[MAX_TAPS], tap_koff[MAX_TAPS];
    // several launches that differ only in their tap list and output phase (the four sub-pixel phases of a stride-2 data
    // gradient / transposed conv) run as ONE grid: blockIdx.z selects the phase, taps [tap0[z], tap0[z+1]) of the table
    int n_phases, phase_tap0[5], phase_oy0[4], phase_ox0[4];
    int sy, sx, Hg, Wg, th, tw, tiles_y, tiles_x;
    int BK, kchunks, BN, Co, stages, swizzle;
    __half* out;
    int Ho, Wo, Cs, c_off, oys, oy0, oxs, ox0;
    int pre_act, post_act;           // 0 none, 1 relu, 2 leaky-relu(0.2); pre: before the affine, post: after residual
    float* stats;
    const float* bias;
    const float* scale;              // optional per-channel affine (eval-mode BatchNorm folded into the epilogue)
    const float* shift;
    const __half* res;               // optional residual, NHWC [N,Ho,Wo,Co] or [N,Ho/2,Wo/2,Co] when res_up
    int res_up;
    // split-fp16 ("x3") mode, template parameter HP: `out` and `res` are fp32 tensors (same indexing)
    float* out32;
    const float* res32;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v > 0.f ? v : 0.2f * v) : v);
}

// HP ("x3", evaluation at fp32-level accuracy): every operand is an unevaluated sum of two fp16 tensors (x = x_hi + x_lo,
// w = w_hi + w_lo, |lo| <= 2^-11 |hi|); the K loop runs three times - x_hi w_hi, x_lo w_hi, x_hi w_lo - into the same fp32
// TMEM accumulator (the dropped x_lo w_lo term is 2^-22 relative), and the epilogue reads / writes fp32 tensors.
template <bool HP>
__global__ void __launch_bounds__(THREADS, 2)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA_lo, const __grid_constant__ CUtensorMap tmB_lo, const KArgs a) {
    mg::pdl_launch();   // the next kernel may be scheduled; its own griddepcontrol.wait orders the memory accesses
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: required by the 128B swizzle pattern shared by TMA and the UMMA descriptors
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = BM * a.BK * 2, b_bytes = a.BN * a.BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = sA + a.stages * a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + a.stages * b_bytes);  // full[stages], empty[stages], tmem_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * a.stages + 1);
    float* s_stage = reinterpret_cast<float*>(tmem_slot + 4);               // [4 warps][32][17]
    float* s_part = s_stage + 4 * 32 * 17;                                  // [4 warps][2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * a.stages, tfull = empty0 + 8 * a.stages;

    // tile coordinates
    int t = blockIdx.x;
    const int tx = t % a.tiles_x;
    t /= a.tiles_x;
    const int ty = t % a.tiles_y, img = t / a.tiles_y;
    const int y0 = ty * a.th, x0 = tx * a.tw, n0 = blockIdx.y * a.BN;
    const int tap_first = a.phase_tap0[blockIdx.z];
    const int nkb1 = (a.phase_tap0[blockIdx.z + 1] - tap_first) * a.kchunks;   // k-blocks of one pass over the taps
    const int nkb = HP ? 3 * nkb1 : nkb1;
    const int oy0 = a.phase_oy0[blockIdx.z], ox0 = a.phase_ox0[blockIdx.z];

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (HP) {
            prefetch_tmap(&tmA_lo);
            prefetch_tmap(&tmB_lo);
        }
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), a.BN < 32 ? 32 : a.BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mg::pdl_wait();   // barriers, TMEM and descriptors are set up while the previous kernel drains; now wait for its data
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int s = 0, ph = 0, term = 0, tap = tap_first, c = 0;   // counters: no run-time divisions on the issue path
            const int tap_last = a.phase_tap0[blockIdx.z + 1];
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                mbar_expect_tx(full0 + 8 * s, a_bytes + b_bytes);
                tma_load_4d(smem_u32(sA + s * a_bytes), (HP && term == 1) ? &tmA_lo : &tmA, full0 + 8 * s, c * a.BK,
                            x0 * a.sx + a.tap_dx[tap], y0 * a.sy + a.tap_dy[tap], img);
                tma_load_2d(smem_u32(sB + s * b_bytes), (HP && term == 2) ? &tmB_lo : &tmB, full0 + 8 * s,
                            a.tap_koff[tap] + c * a.BK, n0);
                if (++s == a.stages) s = 0, ph ^= 1;
                if (++c == a.kchunks) {
                    c = 0;
                    if (++tap == tap_last) tap = tap_first, ++term;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop (all operands provably warp-uniform -> uniform registers, no
        // elect / R2UR waterfall around every tcgen05.mma), one elected lane issues =====
        const uint32_t tmem_u = uniform_u32(tmem_base);
        const uint32_t idesc = instr_desc_f16(BM, a.BN, 0, 0);
        const uint32_t layout = swizzle_layout(a.swizzle), sbo = 8 * a.swizzle;
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        const int ksteps = a.BK / 16;
        int s = 0, ph = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            const uint64_t da = smem_desc(sA_u + s * a_bytes, 0, sbo, layout);
            const uint64_t db = smem_desc(sB_u + s * b_bytes, 0, sbo, layout);
            if (elect_one()) {
                mma_f16(tmem_u, da, db, idesc, kb != 0);
                for (int k = 1; k < ksteps; ++k) mma_f16(tmem_u, da + 2 * k, db + 2 * k, idesc, 1u);
                mma_commit(empty0 + 8 * s);  // frees the smem slot when these MMAs retire
            }
            if (++s == a.stages) s = 0, ph ^= 1;
        }
        if (elect_one()) mma_commit(tfull);  // accumulator complete
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int m = q * 32 + lane;                  // accumulator row = pixel within the tile
        const int py = y0 + m / a.tw, px = x0 + m % a.tw;
        const bool valid = (py < a.Hg) && (px < a.Wg);
        const int oy = py * a.oys + oy0, ox = px * a.oxs + ox0;
        const size_t ooff = (((size_t)img * a.Ho + oy) * a.Wo + ox) * a.Cs + a.c_off + n0;
        const size_t roff = (a.res_up ? (((size_t)img * (a.Ho >> 1) + (oy >> 1)) * (a.Wo >> 1) + (ox >> 1))
                                      : (((size_t)img * a.Ho + oy) * a.Wo + ox)) * a.Co + n0;
        __half* orow = HP ? nullptr : a.out + ooff;
        const __half* rrow = (!HP && a.res) ? a.res + roff : nullptr;
        float* orow32 = HP ? a.out32 + ooff : nullptr;
        const float* rrow32 = (HP && a.res32) ? a.res32 + roff : nullptr;
        float* stg = s_stage + q * 32 * 17;
        float* part = s_part + q * 2 * a.BN;
        mbar_wait(tfull, 0);
        tc_fence_after();
        for (int c0 = 0; c0 < a.BN; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(r[i]);
                if (a.bias) v[i] += __ldg(a.bias + n0 + c0 + i);
                v[i] = apply_act(v[i], a.pre_act);
            }
            if (a.stats) {
                // per-channel sum / sum of squares over this warp's 32 rows (rows outside the image are masked: a bias
                // would make them non-zero)
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 16; ++i) stg[lane * 17 + i] = valid ? v[i] : 0.f;
                __syncwarp();
                const int col = lane & 15;
                float acc = 0.f;
                if (lane < 16) {
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) acc += stg[rr * 17 + col];
                } else {
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) { const float z = stg[rr * 17 + col]; acc += z * z; }
                }
                part[(lane >> 4) * a.BN + c0 + col] = acc;
            }
            if (valid && n0 + c0 < a.Co) {
                if (a.scale) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], __ldg(a.scale + n0 + c0 + i), __ldg(a.shift + n0 + c0 + i));
                }
                if (HP && rrow32) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 f = __ldg(reinterpret_cast<const float4*>(rrow32 + c0) + i);
                        v[4 * i] += f.x, v[4 * i + 1] += f.y, v[4 * i + 2] += f.z, v[4 * i + 3] += f.w;
                    }
                }
                if (!HP && rrow) {
                    const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rrow + c0)), r1 = __ldg(reinterpret_cast<const uint4*>(rrow + c0) + 1);
                    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[i]));
                        v[2 * i] += f.x, v[2 * i + 1] += f.y;
                    }
                }
                if (a.post_act) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], a.post_act);
                }
                if (HP) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        reinterpret_cast<float4*>(orow32 + c0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    continue;
                }
                uint4 o0, o1;
                __half2 h;
                h = __floats2half2_rn(v[0], v[1]);   o0.x = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(v[2], v[3]);   o0.y = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(v[4], v[5]);   o0.z = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(v[6], v[7]);   o0.w = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(v[8], v[9]);   o1.x = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(v[10], v[11]); o1.y = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(v[12], v[13]); o1.z = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(v[14], v[15]); o1.w = *reinterpret_cast<uint32_t*>(&h);
                reinterpret_cast<uint4*>(orow + c0)[0] = o0;
                reinterpret_cast<uint4*>(orow + c0)[1] = o1;
            }
        }
        if (a.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
            const int et = threadIdx.x - 64;                 // 0..127
            float* dst = a.stats + (size_t)(blockIdx.x % STAT_COPIES) * 2 * a.Co;
            for (int i = et; i < 2 * a.BN; i += 128) {
                const int kind = i / a.BN, c = i - kind * a.BN;
                if (n0 + c < a.Co) {
                    const float tot = s_part[0 * 2 * a.BN + i] + s_part[1 * 2 * a.BN + i] + s_part[2 * 2 * a.BN + i] +
                                      s_part[3 * 2 * a.BN + i];
                    atomicAdd(dst + kind * a.Co + n0 + c, tot);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.BN < 32 ? 32 : a.BN);
    }
}

}  // namespace

namespace mg {
int conv_halo_launch(const mg_conv_desc* d, void* stream, bool* handled);     // k2b_conv_halo.cu
int conv_mid_launch(const mg_conv_desc* d, void* stream, bool* handled);      // k2h_conv_mid.cu
int conv_midt_launch(const mg_conv_desc* d, void* stream, bool* handled);     // k2t_conv_mid_t.cu
}

static int conv_launch_impl(const mg_conv_desc* d, const void* x_lo, const void* w_lo, void* stream) {
    const bool hp = x_lo != nullptr;
    MG_REQUIRE(d && d->x && d->w && d->out, "mg_conv_fprop: null pointer");
    MG_REQUIRE(d->n_taps >= 1 && d->n_taps <= MAX_TAPS, "mg_conv_fprop: n_taps %d out of range", d->n_taps);
    MG_REQUIRE(d->Ci % 16 == 0 && d->Co % 16 == 0, "mg_conv_fprop: Ci (%d) and Co (%d) must be multiples of 16", d->Ci, d->Co);
    MG_REQUIRE(d->Ktot % 8 == 0 && d->Cs % 8 == 0 && d->c_off % 8 == 0, "mg_conv_fprop: Ktot/Cs/c_off must be multiples of 8");
    MG_REQUIRE(d->sy >= 1 && d->sx >= 1 && d->sy <= 2 && d->sx <= 2, "mg_conv_fprop: stride must be 1 or 2");
    MG_REQUIRE(d->N > 0 && d->Hg > 0 && d->Wg > 0, "mg_conv_fprop: empty problem");
    mg::EncodeTiledFn enc = mg::get_encode();
    if (!enc) {
        mg::set_error("mg_conv_fprop: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return MG_ERR_CUDA;
    }
    MG_REQUIRE((d->scale == nullptr) == (d->shift == nullptr), "mg_conv_fprop: scale and shift go together");
    if (!hp) {
        // high-resolution, low-channel stride-1 layers: halo-resident persistent kernel (K2b)
        bool handled = false;
        const int rc = mg::conv_halo_launch(d, stream, &handled);
        if (rc != MG_OK || handled) return rc;
    }
    if (!hp) {
        // the same layers with Co a multiple of 128: transposed form (weights as M, patch pixels as N <= 256), K2t
        bool handled = false;
        const int rc = mg::conv_midt_launch(d, stream, &handled);
        if (rc != MG_OK || handled) return rc;
    }
    if (!hp) {
        // mid-resolution stride-1 3x3 layers with >= 128 input channels: halo-patch / weight-streaming kernel (K2h)
        bool handled = false;
        const int rc = mg::conv_mid_launch(d, stream, &handled);
        if (rc != MG_OK || handled) return rc;
    }
    KArgs a;
    a.n_taps = d->n_taps;
    for (int t = 0; t < d->n_taps; ++t) a.tap_dy[t] = d->tap_dy[t], a.tap_dx[t] = d->tap_dx[t], a.tap_koff[t] = d->tap_koff[t];
    a.sy = d->sy, a.sx = d->sx, a.Hg = d->Hg, a.Wg = d->Wg;
    if (d->Wg > 8) a.th = 8, a.tw = 16; else a.th = 16, a.tw = 8;
    a.tiles_y = mg::ceil_div(d->Hg, a.th), a.tiles_x = mg::ceil_div(d->Wg, a.tw);
    a.BK = (d->Ci % 64 == 0) ? 64 : ((d->Ci % 32 == 0) ? 32 : 16);
    a.kchunks = d->Ci / a.BK;
    a.swizzle = a.BK * 2;
    a.Co = d->Co;
    a.BN = d->Co >= 128 ? 128 : (d->Co >= 64 ? 64 : 32);
    MG_REQUIRE(d->Co % a.BN == 0 || d->Co < a.BN, "mg_conv_fprop: Co=%d not a multiple of the N tile %d", d->Co, a.BN);
    if (d->Co < a.BN) a.BN = d->Co;  // 16 (TMEM allocation is rounded up to 32 columns)
    a.out = static_cast<__half*>(d->out);
    a.Ho = d->Ho, a.Wo = d->Wo, a.Cs = d->Cs, a.c_off = d->c_off;
    a.oys = d->oys, a.oy0 = d->oy0, a.oxs = d->oxs, a.ox0 = d->ox0;
    a.n_phases = 1;
    a.phase_tap0[0] = 0, a.phase_tap0[1] = d->n_taps, a.phase_oy0[0] = d->oy0, a.phase_ox0[0] = d->ox0;
    if (d->n_phases > 1) {
        MG_REQUIRE(d->n_phases <= 4, "mg_conv_fprop: at most 4 phases");
        a.n_phases = d->n_phases;
        for (int p = 0; p < d->n_phases; ++p) {
            a.phase_tap0[p] = d->phase_tap0[p], a.phase_oy0[p] = d->phase_oy0[p], a.phase_ox0[p] = d->phase_ox0[p];
            MG_REQUIRE(d->phase_tap0[p + 1] > d->phase_tap0[p], "mg_conv_fprop: every phase needs at least one tap");
        }
        a.phase_tap0[d->n_phases] = d->phase_tap0[d->n_phases];
        MG_REQUIRE(a.phase_tap0[0] == 0 && a.phase_tap0[d->n_phases] == d->n_taps, "mg_conv_fprop: phase tap ranges must cover the tap table");
    }
    a.pre_act = d->pre_act, a.post_act = d->post_act, a.stats = d->stats, a.bias = d->bias;
    a.scale = d->scale, a.shift = d->shift, a.res = static_cast<const __half*>(d->res), a.res_up = d->res_up;
    a.out32 = static_cast<float*>(d->out), a.res32 = static_cast<const float*>(d->res);
    MG_REQUIRE(!d->res || (d->c_off == 0 && d->Cs == d->Co), "mg_conv_fprop: residual needs a dense [N,Ho,Wo,Co] output");

    const int a_bytes = BM * a.BK * 2, b_bytes = a.BN * a.BK * 2;
    const int fixed = 1024 /*align*/ + 256 /*barriers*/ + 4 * 32 * 17 * 4 + 4 * 2 * a.BN * 4;
    // this non-persistent kernel hides its prologue / epilogue behind the main loop of co-resident CTAs: keep the
    // per-CTA footprint at <= 111 KB so that two CTAs (TMEM: 2 x BN <= 512 columns) share an SM and the 128 x 128 x 64
    // layers still get a three-stage ring (measured on the C2 step: 530.6 vs 511.2 frames/s with a 100 KB budget = two
    // stages; MAGGIE_B200_CONV_SMEM_KB overrides, for experiments)
    static const int budget_kb = [] { const char* e = std::getenv("MAGGIE_B200_CONV_SMEM_KB"); const int v = e ? std::atoi(e) : 0; return v >= 64 && v <= 220 ? v : 111; }();
    a.stages = std::max(2, std::min(4, (budget_kb * 1024 - fixed) / (a_bytes + b_bytes)));
    const size_t smem = (size_t)fixed + (size_t)a.stages * (a_bytes + b_bytes);

    CUtensorMap tmA, tmB, tmA_lo, tmB_lo;
    auto encode_a = [&](CUtensorMap* tm, const void* ptr) -> int {
        cuuint64_t dims[4] = {(cuuint64_t)d->Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->Ci * 2, (cuuint64_t)d->Wi * d->Ci * 2, (cuuint64_t)d->Hi * d->Wi * d->Ci * 2};
        cuuint32_t box[4] = {(cuuint32_t)a.BK, (cuuint32_t)(a.tw * a.sx), (cuuint32_t)(a.th * a.sy), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)a.sx, (cuuint32_t)a.sy, 1};
        CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, mg::swz_enum(a.swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            mg::set_error("mg_conv_fprop: cuTensorMapEncodeTiled(A) failed (%d): x=%p N=%d H=%d W=%d C=%d box=%u,%u,%u", (int)r,
                          ptr, d->N, d->Hi, d->Wi, d->Ci, box[0], box[1], box[2]);
            return MG_ERR_CUDA;
        }
        return MG_OK;
    };
    auto encode_b = [&](CUtensorMap* tm, const void* ptr) -> int {
        cuuint64_t dims[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Co};
        cuuint64_t strides[1] = {(cuuint64_t)d->Ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)a.BK, (cuuint32_t)a.BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, mg::swz_enum(a.swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            mg::set_error("mg_conv_fprop: cuTensorMapEncodeTiled(W) failed (%d): Ktot=%d Co=%d", (int)r, d->Ktot, d->Co);
            return MG_ERR_CUDA;
        }
        return MG_OK;
    };
    if (int e = encode_a(&tmA, d->x)) return e;
    if (int e = encode_b(&tmB, d->w)) return e;
    if (hp) {
        if (int e = encode_a(&tmA_lo, x_lo)) return e;
        if (int e = encode_b(&tmB_lo, w_lo)) return e;
    } else {
        tmA_lo = tmA, tmB_lo = tmB;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
            mg::set_error("mg_conv_fprop: cannot raise dynamic shared memory limit");
            return MG_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid(d->N * a.tiles_y * a.tiles_x, mg::ceil_div(d->Co, a.BN), a.n_phases);
    if (hp)
        MG_LAUNCH(conv_tcgen05_kernel<true>, grid, THREADS, smem, stream, tmA, tmB, tmA_lo, tmB_lo, a);
    else
        MG_LAUNCH(conv_tcgen05_kernel<false>, grid, THREADS, smem, stream, tmA, tmB, tmA_lo, tmB_lo, a);
    MG_CHECK_LAUNCH("mg_conv_fprop");
    return MG_OK;
}

extern "C" int mg_conv_fprop(const mg_conv_desc* d, void* stream) { return conv_launch_impl(d, nullptr, nullptr, stream); }

extern "C" int mg_conv_fprop_x3(const mg_conv_desc* d, const void* x_lo, const void* w_lo, void* stream) {
    MG_REQUIRE(x_lo && w_lo, "mg_conv_fprop_x3: null pointer");
    MG_REQUIRE(d && !d->stats, "mg_conv_fprop_x3: evaluation only (no BatchNorm statistics epilogue)");
    return conv_launch_impl(d, x_lo, w_lo, stream);
}
