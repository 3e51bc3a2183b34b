/*
 * maggie_b200.h - C ABI of libmaggie_b200.so: hand-written sm_100a CUDA kernels for the MaGGIe
 * forward/backward hot path (encoder/ASPP convs, mask-guided attention decoder, sparse refinement).
 *
 * The reference (hmchuong/MaGGIe) has NO FFI: its native work is library calls made from Python
 * (cuDNN/ATen through torch, spconv, cv2.dilate).  Each entry point below therefore replaces a Python
 * call site of the reference; the citation after "replaces:" is relative to /root/reference/maggie.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; all buffers (inputs, outputs,
 *     workspaces) are owned and allocated by the caller (PyTorch) and only borrowed for the call;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no internal synchronisation,
 *     no allocation: safe for CUDA-graph capture;
 *   - return value 0 = OK, non-zero = error (message via mg_last_error(), thread local); the library never
 *     exits or throws across the ABI;
 *   - activations are NHWC fp16 unless stated, statistics / logits / alphas fp32, index data int32.
 */
#ifndef MAGGIE_B200_H
#define MAGGIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_OK 0
#define MG_ERR_ARG 1
#define MG_ERR_CUDA 2
#define MG_ERR_UNSUPPORTED 3

/* ---- library ---- */
int mg_version(void);
const char* mg_last_error(void);
/* number of kernels this library has launched since load / last reset (bench.py's "gpu_launches") */
unsigned long long mg_launch_count(void);
void mg_reset_launch_count(void);

/* ---- K8a: uncertainty ("unknown") mask -------------------------------------------------------------
 * replaces: utils/utils.py:28-55 compute_unknown (threshold 1/255 < a < 254/255, then per-slice
 *           cv2.dilate with MORPH_ELLIPSE(width) on the CPU, incl. the D2H/H2D round trip).
 * alpha     [slices,H,W] fp32.          widths [slices] int32, ellipse size 1..29 per slice.
 * and_mask  optional [slices,H,W] uint8: result is AND-ed with (and_mask != 0)   (fuse(): "* detail_mask",
 *           decoder/resnet_inst_matt_spconv.py:279,285,335-336).
 * out_u8    optional [slices,H,W] uint8 in {0,1}.   out_bits optional [slices,H,ceil(W/32)] uint32, bit x%32
 *           of word x/32.  At least one output must be given.  Bit-exact.                              */
int mg_unknown_mask(const float* alpha, int slices, int H, int W, const int32_t* widths,
                    const uint8_t* and_mask, uint8_t* out_u8, uint32_t* out_bits, void* stream);
/* the same, reading the slices from `alt` instead of `alpha` when the DEVICE flag *use_alt != 0: the warm-up switch of
 * decoder/resnet_inst_matt_spconv.py:311-316 ("guide with the ground truth when the predicted alpha is all zero") decided
 * on the device, no host read. */
int mg_unknown_mask_select(const float* alpha, const float* alt, const int32_t* use_alt, int slices, int H, int W,
                           const int32_t* widths, const uint8_t* and_mask, uint8_t* out_u8, uint32_t* out_bits,
                           void* stream);

/* ---- K10: one stage of the progressive fusion -----------------------------------------------------------------
 * replaces: fuse() of decoder/resnet_inst_matt_spconv.py:272-290 (compute_unknown on the CPU + two full-size blend passes
 *           per stage): w = dilate(1/255 < src < 254/255, ellipse(width)) AND and_mask;  out_alpha = w ? finer : coarser
 *           (`finer * w + coarser * (1 - w)` with w in {0,1}).  src / finer / coarser / out_alpha fp32 [slices,H,W]
 *           (out_alpha may not alias src: neighbouring CTAs still read its halo rows), out_w_u8 the {0,1} mask.  Bit-exact. */
int mg_fuse_stage(const float* src, const float* finer, const float* coarser, int slices, int H, int W,
                  const int32_t* widths, const uint8_t* and_mask, uint8_t* out_w_u8, float* out_alpha, void* stream);

/* ---- K18: transition ("trimap") ground truth of the loaders ---------------------------------------------------
 * replaces: gen_transition_gt (dataloader/utils.py:15-35, called by dataloader/him.py:187-193 and vim.py:199: cv2.dilate /
 *           cv2.erode of every instance alpha with MORPH_ELLIPSE (k, k), `iterations` times, on the CPU):
 *           trans = (dilate(alpha) - erode(alpha) > 0) | ((alpha > 127) != (mask == 255)), uint8 {0,1}.
 * alpha_u8 [planes, H, W]; mask_u8 NULL or [planes, H / mask_div, W / mask_div] (mask_div 1 or 8: the loader's down-scaled
 * masks, repeated 8 x 8 as `torch.repeat_interleave` does); tmp_u8: 4 * planes * H * W bytes of scratch (may be NULL for one
 * iteration); k_size 1..31.  Bit-exact against OpenCV (tests/test_io_host.py pins the oracle, tests/test_gpu_io.py the kernel). */
int mg_transition_gt(const void* alpha_u8, const void* mask_u8, int mask_div, int planes, int H, int W, int k_size,
                     int iterations, void* tmp_u8, void* trans_u8, void* stream);

/* ---- K11: element-wise half of the video model (MaGGIe_Temp) --------------------------------------------------
 * replaces: the gating arithmetic of ConvGRU.forward_single_frame (module/conv_gru.py:50-58: sigmoid / split / r*h /
 *           torch.cat / tanh / blend, ~12 torch launches per step and direction) around the two native convolutions, and
 *           the forward / backward alpha recurrences of bidirectional_fusion
 *           (decoder/resnet_inst_matt_spconv_temp.py:122-142: 4 (n_f - 1) full-resolution blend passes).
 * GRU tensors are NHWC fp16 rows [P = N*h*w][channels]; C = hidden channels (multiple of 8).
 *   cat1 = [x | h] ([P][2C]),  rz = conv_ih(cat1) + bias ([P][2C], r then z),  cat2 = [x | sigmoid(r) * h],
 *   c_pre = conv_hh(cat2) + bias ([P][C]),  h_new = (1 - z) h + z tanh(c_pre).
 * Backward: gate2_bwd turns dh_new into drz[:, C:], dc_pre and dh_direct = dh_new (1 - z); gate1_bwd turns
 *   dcat2 (data gradient of conv_hh) into drz[:, :C] and dpart = [dcat2[:, :C] | dh_direct + dcat2[:, C:] sigmoid(r)], which
 *   the data gradient of conv_ih takes as its residual input: its output is [dx | dh]. */
int mg_gru_concat2(const void* x, const void* h, void* cat1, size_t P, int C, void* stream);
int mg_gru_gate1_fwd(const void* rz, const void* cat1, void* cat2, size_t P, int C, void* stream);
int mg_gru_gate2_fwd(const void* rz, const void* c_pre, const void* cat1, void* h_new, size_t P, int C, void* stream);
int mg_gru_gate2_bwd(const void* dh_new, const void* rz, const void* c_pre, const void* cat1, void* drz, void* dc_pre,
                     void* dh_direct, size_t P, int C, void* stream);
int mg_gru_gate1_bwd(const void* dcat2, const void* rz, const void* cat1, const void* dh_direct, void* drz, void* dpart,
                     size_t P, int C, void* stream);
/* fd / bd: fp32 logits [B][F][HW] of the forward / backward temporal-difference maps (plane 0 of fd and plane F-1 of bd
 * are not read); preds / fused fp32 [B][F][n_i][HW]:  fp_0 = p_0, fp_i = fp_{i-1} (1 - s(fd_i)) + p_i s(fd_i);
 * bp_{F-1} = p_{F-1}, bp_i = bp_{i+1} (1 - s(bd_i)) + p_i s(bd_i);  fused = [fp_0, (fp_i + bp_i) / 2 ..., bp_{F-1}].
 * 2 <= F <= 8.  The backward recomputes both recurrences and returns the gradients w.r.t. preds and the logits. */
int mg_temporal_fuse_fwd(const float* fd, const float* bd, const float* preds, float* fused, int B, int F, int n_i,
                         size_t HW, void* stream);
int mg_temporal_fuse_bwd(const float* fd, const float* bd, const float* preds, const float* dfused, float* dpreds,
                         float* dfd, float* dbd, int B, int F, int n_i, size_t HW, void* stream);

/* ---- K8b: active-site lists for the sparse refinement ---------------------------------------------
 * replaces: decoder/resnet_inst_matt_spconv.py:203-218 (torch.nonzero + spconv `dummy_downscale`, whose only
 *           live product is the OS1/OS2/OS4/OS8 index sets and the (in,out,tap) pair tables).
 * Level l (0..3) has spatial size (H>>l, W>>l) (H, W multiples of 8).  Level l+1 site q is active iff any
 * level-l site lies in rows/cols 2q-1..2q+1 (SparseConv2d k3 s2 p1).  Sites of a level are numbered in
 * lexicographic (slot,y,x) order == torch.nonzero order.
 *
 * mg_sites_workspace : bytes needed for `ws`.
 * mg_sites_count     : builds the bit pyramids + rank structure in `ws`, writes counts[4] (device int32).
 * mg_sites_tables    : after the caller has read counts[] (one 16-byte D2H) and allocated exact-size outputs:
 *   coords[l]  [N_l,3] int32 (slot,y,x)                                     (l = 0..3, any may be NULL)
 *   nbr[l]     [N_l,9] int32  row of the active neighbour at tap (ky,kx) = (y+ky-1, x+kx-1), else -1
 *              (SubMConv2d 3x3 rulebook; requested for l = 0 and l = 2)
 *   parent[l]  [N_l,9] int32  for l = 0..2: row (at level l+1) of q with p = 2q-1+k for tap k, else -1
 *              (SparseInverseConv2d rulebook, forward)
 *   child[l]   [N_l,9] int32  for l = 1..3: row (at level l-1) of p = 2q-1+k if active, else -1
 *              (SparseInverseConv2d backward / SparseConv2d forward rulebook)                            */
size_t mg_sites_workspace(int slots, int H, int W);
int mg_sites_count(const uint8_t* roi, int slots, int H, int W, void* ws, int32_t* counts, void* stream);
int mg_sites_tables(const void* ws, int slots, int H, int W, const int32_t* counts_host,
                    int32_t* const* coords, int32_t* const* nbr, int32_t* const* parent,
                    int32_t* const* child, void* stream);

/* ---- K1: mask-id embedding + NHWC pack -------------------------------------------------------------
 * replaces: arch/maggie.py:200-235 (zero-padded 10-slot mask tensor + cat) and
 *           encoder/resnet.py:211-229 (Embedding gather, masked mean over instances, permute, cat).
 * image [B,3,H,W] fp32 NCHW; masks [B,M,H,W] fp32 {0,1}; slot_ids[M] (DEVICE int32) = slot (0..9) of each given mask;
 * table [11,3] fp32; out [B,H,W,C] fp16 NHWC, C >= 6: ch 0-2 image, 3-5 mean embedding, rest zero.     */
int mg_mask_embed_fwd(const float* image, const float* masks, const int32_t* slot_ids, int M,
                      const float* table, void* out_f16, int B, int H, int W, int C, void* stream);
/* the same with an fp32 NHWC output (evaluation at fp32-level accuracy)                                  */
int mg_mask_embed_fwd_f32(const float* image, const float* masks, const int32_t* slot_ids, int M,
                          const float* table, float* out_f32, int B, int H, int W, int C, void* stream);
/* grad_table [11,3] fp32 += d(out[...,3:6]) / d(table)   (caller zeroes grad_table)                     */
int mg_mask_embed_bwd(const void* grad_out_f16, const float* masks, const int32_t* slot_ids, int M,
                      float* grad_table, int B, int H, int W, int C, void* stream);


/* ---- K0: grouped weight preparation (spectral norm + fp16 operand packs) and its backward -------------
 * replaces: module/spectral_norm.py:22-35,73-80 (SpectralNorm._update_u_v + forward, ~12 tiny launches per layer x 54
 *           layers), the implicit fp32->fp16 weight casts of autocast, and autograd's backward through W_bar / sigma.
 * One call prepares EVERY dense conv layer: `layers` is a DEVICE array of descriptors, items_* are DEVICE int32[4]
 * work-item tables built by the host ((layer, row0, col0, first)): vt tiles 64 rows x 256 columns, u tiles 8 rows,
 * pack tiles 16 (dim0) x 32 (dim1) with dim1 (conv) / dim0 (transposed conv) running up to ci_pad.
 * mg_wprep_fwd : per layer  v <- norm(W^T u), u <- norm(W v) (in place), sigma = u.Wv;  P[p_off..] = fp16 W/sigma as
 *                [Co][taps_out][ci_pad];  D[d_off..] = fp16 W/sigma as [ci_pad][taps_out][Co] (D may be NULL).
 *                vec: scratch (zeroed by the caller), per layer v_raw[width] then t[dim0];  scal: [n_layers][4] =
 *                {sigma, 1/(|v_raw|+eps), 1/(|t|+eps), <G,W_bar> accumulator}.  fold: the stored 1x1 weight is emitted
 *                as 4 taps of 0.25 W (AvgPool2d(2) + 1x1 conv == 2x2 stride-2 conv).  u == NULL: plain conv, sigma = 1.
 * mg_wprep_bwd : G (fp32, P layout, at g_off) -> grad[grad_off..] = dL/dW_bar in the torch layout of `w`.             */
typedef struct mg_wprep_layer {
    const float* w;            /* master weight, torch layout [dim0][dim1][kh][kw]: conv dim0 = Co, transposed conv dim0 = Ci */
    float* u; float* v;        /* spectral-norm vectors [dim0], [dim1*kh*kw] (updated in place) or NULL */
    int32_t Co, Ci, taps, transposed, fold, ci_pad;
    int32_t vec_off, pad_;
    int64_t p_off, d_off, g_off, grad_off;
} mg_wprep_layer;
int mg_wprep_fwd(const mg_wprep_layer* layers, const int32_t* items_vt, int n_vt, const int32_t* items_u, int n_u,
                 const int32_t* items_tile, int n_tile, float* vec, float* scal, void* P_f16, void* D_f16, void* stream);
int mg_wprep_bwd(const mg_wprep_layer* layers, const int32_t* items_tile, int n_tile, const float* G, const float* vec,
                 float* scal, float* grad, void* stream);

/* ---- K2: dense convolution, implicit GEMM on tcgen05 tensor cores with TMA-staged operands ---------
 * replaces: every nn.Conv2d / nn.ConvTranspose2d call of the hot path (cuDNN via ATen): encoder
 *           encoder/resnet.py:177-200, ASPP module/aspp.py:35-56, decoder decoder/resnet.py:33-45,
 *           module/instance_matte_decoder.py:81-88; and, with the transposed weight pack, their data gradients.
 * x    NHWC fp16 [N,Hi,Wi,Ci], Ci % 16 == 0.      w  packed fp16 [Co][Ktot] (K contiguous), Co % 16 == 0.
 * The op computes, for every image n and logical grid point (y,x), 0<=y<Hg, 0<=x<Wg:
 *     acc[c] = sum_t sum_ci  x[n, y*sy + tap_dy[t], x*sx + tap_dx[t], ci] * w[c][tap_koff[t] + ci]   (zero outside x)
 *     v = pre_act(acc[c] + bias[c]);  [stats see v];  v = v*scale[c] + shift[c];  v += res[...];  v = post_act(v)
 *     out[n, y*oys + oy0, x*oxs + ox0, c_off + c] = fp16(v)
 * out is NHWC fp16 [N,Ho,Wo,Cs].  3x3 pad 1: taps (ky-1,kx-1); dilation d: d*(k-1); stride 2: sy=sx=2;
 * 4x4 stride-2 transposed conv: four launches (one per output parity) with oys=oxs=2.
 * stats (optional): float [MG_CONV_STAT_COPIES][2][Co], must be zeroed by the caller; the kernel accumulates
 * per-channel sum (row 0) and sum of squares (row 1) of the stored values' fp32 pre-images, spread over the
 * copies to keep atomics uncontended (mg_bn_finalize adds the copies up).                                  */
#define MG_CONV_MAX_TAPS 16
#define MG_CONV_STAT_COPIES 16
typedef struct mg_conv_desc {
    const void* x; int32_t N, Hi, Wi, Ci;
    const void* w; int32_t Co, Ktot;
    int32_t n_taps; int32_t tap_dy[MG_CONV_MAX_TAPS], tap_dx[MG_CONV_MAX_TAPS], tap_koff[MG_CONV_MAX_TAPS];
    int32_t sy, sx, Hg, Wg;
    void* out; int32_t Ho, Wo, Cs, c_off, oys, oy0, oxs, ox0;
    int32_t pre_act, post_act;   /* 0 none, 1 ReLU, 2 LeakyReLU(0.2) */
    float* stats;
    const float* bias;
    const float* scale; const float* shift;   /* optional per-channel affine (eval-mode BN folded in) */
    const void* res; int32_t res_up;          /* optional fp16 residual [N,Ho,Wo,Co] ([N,Ho/2,Wo/2,Co] if res_up) */
    /* n_phases in 2..4: ONE launch computes several output phases that differ only in their taps and (oy0, ox0) - the
     * sub-pixel phases of a stride-2 data gradient or of the 4x4 stride-2 transposed conv.  Phase p uses the taps
     * [phase_tap0[p], phase_tap0[p+1]) of the table and the offsets phase_oy0[p], phase_ox0[p]; 0 or 1: plain launch. */
    int32_t n_phases, phase_tap0[5], phase_oy0[4], phase_ox0[4];
} mg_conv_desc;
int mg_conv_fprop(const mg_conv_desc* desc, void* stream);
/* Evaluation at fp32-level accuracy ("x3" mode; the reference evaluates in fp32: engine/test.py:131 runs without autocast).
 * Every operand is an unevaluated sum of two fp16 tensors, x = desc->x + x_lo and w = desc->w + w_lo (mg_split_f32 below),
 * and the tensor cores accumulate x_hi*w_hi + x_lo*w_hi + x_hi*w_lo into one fp32 TMEM accumulator (the dropped x_lo*w_lo
 * term is 2^-22 relative).  desc->out and desc->res are FP32 tensors here (same shapes / indexing); desc->stats must be NULL;
 * always the generic kernel (no K2b / K2s routing).  Everything else as in mg_conv_fprop. */
int mg_conv_fprop_x3(const mg_conv_desc* desc, const void* x_lo, const void* w_lo, void* stream);
/* K17: hi = fp16(x), lo = fp16(x - hi) for n fp32 elements (n % 4 == 0). */
int mg_split_f32(const float* x, void* hi_f16, void* lo_f16, size_t n, void* stream);
/* Stride-1 layers with Ci, Co <= 64 and taps within +-1 pixel (the 512^2 .. 128^2 3x3 convolutions and their data
 * gradients) are routed by mg_conv_fprop to K2b, a persistent kernel that keeps a halo patch of the activations and all
 * the weights resident in shared memory (csrc/k2b_conv_halo.cu; MAGGIE_B200_NO_HALO_CONV=1 disables the routing).
 * mg_conv_halo_launches: how many launches took that path (tests / profiling). */
unsigned long long mg_conv_halo_launches(void);
/* Stride-1 layers with taps within +-1 pixel, Ci a multiple of 64 (>= 128), Co a multiple of 64 and width <= 64 (the 3x3
 * convolutions of the trunk at 64^2 / 32^2 / 16^2, of the dense decoder, and their data gradients) are routed to K2h, a
 * persistent kernel that keeps a halo patch of a whole row slab in shared memory, streams the weights through a ring and
 * accumulates up to five 128-pixel blocks per weight block (csrc/k2h_conv_mid.cu).  OPT-IN since K2t (MAGGIE_B200_MID_CONV=h):
 * measured slower than the generic kernel on every layer once K2 issued its MMAs warp-uniformly. */
unsigned long long mg_conv_mid_launches(void);
/* K2t (csrc/k2t_conv_mid_t.cu): the K2h layers whose Co is a multiple of 128 run in the TRANSPOSED form - weights as the
 * M = 128 operand, the linear halo-patch pixels as the N operand (N = 128..256 per tcgen05.mma, the range where the
 * instruction is paced by the tensor pipe instead of its shared-memory operand reads), BatchNorm sums as per-thread
 * accumulations, one TMA store per item.  MAGGIE_B200_NO_MIDT_CONV=1 sends them back to K2h. */
unsigned long long mg_conv_midt_launches(void);
/* profiling aid: K2t launches write 8 globaltimer stamps per CTA into buf (device, >= 148 * 8 uint64; NULL = off) */
void mg_conv_midt_trace(unsigned long long* buf);
/* profiling aid: K2h launches write 8 globaltimer stamps per CTA into buf (device, >= 148 * 8 uint64; NULL = off) */
void mg_conv_mid_trace(unsigned long long* buf);

/* ---- K4: convolution weight gradient (tcgen05, split-K over pixels) ---------------------------------
 * replaces: cuDNN's wgrad behind `loss.backward()` for every conv above (engine/train.py:266).
 * dw[co][tap_koff[t] + ci] += sum_{n,y,x} dy[n, y*ays + ay0, x*axs + ax0, co] * x[n, y*sy + tap_dy[t], x*sx + tap_dx[t], ci]
 * over the logical grid 0<=y<Hg, 0<=x<Wg (zero outside either tensor).  dy NHWC fp16 [N,Hy,Wy,Co]; x NHWC fp16
 * [N,Hi,Wi,Ci] (Ci % 16 == 0, Co % 8 == 0); dw fp32 [Co][Ktot], accumulated with atomics (caller zeroes it).  */
typedef struct mg_wgrad_desc {
    const void* dy; int32_t N, Hy, Wy, Co;
    const void* x; int32_t Hi, Wi, Ci;
    float* dw; int32_t Ktot;
    int32_t n_taps; int32_t tap_dy[MG_CONV_MAX_TAPS], tap_dx[MG_CONV_MAX_TAPS], tap_koff[MG_CONV_MAX_TAPS];
    int32_t sy, sx, ays, ay0, axs, ax0, Hg, Wg;
} mg_wgrad_desc;
int mg_conv_wgrad(const mg_wgrad_desc* desc, void* stream);
/* 3x3 stride-1 layers with 32 input channels at widths that are multiples of 128 are routed to K4b
 * (csrc/k4b_wgrad_halo.cu); mg_wgrad_halo_launches counts those launches (tests / profiling). */
unsigned long long mg_wgrad_halo_launches(void);

/* ---- K3: BatchNorm pieces around the conv kernel (NHWC fp16 activations, fp32 statistics) ----------
 * replaces: nn.BatchNorm2d forward/backward (cuDNN via ATen) + the separate ReLU/LeakyReLU/add passes of
 *           encoder/resnet.py:23-39, decoder/resnet.py:29-45, module/aspp.py:35-56.
 * mg_bn_finalize : stats = conv-epilogue copies [MG_CONV_STAT_COPIES][2][C] (training; also updates the running
 *                  statistics with `momentum`, unbiased variance) or NULL (eval: use running statistics);
 *                  writes scale = gamma*invstd, shift = beta - mean*scale, and (optionally) mean / invstd.
 * mg_bn_apply    : y = act(x*scale + shift + res)      act: 0 none, 1 ReLU, 2 LeakyReLU(0.2); res optional,
 *                  res_up: residual is [N,H/2,W/2,C] and replicated 2x2 (nearest upsampling).
 * mg_bn_bwd_reduce / mg_bn_bwd_apply : dz = dy*act'(y); sums = [sum dz ; sum dz*xhat] (caller zeroes sums [2][C]);
 *                  dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)) (* pre_act'(conv_out) if pre_act),
 *                  dres = dz (optional).  dgamma = sums[1], dbeta = sums[0].
 * count_dev (mg_bn_finalize, mg_bn_bwd_apply; may be NULL): device scalar that replaces the host-side element count.
 *                  SyncBatchNorm-equivalent training (engine/train.py:160-161): `stats` / `sums` then hold ONE copy
 *                  [2][C] of sums taken over all ranks (mg_stats_exchange, K15) and count_dev the global element count,
 *                  so that ranks with different numbers of active sites need no host synchronisation.             */
int mg_bn_finalize(const float* stats, float count, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float momentum, float eps, float* scale, float* shift, float* save_mean,
                   float* save_invstd, int C, const float* count_dev, void* stream);
int mg_bn_apply(const void* x, const float* scale, const float* shift, const void* res, int res_up, void* y, int N,
                int H, int W, int C, int act, void* stream);
/* Training forward in one launch (local statistics): sums the MG_CONV_STAT_COPIES copies, derives scale / shift, updates the
 * running statistics (optional) and applies y = act(x*scale + shift (+ res)); out4 = float [4][C]: scale, shift, mean, 1/std
 * (saved for the backward).  Same arithmetic as mg_bn_finalize followed by mg_bn_apply.  C % 8 == 0, C <= 512.            */
int mg_bn_train_apply(const float* stats, float count, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float momentum, float eps, float* out4, const void* x, const void* res, int res_up,
                      void* y, int N, int H, int W, int C, int act, void* stream);
int mg_bn_bwd_reduce(const void* dy, const void* y, const void* conv_out, const float* mean, const float* invstd,
                     float* sums, int N, int H, int W, int C, int act, void* stream);
int mg_bn_bwd_apply(const void* dy, const void* y, const void* conv_out, const float* mean, const float* invstd,
                    const float* gamma, const float* sums, void* dx, void* dres, int N, int H, int W, int C, int act,
                    int pre_act, const float* count_dev, void* stream);

/* ---- K9: sparse refinement on the active-site lists (no spconv) -------------------------------------
 * replaces: spconv SubMConv2d / SparseInverseConv2d / SparseConvTensor.dense() and the dense<->sparse gathers of
 *           decoder/resnet_inst_matt_spconv.py:161-270, plus their backward passes.
 * mg_gather_rows      : out[r, c_off:c_off+C] = dense[coords[r].slot / n_i, y, x, :]   (dense NHWC fp16 [B,H,W,C])
 * mg_scatter_rows_add : ddense[slot / n_i, y, x, :] += g[r, c_off:c_off+C]              (backward of the gather)
 * mg_sparse_conv      : rulebook convolution  out[p][co] = sum_t sum_ci src[table[p][t]][ci] * w[co][t*Cin + ci] + bias
 *     src fp16 rows (row stride src_stride); table int32 [No][T] with -1 = none, or NULL with T = 1 (1x1 / Linear);
 *     w fp16 [ceil16(Cout)][T*Cin] (rows beyond Cout zero); pre_act 1 = ReLU on the result;
 *     out fp16 rows written at column c_off (row stride out_stride), or, when `map` is given (the 32->1 heads), the
 *     fp32 logit map[slot][mapH][mapW] (pre-filled with -99 by the caller) gets ((v - 99) + 99) at coords[p];
 *     stats optional [MG_CONV_STAT_COPIES][2][Cout] as in mg_conv_fprop (BatchNorm1d batch statistics).
 *     SubMConv2d: table = nbr;  SparseInverseConv2d: table = parent;  their data gradients: the same call with the
 *     mirrored nbr taps / the child table and the transposed weight pack.
 * mg_sparse_wgrad     : dw[co][t*Cin + ci] += sum_p dout[p][co] * src[table[p][t]][ci]   (fp32, atomics; caller zeroes) */
typedef struct mg_sparse_conv_desc {
    const void* src; int32_t src_stride;
    const int32_t* table; int32_t T, No, Cin, Cout;
    const void* w; const float* bias;
    void* out; int32_t out_stride, c_off;
    float* stats;
    float* map; const int32_t* coords; int32_t mapH, mapW;
    int32_t pre_act;
} mg_sparse_conv_desc;
int mg_gather_rows(const void* dense, const int32_t* coords, int n, int n_i, int H, int W, int C, void* out, int out_stride,
                   int c_off, void* stream);
int mg_scatter_rows_add(const void* g, int g_stride, int c_off, const int32_t* coords, int n, int n_i, int H, int W, int C,
                        void* ddense, void* stream);
int mg_sparse_conv(const mg_sparse_conv_desc* desc, void* stream);
int mg_sparse_wgrad(const void* dout, int dout_stride, int Cout, const void* src, int src_stride, int Cin,
                    const int32_t* table, int T, int No, float* dw, void* stream);

/* ---- K14: optimizer tail (GradScaler.unscale_ + clip_grad_norm_ + AdamW.step) ------------------------------
 * replaces: engine/train.py:265-283 (scaler.unscale_, torch.nn.utils.clip_grad_norm_(all_params, 0.01), scaler.step) with
 *           engine/optim.py:118 (torch.optim.AdamW) - ~600 parameter tensors -> two multi-tensor launches, no host sync.
 * grad / m / v: flat fp32 buffers of n_flat elements (tensor t occupies [flat_off, flat_off + numel)); tensors[] gives each
 * parameter's address; items[k] = (tensor index, element offset) work items of <= 16384 elements.
 * acc [2] (zero before the first call; the call leaves it zeroed), step [1] (number of updates applied so far, advanced on
 * the device unless a gradient was inf / nan, in which case nothing is updated), report [2] = (unscaled gradient norm,
 * found_inf) of this call.  inv_scale = 1 / loss scale.  skip (optional, uint8 per tensor): tensors that received no gradient
 * this step are left untouched (no weight decay, no moment decay), as torch.optim.AdamW does for `grad is None`. */
typedef struct mg_optim_tensor {
    float* param;
    int64_t flat_off, numel;
} mg_optim_tensor;
int mg_optim_adamw_step(const mg_optim_tensor* tensors, const int32_t* items, int n_items, const float* grad, size_t n_flat,
                        float* m, float* v, float* acc, float* step, float* report, float lr, float beta1, float beta2,
                        float eps, float weight_decay, float max_norm, float inv_scale, const uint8_t* skip, void* stream);

/* ---- K16: loader / evaluator ends of the path ------------------------------------------------------------------
 * mg_input_stage    replaces ToTensor + Normalize (dataloader/transforms.py:720-783) and the dataset's scaling / nearest 1/8
 *                   mask down-sampling (dataloader/him.py:156-157, 175-176): frames uint8 [B][H][W][3], alphas / masks uint8
 *                   [B][n_i][H][W] (either may be NULL together with its output) -> image fp32 [B][3][H][W] =
 *                   (v / 255 - mean) / std, alpha = (a < 5 ? 0 : a) / 255, mask = m / 255 at full size (mask_div 1) or
 *                   [B][n_i][H/8][W/8] (mask_div 8).  mean3 / std3 are HOST arrays of three floats.
 * mg_alpha_finalize replaces reverse_transform_tensor (utils/postprocessing.py:36-64) + the clamps of engine/test.py:141-142:
 *                   in fp32 [planes][h][w]; the (h - pad_h) x (w - pad_w) top-left crop is resized bilinearly
 *                   (align_corners = True) to out_h x out_w (0: no resize, out = the crop); values <= lo become 0, values
 *                   >= hi become 1.                                                                                  */
int mg_input_stage(const void* frames_u8, const void* alphas_u8, const void* masks_u8, float* image, float* alpha, float* mask,
                   const float* mean3, const float* std3, int B, int n_i, int H, int W, int mask_div, void* stream);
int mg_alpha_finalize(const float* in, float* out, int planes, int h, int w, int pad_h, int pad_w, int out_h, int out_w,
                      float lo, float hi, void* stream);

/* ---- K15: SyncBatchNorm-equivalent statistics exchange over peer memory ---------------------------------------
 * replaces: the per-layer all-reduces of nn.SyncBatchNorm (engine/train.py:160-161 converts all 71 BatchNorms when
 *           `model.sync_bn` is true): forward sum / sum-of-squares / element count, backward sum dz / sum dz*xhat.
 * Each rank creates ONE exchange window (device memory of mg_xchg_window_bytes() bytes, exported as a CUDA IPC handle of
 * MG_XCHG_HANDLE_BYTES bytes), the handles travel over the host-side process group, every rank maps every peer's window.
 * mg_stats_exchange : in = `n_copies` copies of [2][C] partial sums (conv epilogue: MG_CONV_STAT_COPIES; backward: 1),
 *                     count = this rank's element count (< 0: none).  The kernel reduces the copies, stores the result
 *                     into every rank's window (NVLink / NVSwitch peer stores), waits for the peers' parts and sums them in
 *                     rank order: out [2][C] (+ out[2*C] = global count) is bit-identical on all ranks.  x == NULL or
 *                     world == 1: local reduction of the copies only (the caller then all-reduces `out` itself).
 *                     Stream-ordered, no host synchronisation, CUDA-graph capturable (the exchange counter lives in the
 *                     window); all ranks must issue the same sequence of exchanges.  Like a collective the kernel WAITS
 *                     for late peers (rank-0-only validation, slow loader workers); a watchdog of NCCL order
 *                     (MAGGIE_B200_XCHG_TIMEOUT_S seconds, default 600 as torch's process groups, 0 = none) traps the kernel
 *                     when a peer never shows up (the next synchronisation reports the failure).                    */
#define MG_XCHG_MAX_RANKS 16
#define MG_XCHG_HANDLE_BYTES 64
typedef struct mg_xchg_desc {
    void* window[MG_XCHG_MAX_RANKS]; /* window[r]: rank r's window as mapped into this process (window[rank]: the local one) */
    int32_t rank, world;
} mg_xchg_desc;
size_t mg_xchg_window_bytes(void);
int mg_xchg_window_create(void** window, void* ipc_handle);
int mg_xchg_window_open(const void* ipc_handle, void** mapped);
int mg_xchg_window_close(void* mapped);
int mg_xchg_window_destroy(void* window);
int mg_stats_exchange(const mg_xchg_desc* x, const float* in, int n_copies, int C, float count, float* out, void* stream);

/* ---- K13: row-wise helpers (LayerNorm with fused residual, column sums) ---------------------------------
 * replaces: nn.LayerNorm after the residual add of every post-norm attention / FFN layer (module/mask_attention.py:
 *           `tgt = self.norm(tgt + self.dropout(tgt2))`, 11 per forward) with its backward, and the bias-gradient
 *           reductions of the sparse layers.
 * mg_layer_norm_fwd : y = LN(a + b) * gamma + beta over rows of E (64 | 128) fp16 elements; b optional; sum_out (optional,
 *                     needed for the backward when b is given) receives the fp16 sum; stat [rows][2] = (mean, rstd).
 * mg_layer_norm_bwd : dx (fp16) from s = a + b, gy; dgb [2][E] fp32 += (dgamma ; dbeta)  (caller zeroes).
 * mg_col_sum        : out[c] += sum_r x[r][c], x fp16 rows with row stride `stride` (caller zeroes out).            */
int mg_layer_norm_fwd(const void* a, const void* b, const float* gamma, const float* beta, float eps, void* sum_out, void* y,
                      float* stat, int rows, int E, void* stream);
int mg_layer_norm_bwd(const void* s, const void* gy, const float* gamma, const float* stat, void* dx, float* dgb, int rows,
                      int E, void* stream);
int mg_col_sum(const void* x, int stride, int rows, int C, float* out, void* stream);
/* fp32 rows, forward only (evaluation at fp32-level accuracy): y = LN(a + b) * gamma + beta                          */
int mg_layer_norm_fwd_f32(const float* a, const float* b, const float* gamma, const float* beta, float eps, float* y, int rows,
                          int E, void* stream);
/* mg_token_logits_fwd/bwd : logits[bt][q][p] = sum_c tok[bt / n_f][q][c] * x[bt][p][c] - the `einsum('bqc,btchw->btqhw')` of
 *     the OS8 head (module/instance_matte_decoder.py:302); x NHWC fp16 with C = 64, tok fp32 [B][Q][C], logits / g fp32
 *     [BT][Q][HW].  bwd: dx fp16 [BT][HW][C] and / or dtok fp32 [B][Q][C] (+=, caller zeroes); either may be NULL. */
int mg_token_logits_fwd(const float* tok, const void* x, float* logits, int BT, int n_f, int Q, int HW, int C, void* stream);
int mg_token_logits_fwd_f32(const float* tok, const float* x, float* logits, int BT, int n_f, int Q, int HW, int C,
                            void* stream);   /* x NHWC fp32 */
int mg_token_logits_bwd(const float* tok, const void* x, const float* g, void* dx, float* dtok, int BT, int n_f, int Q, int HW,
                        int C, void* stream);

/* ---- K7: alpha heads (bilinear upsampling + (tanh + 1) / 2 + per-plane factor) ---------------------------------
 * replaces: F.interpolate(..., mode='bilinear', align_corners=False) + (tanh(x) + 1) / 2 (+ `* valid_masks`) of the three
 *           alpha scales (decoder/resnet_inst_matt_spconv.py:302-309, 357-362) and their autograd backward.
 * logits fp32 [planes,h,w]; out / g fp32 [planes,h*S,w*S]; S in {1,2,4,8}; plane_scale optional fp32 [planes].
 * fwd: out = (tanh(bilerp(logits)) + 1) / 2 * plane_scale.   bwd: glogits = d out / d logits applied to g (gather form). */
/* all_zero (optional, device int32, preset to 1 by the caller): cleared when some output value is non-zero - the device-side
 * form of the reference's `x_os8.sum() == 0` host test (decoder/resnet_inst_matt_spconv.py:314). */
int mg_upsample_tanh_fwd(const float* logits, const float* plane_scale, float* out, int planes, int h, int w, int S,
                         int32_t* all_zero, void* stream);
int mg_upsample_tanh_bwd(const float* logits, const float* plane_scale, const float* g, float* glogits, int planes, int h,
                         int w, int S, void* stream);

/* ---- K12: fused training losses (weighted L1 + 3-level Laplacian pyramid + Sobel gradient, 3 alpha scales) ----
 * replaces: arch/maggie.py:237-346 (regression_loss, compute_loss) + loss.py:67-191 (GradientLoss, LapLoss), ~9
 *           pyramid / stencil passes of tiny cuDNN convolutions per scale, and their autograd backward.
 * a1/a4/a8, target, w1/w4/w8: fp32 [S,H,W] (predictions per scale, GT alpha, per-scale weights).
 * mg_loss_fwd : sums [MG_LOSS_COPIES][3 scales][8] (caller zeroes; add the copies up) with, per scale,
 *               [0] sum|p*w - t*w|, [1..3] sum|L_k|*w_k for pyramid levels k=0..2 (w_k = w[::2^k, ::2^k]),
 *               [4] sum|sobel(p*w) - sobel(t*w)|, [5..7] sum w_k.   ws: mg_loss_workspace_floats() floats;
 *               sg: fp16 scratch of 3*S*H*W*(1 + 1/4 + 1/16) elements (sign(L_k)*w_k, kept for the backward).
 * mg_loss_bwd : coef [3][5] (device) = upstream gradient of numerators [0..4]; writes d/d(a1), d/d(a4), d/d(a8).   */
#define MG_LOSS_COPIES 32
size_t mg_loss_workspace_floats(int S, int H, int W);
/* plane_scale (optional, fp32 [S]): the predictions enter as a * plane_scale[slice] (the reference's `pred * valid_masks`,
 * arch/maggie.py:112-117, without materialising the products); the returned gradients are w.r.t. the unscaled a. */
int mg_loss_fwd(const float* a1, const float* a4, const float* a8, const float* target, const float* w1, const float* w4,
                const float* w8, const float* plane_scale, int S, int H, int W, float* ws, void* sg_f16, float* sums,
                void* stream);
int mg_loss_bwd(const float* a1, const float* a4, const float* a8, const float* target, const float* w1, const float* w4,
                const float* w8, const float* plane_scale, int S, int H, int W, float* ws, const void* sg_f16,
                const float* coef, float* g1_out, float* g4_out, float* g8_out, void* stream);

/* ---- K6: mask-guided attention cores (single head, E = 128) ------------------------------------------
 * replaces: nn.MultiheadAttention's QK^T / softmax / PV (unfused, weights materialised) inside
 *           module/mask_attention.py:99-102 as used by module/instance_matte_decoder.py:219-267, and the
 *           attention-max statistic of :101-109.  (The 128x128 in/out projections run on the K9 rows GEMM.)
 * "tq": FEW queries (tokens, F <= 16, fp32 [B,F,E]) over MANY keys/values (fp16 rows [B,S,E]); key_pad [B,S] and
 *       guidance [B,F,S] optional uint8; out fp32 [B,F,E], stat[b,f] = sum_k guidance*softmax; row_max/row_sum are
 *       kept for the backward; ws: mg_attn_tq_workspace_floats().  Backward: dq fp32 (atomics, caller zeroes),
 *       dk/dv fp16 rows.
 * "fq": MANY queries (fp16 rows [B,S,E]) over FEW keys/values (fp32 [B,F,E], key_pad [B,F]); out fp16 rows.
 *       Backward: dq fp16 rows, dk/dv fp32 (atomics, caller zeroes).                                           */
size_t mg_attn_tq_workspace_floats(int B, int F, int S);
int mg_attn_tq_fwd(const float* q, const void* k, const void* v, const uint8_t* key_pad, const uint8_t* guidance, int B, int F,
                   int S, int E, float* out, float* stat, float* row_max, float* row_sum, float* ws, void* stream);
int mg_attn_tq_bwd(const float* q, const void* k, const void* v, const uint8_t* key_pad, const uint8_t* guidance,
                   const float* out, const float* stat, const float* row_max, const float* row_sum, const float* d_out,
                   const float* d_stat, int B, int F, int S, int E, float* dq, void* dk, void* dv, void* stream);
int mg_attn_fq_fwd(const void* q, const float* k, const float* v, const uint8_t* key_pad, int B, int F, int S, int E,
                   void* out, void* stream);
int mg_attn_fq_bwd(const void* q, const float* k, const float* v, const uint8_t* key_pad, const void* d_out, int B, int F,
                   int S, int E, void* dq, float* dk, float* dv, void* stream);
/* forward cores with an fp32 many side (evaluation at fp32-level accuracy; accurate expf)                          */
int mg_attn_tq_fwd_f32(const float* q, const float* k, const float* v, const uint8_t* key_pad, const uint8_t* guidance, int B,
                       int F, int S, int E, float* out, float* stat, float* row_max, float* row_sum, float* ws, void* stream);
int mg_attn_fq_fwd_f32(const float* q, const float* k, const float* v, const uint8_t* key_pad, int B, int F, int S, int E,
                       float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGGIE_B200_H */
