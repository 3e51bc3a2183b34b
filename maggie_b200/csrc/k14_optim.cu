// K14: optimizer tail of a training step in two multi-tensor launches:
//     GradScaler.unscale_  ->  clip_grad_norm_(all params, max_norm)  ->  AdamW.step (skipped when a gradient is inf/nan)
// (engine/train.py:265-283, engine/optim.py:118).  The gradients live in ONE flat fp32 buffer (the buffer the data-parallel
// all-reduce uses), the moments in two more; the parameters stay where torch put them (a table of pointers + offsets).
//   pass 1: sum of squares of the unscaled gradients + inf/nan flag          (one read of the gradients)
//   pass 2: clip coefficient from pass 1, AdamW update of p, m, v in place; the step counter advances on the device only
//           when no gradient was inf/nan - no host synchronisation anywhere.
#include "common.cuh"

namespace {

constexpr int CHUNK = 16384;   // elements per work item

__global__ void __launch_bounds__(256)
optim_norm_kernel(const float* __restrict__ grad, size_t n, float inv_scale, float* __restrict__ acc) {
    mg::pdl_prologue();
    float s = 0.f;
    int bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float g = grad[i] * inv_scale;
        bad |= !isfinite(g);
        s += g * g;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    __shared__ float red[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = s;
    const int anybad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        if (isfinite(t) && t != 0.f) atomicAdd(acc, t);
        if (anybad || !isfinite(t)) acc[1] = 1.f;
    }
}

struct OptimHyper {
    float lr, beta1, beta2, eps, weight_decay, max_norm, inv_scale;
};

// items[k] = (tensor index, element offset inside the tensor); tensors[t] = (param pointer, flat offset, numel)
__global__ void __launch_bounds__(256)
optim_adamw_kernel(const mg_optim_tensor* __restrict__ tensors, const int2* __restrict__ items, const float* __restrict__ grad,
                   float* __restrict__ m, float* __restrict__ v, const float* __restrict__ acc, float* __restrict__ step,
                   const uint8_t* __restrict__ skip, OptimHyper h) {
    mg::pdl_prologue();
    if (acc[1] != 0.f) return;                          // a gradient was inf / nan: the whole step is skipped
    const int2 it = items[blockIdx.x];
    if (skip && skip[it.x]) return;                     // no gradient this step: torch.optim.AdamW leaves the tensor alone
    const mg_optim_tensor T = tensors[it.x];
    const float total_norm = sqrtf(acc[0]);
    const float clip = fminf(1.f, h.max_norm / (total_norm + 1e-6f)) * h.inv_scale;   // torch: clamp(max_norm / (norm + 1e-6), max=1)
    const float t = step[0] + 1.f;                      // this update's step number (all blocks read the old value)
    const float bc1 = 1.f - powf(h.beta1, t), bc2 = 1.f - powf(h.beta2, t);
    const float step_size = h.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - h.lr * h.weight_decay;
    float* p = T.param + it.y;
    const size_t fo = T.flat_off + it.y;
    const int n = min(CHUNK, (int)(T.numel - it.y));
    for (int i = threadIdx.x; i < n; i += 256) {
        const float g = grad[fo + i] * clip;
        const float mi = h.beta1 * m[fo + i] + (1.f - h.beta1) * g;
        const float vi = h.beta2 * v[fo + i] + (1.f - h.beta2) * g * g;
        m[fo + i] = mi, v[fo + i] = vi;
        p[i] = p[i] * decay - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + h.eps);
    }
}

// runs after the update: advance the step counter unless the step was skipped; publish (norm, found_inf); reset pass 1
__global__ void optim_finish_kernel(float* __restrict__ acc, float* __restrict__ step, float* __restrict__ report) {
    mg::pdl_prologue();
    if (threadIdx.x == 0) {
        report[0] = sqrtf(acc[0]), report[1] = acc[1];
        if (acc[1] == 0.f) step[0] += 1.f;
        acc[0] = 0.f, acc[1] = 0.f;
    }
}

}  // namespace

extern "C" int mg_optim_adamw_step(const mg_optim_tensor* tensors, const int32_t* items, int n_items, const float* grad,
                                   size_t n_flat, float* m, float* v, float* acc, float* step, float* report, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, float max_norm, float inv_scale,
                                   const uint8_t* skip, void* stream) {
    MG_REQUIRE(tensors && items && grad && m && v && acc && step && report, "mg_optim_adamw_step: null pointer");
    MG_REQUIRE(n_items > 0 && n_flat > 0, "mg_optim_adamw_step: no work");
    const int grid = (int)std::min<size_t>((n_flat + 255) / 256, (size_t)mg::kNumSMs * 8);
    MG_LAUNCH(optim_norm_kernel, grid, 256, 0, stream, grad, n_flat, inv_scale, acc);
    OptimHyper h{lr, beta1, beta2, eps, weight_decay, max_norm, inv_scale};
    MG_LAUNCH(optim_adamw_kernel, n_items, 256, 0, stream, tensors, reinterpret_cast<const int2*>(items), grad, m, v,
              const_cast<const float*>(acc), step, skip, h);
    MG_LAUNCH(optim_finish_kernel, 1, 32, 0, stream, acc, step, report);
    MG_CHECK_LAUNCH("mg_optim_adamw_step");
    return MG_OK;
}
