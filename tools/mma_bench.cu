// Dev tool: issue-rate microbenchmark of tcgen05.mma (kind::f16, cta_group::1, M = 128, K = 16, SS operands) on sm_100a.
// One CTA per SM; one thread issues `iters` x `per_iter` MMAs on (garbage) shared-memory operands and commits once.
// Prints cycles per MMA for N in {32, 64, 128, 256}, aligned and shifted (non-atom-aligned) A starts, 128-B and 64-B
// swizzle, with the A or the B descriptor advancing / fixed.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o /tmp/mma_bench tools/mma_bench.cu -I maggie_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace mg::ptx;

struct Cfg {
    int N, a_shift_bytes, swizzle, a_rows_step, b_rows_step, iters, per_iter, ts, nacc;
};

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    // operands: zeros (finite)
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1) {
        // whole warp runs the loop, one elected lane issues: every operand stays in uniform registers (no waterfall)
        const uint32_t tm = uniform_u32(tmem);
        const uint32_t idesc = instr_desc_f16(128, c.N, 0, 0);
        const uint32_t lay = swizzle_layout(c.swizzle), sbo = 8 * c.swizzle;
        const uint32_t a0 = smem_u32(smem) + 1024 + c.a_shift_bytes, b0 = smem_u32(smem) + 96 * 1024;
        const uint64_t adesc = smem_desc(a0, 0, sbo, lay), bdesc = smem_desc(b0, 0, sbo, lay);
        const uint32_t astep = (c.a_rows_step * c.swizzle) >> 4, bstep = (c.b_rows_step * c.swizzle) >> 4;
        const uint32_t acc1 = c.nacc > 1 ? c.N : 0, acc2 = c.nacc > 2 ? 2 * c.N : 0, acc3 = c.nacc > 2 ? 3 * c.N : acc1;
        const long long t0 = clock64();
        for (int it = 0; it < c.iters; ++it) {
            if (elect_one()) {
                if (c.ts) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                                     ::"r"(tm), "r"(tm + 256), "l"(bdesc + j * bstep), "r"(idesc), "r"(1) : "memory");
                } else {
                    mma_f16(tm, adesc, bdesc, idesc, 1);
                    mma_f16(tm + acc1, adesc + astep, bdesc + bstep, idesc, 1);
                    mma_f16(tm + acc2, adesc + 2 * astep, bdesc + 2 * bstep, idesc, 1);
                    mma_f16(tm + acc3, adesc + 3 * astep, bdesc + 3 * bstep, idesc, 1);
                }
            }
        }
        if (elect_one()) mma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int Ns[] = {32, 64, 128, 256};
    struct V { const char* name; int shift, swz, astep, bstep, ts, nacc; };
    const V vs[] = {
        {"aligned, one accumulator, sw128", 0, 128, 0, 0, 0, 1},
        {"aligned, 2 accumulators round robin", 0, 128, 0, 0, 0, 2},
        {"aligned, 4 accumulators round robin", 0, 128, 0, 0, 0, 4},
        {"A shifted 1 row, 4 accumulators", 128, 128, 0, 0, 0, 4},
        {"A shifted 66 rows, 4 acc, A advances", 66 * 128, 128, 128, 0, 0, 4},
        {"sw64, A shifted 1 row, 4 acc", 64, 64, 0, 0, 0, 4},
        {"A in TMEM (TS), 1 acc", 0, 128, 0, 0, 1, 1},
    };
    for (const V& v : vs) {
        printf("%-40s", v.name);
        for (int N : Ns) {
            Cfg c{N, v.shift, v.swz, v.astep, v.bstep, 200, 4, v.ts, (v.nacc * N <= 512) ? v.nacc : 512 / N};
            bench<<<148, 128, 210 * 1024>>>(c, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf(" N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
            long long cyc;
            cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
            printf("  N=%3d: %6.1f cyc/MMA", N, (double)cyc / (c.iters * c.per_iter));
        }
        printf("\n");
    }
    return 0;
}
