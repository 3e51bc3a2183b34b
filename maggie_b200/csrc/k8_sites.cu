// K8b: active-site lists + rulebook tables for the sparse refinement, without spconv's hash tables.
//
// Representation: per level l (OS1, OS2, OS4, OS8) a bit image [slots, H>>l, ceil((W>>l)/32)] plus a rank
// structure (exclusive popcount prefix per word).  row(site) = rank[word] + popc(word & below(bit)): sites
// are numbered in lexicographic (slot,y,x) order, i.e. exactly torch.nonzero order.  0.25 B/px of metadata
// instead of a 4 B/px dense index map.  Tables are then produced with one thread per (site, tap).
#include "common.cuh"

namespace {

constexpr int NLEV = 4;
constexpr int SCAN_BLOCK = 1024;  // words per scan block; level regions are padded to this

struct Layout {
    int H[NLEV], W[NLEV], Wd[NLEV];
    unsigned words[NLEV];      // real words per level
    unsigned wbase[NLEV + 1];  // padded word base per level (multiples of SCAN_BLOCK)
    unsigned nblocks;          // total scan blocks
    size_t off_bits, off_rank, off_blocksum, off_sitebase, bytes;
};

Layout make_layout(int slots, int H, int W) {
    Layout L;
    unsigned base = 0;
    for (int l = 0; l < NLEV; ++l) {
        L.H[l] = H >> l, L.W[l] = W >> l, L.Wd[l] = (L.W[l] + 31) / 32;
        L.words[l] = (unsigned)slots * L.H[l] * L.Wd[l];
        L.wbase[l] = base;
        base += (L.words[l] + SCAN_BLOCK - 1) / SCAN_BLOCK * SCAN_BLOCK;
    }
    L.wbase[NLEV] = base;
    L.nblocks = base / SCAN_BLOCK;
    L.off_bits = 0;
    L.off_rank = mg::align_up(L.off_bits + (size_t)base * 4, 256);
    L.off_blocksum = mg::align_up(L.off_rank + (size_t)base * 4, 256);
    L.off_sitebase = mg::align_up(L.off_blocksum + (size_t)(L.nblocks + 1) * 4, 256);
    L.bytes = L.off_sitebase + 256;
    return L;
}

// ---- level 0 bits from the uint8 roi ------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bits_from_u8_kernel(const uint8_t* __restrict__ roi, uint32_t* __restrict__ bits, int rows, int W, int Wd) {
    mg::pdl_prologue();
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)rows * Wd) return;
    const int row = idx / Wd, j = idx - row * Wd;
    const uint8_t* p = roi + (size_t)row * W + (j << 5);
    uint32_t acc = 0u;
    if ((W & 31) == 0) {
        const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(p)), m1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        const uint32_t wds[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t t = wds[q];
            acc |= (((t & 0xffu) != 0) | (((t & 0xff00u) != 0) << 1) | (((t & 0xff0000u) != 0) << 2) |
                    (((t & 0xff000000u) != 0) << 3)) << (4 * q);
        }
    } else {
        for (int b = 0; b < 32 && (j << 5) + b < W; ++b) acc |= (uint32_t)(p[b] != 0) << b;
    }
    bits[idx] = acc;
}

__device__ __forceinline__ uint32_t compress_even(uint64_t x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)x;
}

// ---- level l+1 bits: 3x3 / stride 2 / pad 1 "any" pooling of level l (SparseConv2d k3 s2 p1 index rule)
__global__ void __launch_bounds__(256)
downscale_bits_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int slots, int Hi, int Wdi,
                      int Ho, int Wo, int Wdo) {
    mg::pdl_prologue();
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)slots * Ho * Wdo) return;
    const int jw = idx % Wdo, qy = (idx / Wdo) % Ho, s = idx / (Wdo * Ho);
    uint32_t wl = 0, w0 = 0, w1 = 0, wr = 0;
    for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * qy + dy;
        if (y < 0 || y >= Hi) continue;
        const uint32_t* row = in + ((size_t)s * Hi + y) * Wdi;
        const int c = 2 * jw;
        if (c - 1 >= 0) wl |= row[c - 1];
        if (c < Wdi) w0 |= row[c];
        if (c + 1 < Wdi) w1 |= row[c + 1];
        if (c + 2 < Wdi) wr |= row[c + 2];
    }
    const uint64_t V = ((uint64_t)w1 << 32) | w0;
    const uint64_t h = V | ((V << 1) | (wl >> 31)) | ((V >> 1) | ((uint64_t)(wr & 1u) << 63));
    uint32_t o = compress_even(h);
    const int x0 = jw << 5;
    if (x0 + 32 > Wo) o &= (1u << (Wo - x0)) - 1u;
    out[idx] = o;
}

// ---- rank structure ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scan_local_kernel(const uint32_t* __restrict__ bits, uint32_t* __restrict__ rank, uint32_t* __restrict__ blocksum) {
    mg::pdl_prologue();
    // one block = SCAN_BLOCK words, 4 per thread
    __shared__ uint32_t warp_tot[8];
    const unsigned base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    const uint4 w = *reinterpret_cast<const uint4*>(bits + base);
    const uint32_t c0 = __popc(w.x), c1 = __popc(w.y), c2 = __popc(w.z), c3 = __popc(w.w);
    const uint32_t tot = c0 + c1 + c2 + c3;
    uint32_t inc = tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0;
    for (int i = 0; i < warp; ++i) wbase += warp_tot[i];
    const uint32_t ex = wbase + inc - tot;
    uint4 r;
    r.x = ex, r.y = ex + c0, r.z = ex + c0 + c1, r.w = ex + c0 + c1 + c2;
    *reinterpret_cast<uint4*>(rank + base) = r;
    if (threadIdx.x == 255) blocksum[blockIdx.x] = wbase + inc;
}

struct LevelBlocks {
    unsigned start[NLEV + 1];  // first scan block of each level
};

__global__ void __launch_bounds__(1024)
scan_blocks_kernel(uint32_t* __restrict__ blocksum, unsigned nblocks, LevelBlocks lb, uint32_t* __restrict__ sitebase,
                   int32_t* __restrict__ counts) {
    mg::pdl_prologue();
    // exclusive scan of blocksum[0..nblocks) in place (single CTA), blocksum[nblocks] = total
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned start = 0; start < nblocks; start += 1024) {
        const unsigned i = start + threadIdx.x;
        const uint32_t v = i < nblocks ? blocksum[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
        const uint32_t carry = carry_s;
        if (i < nblocks) blocksum[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        blocksum[nblocks] = carry_s;
        for (int l = 0; l <= NLEV; ++l) sitebase[l] = blocksum[lb.start[l]];
        for (int l = 0; l < NLEV; ++l) counts[l] = (int32_t)(blocksum[lb.start[l + 1]] - blocksum[lb.start[l]]);
    }
}

// ---- tables -------------------------------------------------------------------------------------------
struct LevelView {
    const uint32_t* bits;
    const uint32_t* rank;      // block-local exclusive prefix
    const uint32_t* blocksum;  // global exclusive prefix per scan block (indexed by absolute block)
    unsigned wbase;            // padded word base of this level (for the block index)
    uint32_t sitebase;         // global rank of this level's first site
    int H, W, Wd;
};

__device__ __forceinline__ int site_row(const LevelView& L, int s, int y, int x) {
    if (y < 0 || y >= L.H || x < 0 || x >= L.W) return -1;
    const unsigned widx = ((unsigned)s * L.H + y) * L.Wd + (x >> 5);
    const uint32_t w = L.bits[widx];
    const int b = x & 31;
    if (!((w >> b) & 1u)) return -1;
    return (int)(L.blocksum[(L.wbase + widx) / SCAN_BLOCK] + L.rank[widx] - L.sitebase + __popc(w & ((1u << b) - 1u)));
}

__global__ void __launch_bounds__(256)
coords_kernel(LevelView L, unsigned words, int32_t* __restrict__ coords) {
    mg::pdl_prologue();
    const unsigned widx = blockIdx.x * blockDim.x + threadIdx.x;
    if (widx >= words) return;
    uint32_t w = L.bits[widx];
    if (!w) return;
    const int jw = widx % L.Wd, y = (widx / L.Wd) % L.H, s = widx / (L.Wd * L.H);
    int row = (int)(L.blocksum[(L.wbase + widx) / SCAN_BLOCK] + L.rank[widx] - L.sitebase);
    while (w) {
        const int b = __ffs(w) - 1;
        w &= w - 1;
        int32_t* c = coords + (size_t)row * 3;
        c[0] = s, c[1] = y, c[2] = (jw << 5) + b;
        ++row;
    }
}

// mode 0: SubM 3x3 neighbours in the same level; mode 1: parent (coarser) row per tap; mode 2: child (finer) row.
template <int MODE>
__global__ void __launch_bounds__(256)
table_kernel(LevelView other, const int32_t* __restrict__ coords, int n, int32_t* __restrict__ table) {
    mg::pdl_prologue();
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)n * 9) return;
    const int r = t / 9, k = t - r * 9, ky = k / 3, kx = k - ky * 3;
    const int s = coords[r * 3], y = coords[r * 3 + 1], x = coords[r * 3 + 2];
    int out = -1;
    if (MODE == 0) {
        out = site_row(other, s, y + ky - 1, x + kx - 1);
    } else if (MODE == 1) {  // p = 2q - 1 + k  ->  q = (p + 1 - k) / 2 when even
        const int ty = y + 1 - ky, tx = x + 1 - kx;
        if (ty >= 0 && tx >= 0 && !(ty & 1) && !(tx & 1)) out = site_row(other, s, ty >> 1, tx >> 1);
    } else {
        out = site_row(other, s, 2 * y - 1 + ky, 2 * x - 1 + kx);
    }
    table[t] = out;
}

}  // namespace

extern "C" size_t mg_sites_workspace(int slots, int H, int W) {
    if (slots <= 0 || H <= 0 || W <= 0) return 256;
    return make_layout(slots, H, W).bytes;
}

extern "C" int mg_sites_count(const uint8_t* roi, int slots, int H, int W, void* ws, int32_t* counts, void* stream) {
    MG_REQUIRE(roi && ws && counts, "mg_sites_count: null pointer");
    MG_REQUIRE(slots > 0 && H > 0 && W > 0 && (H % 8) == 0 && (W % 8) == 0, "mg_sites_count: H, W must be positive multiples of 8 (%d x %d x %d)", slots, H, W);
    MG_REQUIRE((double)slots * H * W < 2.0e9, "mg_sites_count: too many pixels for int32 rows");
    const Layout L = make_layout(slots, H, W);
    char* base = static_cast<char*>(ws);
    uint32_t* bits = reinterpret_cast<uint32_t*>(base + L.off_bits);
    uint32_t* rank = reinterpret_cast<uint32_t*>(base + L.off_rank);
    uint32_t* blocksum = reinterpret_cast<uint32_t*>(base + L.off_blocksum);
    uint32_t* sitebase = reinterpret_cast<uint32_t*>(base + L.off_sitebase);
    cudaStream_t st = (cudaStream_t)stream;
    // padding words between levels must be zero for the scan
    if (cudaMemsetAsync(bits, 0, (size_t)L.wbase[NLEV] * 4, st) != cudaSuccess) {
        mg::set_error("mg_sites_count: memset failed");
        return MG_ERR_CUDA;
    }
    MG_LAUNCH(bits_from_u8_kernel, mg::ceil_div((int)L.words[0], 256), 256, 0, stream, roi, bits + L.wbase[0],
              slots * L.H[0], L.W[0], L.Wd[0]);
    for (int l = 1; l < NLEV; ++l)
        MG_LAUNCH(downscale_bits_kernel, mg::ceil_div((int)L.words[l], 256), 256, 0, stream, bits + L.wbase[l - 1],
                  bits + L.wbase[l], slots, L.H[l - 1], L.Wd[l - 1], L.H[l], L.W[l], L.Wd[l]);
    MG_LAUNCH(scan_local_kernel, L.nblocks, 256, 0, stream, bits, rank, blocksum);
    LevelBlocks lb;
    for (int l = 0; l <= NLEV; ++l) lb.start[l] = L.wbase[l] / SCAN_BLOCK;
    MG_LAUNCH(scan_blocks_kernel, 1, 1024, 0, stream, blocksum, L.nblocks, lb, sitebase, counts);
    MG_CHECK_LAUNCH("mg_sites_count");
    return MG_OK;
}

extern "C" int mg_sites_tables(const void* ws, int slots, int H, int W, const int32_t* counts_host,
                               int32_t* const* coords, int32_t* const* nbr, int32_t* const* parent,
                               int32_t* const* child, void* stream) {
    MG_REQUIRE(ws && counts_host && coords, "mg_sites_tables: null pointer");
    const Layout L = make_layout(slots, H, W);
    const char* base = static_cast<const char*>(ws);
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(base + L.off_bits);
    const uint32_t* rank = reinterpret_cast<const uint32_t*>(base + L.off_rank);
    const uint32_t* blocksum = reinterpret_cast<const uint32_t*>(base + L.off_blocksum);
    LevelView V[NLEV];
    uint32_t sb = 0;
    for (int l = 0; l < NLEV; ++l) {
        V[l].bits = bits + L.wbase[l], V[l].rank = rank + L.wbase[l], V[l].blocksum = blocksum;
        V[l].wbase = L.wbase[l], V[l].sitebase = sb, V[l].H = L.H[l], V[l].W = L.W[l], V[l].Wd = L.Wd[l];
        sb += (uint32_t)counts_host[l];
    }
    for (int l = 0; l < NLEV; ++l) {
        const int n = counts_host[l];
        if (n <= 0) continue;
        const bool need = coords[l] != nullptr;
        MG_REQUIRE(need || !((nbr && nbr[l]) || (parent && parent[l]) || (child && child[l])),
                   "mg_sites_tables: level %d tables requested without coords", l);
        if (!need) continue;
        MG_LAUNCH(coords_kernel, mg::ceil_div((int)L.words[l], 256), 256, 0, stream, V[l], L.words[l], coords[l]);
        const int tb = mg::ceil_div(n * 9, 256);
        if (nbr && nbr[l]) MG_LAUNCH(table_kernel<0>, tb, 256, 0, stream, V[l], coords[l], n, nbr[l]);
        if (parent && parent[l] && l + 1 < NLEV)
            MG_LAUNCH(table_kernel<1>, tb, 256, 0, stream, V[l + 1], coords[l], n, parent[l]);
        if (child && child[l] && l >= 1) MG_LAUNCH(table_kernel<2>, tb, 256, 0, stream, V[l - 1], coords[l], n, child[l]);
    }
    MG_CHECK_LAUNCH("mg_sites_tables");
    return MG_OK;
}
