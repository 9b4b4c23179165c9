// ngp_grid.cu -- the Instant-NGP baseline density grid:
// modules/ngp_grid.py:37-152 of the reference (NGPGrid.sample_uniform_and_occupied_cells and
// NGPGrid.update) as a handful of kernels without a host synchronisation.
//
// Random numbers stay on the host side of the boundary (torch) and are passed in; duplicate
// cells in a scatter resolve as on the reference's CPU path (the entry with the largest flat
// index wins) through an atomicMax winner table that is restored to -1 on the way out.
#include "common.cuh"

// ---- occupied-cell sampling, ngp_grid.py:53-60 ----------------------------------------------
// indices2 = nonzero(occ_morton_grid[c] > thr); indices2 = indices2[randint(len(indices2), (M,))]
// Stage 1 counts the occupied cells per 1024-cell block, stage 2 scans the (<= 8192) block
// counts in one block, stage 3 answers M rank queries: the k-th occupied cell in Morton order.
#define NGP_BLOCK_CELLS 1024

__global__ void __launch_bounds__(256) ngp_count_kernel(const float* __restrict__ occ, int64_t n, float thr,
                                                        int32_t* __restrict__ block_counts) {
    __shared__ int32_t sm[8];
    const int64_t base = (int64_t)blockIdx.x * NGP_BLOCK_CELLS + threadIdx.x * 4;
    int32_t c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) c += (base + k < n && occ[base + k] > thr) ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t t = 0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        block_counts[blockIdx.x] = t;
    }
}

// exclusive scan of nb <= 8192 block counts in place, total -> block_counts[nb]
__global__ void __launch_bounds__(1024) ngp_scan_kernel(int32_t* __restrict__ block_counts, int nb) {
    __shared__ int32_t sm[1024];
    const int per = (nb + 1023) / 1024;
    const int lo = threadIdx.x * per;
    int32_t s = 0;
    for (int k = 0; k < per; ++k) if (lo + k < nb) s += block_counts[lo + k];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int32_t v = (threadIdx.x >= off) ? sm[threadIdx.x - off] : 0;
        __syncthreads();
        sm[threadIdx.x] += v;
        __syncthreads();
    }
    int32_t run = sm[threadIdx.x] - s;
    for (int k = 0; k < per; ++k)
        if (lo + k < nb) { const int32_t c = block_counts[lo + k]; block_counts[lo + k] = run; run += c; }
    if (threadIdx.x == 1023) block_counts[nb] = sm[1023];
}

// rank query: out[i] = Morton index of the rand_idx[i]-th occupied cell (rand_idx taken modulo the
// occupied count; -1 when no cell is occupied, the reference then appends nothing, :55)
__global__ void __launch_bounds__(256) ngp_select_kernel(const float* __restrict__ occ, int64_t n, float thr,
                                                         const int32_t* __restrict__ block_starts, int nb,
                                                         const int64_t* __restrict__ rand_idx, int64_t M,
                                                         int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int32_t total = block_starts[nb];
    if (total == 0) { out[i] = -1; return; }
    const int32_t k = (int32_t)(rand_idx[i] % total);
    int lo = 0, hi = nb - 1;                      // last block with start <= k
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (block_starts[mid] <= k) lo = mid; else hi = mid - 1;
    }
    int32_t seen = block_starts[lo];
    const int64_t base = (int64_t)lo * NGP_BLOCK_CELLS;
    int64_t found = -1;
    for (int j = 0; j < NGP_BLOCK_CELLS && base + j < n; ++j) {
        if (occ[base + j] > thr) { if (seen == k) { found = base + j; break; } ++seen; }
    }
    out[i] = found;
}

VN_API int64_t vn_ngp_select_tmp_ints(int64_t n_cells) { return (n_cells + NGP_BLOCK_CELLS - 1) / NGP_BLOCK_CELLS + 4; }

VN_API int vn_ngp_sample_occupied(const float* occ, int64_t n_cells, float threshold, const int64_t* rand_idx, int64_t M,
                                  int32_t* tmp, int64_t* indices_out, void* stream) {
    VN_REQUIRE(n_cells > 0 && M >= 0, "vn_ngp_sample_occupied: bad sizes");
    if (M == 0) return VN_OK;
    VN_REQUIRE(occ && rand_idx && tmp && indices_out, "vn_ngp_sample_occupied: null pointer");
    const int64_t nb = (n_cells + NGP_BLOCK_CELLS - 1) / NGP_BLOCK_CELLS;
    VN_REQUIRE(nb <= 1024 * 1024, "vn_ngp_sample_occupied: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    ngp_count_kernel<<<(unsigned)nb, 256, 0, st>>>(occ, n_cells, threshold, tmp);
    VN_CHECK_LAUNCH("ngp_count_kernel");
    ngp_scan_kernel<<<1, 1024, 0, st>>>(tmp, (int)nb);
    VN_CHECK_LAUNCH("ngp_scan_kernel");
    ngp_select_kernel<<<vn_blocks(M, 256), 256, 0, st>>>(occ, n_cells, threshold, tmp, (int)nb, rand_idx, M, indices_out);
    VN_CHECK_LAUNCH("ngp_select_kernel");
    return VN_OK;
}

// ---- cell -> random world position inside the cell, ngp_grid.py:139-143 ---------------------
//   xyzs_w = (coords / (G - 1) * 2 - 1) * (s - half_grid_size); xyzs_w += (rand * 2 - 1) * half_grid_size
// one IEEE operation per torch op, python scalars (s - hgs), hgs rounded to f32 by the caller
__global__ void __launch_bounds__(256) ngp_positions_kernel(const int32_t* __restrict__ coords, const float* __restrict__ noise,
                                                            int64_t n3, float gm1, float span, float half_cell,
                                                            float* __restrict__ xyzs_w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    float v = vn_div((float)coords[i], gm1);
    v = vn_mul(vn_sub(vn_mul(v, 2.0f), 1.0f), span);
    const float jit = vn_mul(vn_sub(vn_mul(__ldg(noise + i), 2.0f), 1.0f), half_cell);
    xyzs_w[i] = vn_add(v, jit);
}

VN_API int vn_ngp_cell_positions(const int32_t* coords, const float* noise, int64_t M, int grid_size, float span,
                                 float half_cell, float* xyzs_w, void* stream) {
    VN_REQUIRE(M >= 0 && grid_size >= 2, "vn_ngp_cell_positions: bad sizes");
    if (M == 0) return VN_OK;
    VN_REQUIRE(coords && noise && xyzs_w, "vn_ngp_cell_positions: null pointer");
    ngp_positions_kernel<<<vn_blocks(3 * M, 256), 256, 0, (cudaStream_t)stream>>>(coords, noise, 3 * M, (float)(grid_size - 1),
                                                                               span, half_cell, xyzs_w);
    VN_CHECK_LAUNCH("ngp_positions_kernel");
    return VN_OK;
}

// ---- density_grid_tmp[c, indices] = sigma, then the decayed maximum, ngp_grid.py:144-152 -----
__global__ void __launch_bounds__(256) ngp_winner_kernel(const int64_t* __restrict__ indices, int64_t M, int64_t n_cells,
                                                         int32_t* __restrict__ winner) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t c = indices[i];
    if (c >= 0 && c < n_cells) atomicMax(winner + c, (int32_t)i);
}
__global__ void __launch_bounds__(256) ngp_scatter_kernel(const int64_t* __restrict__ indices, const float* __restrict__ sigmas,
                                                          int64_t M, int64_t n_cells, int32_t* __restrict__ winner,
                                                          float* __restrict__ tmp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t c = indices[i];
    if (c >= 0 && c < n_cells && winner[c] == (int32_t)i) { tmp[c] = sigmas[i]; winner[c] = -1; }
}
__global__ void __launch_bounds__(256) ngp_decay_max_kernel(float* __restrict__ occ, float* __restrict__ tmp, int64_t n,
                                                            float decay, const float* __restrict__ decay_cells) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float o = occ[i], t = tmp[i];
    tmp[i] = 0.0f;                                            // density_grid_tmp = zeros_like(...) of the next update
    if (o < 0.0f) return;                                     // torch.where(grid < 0, grid, ...)
    const float d = vn_mul(o, decay_cells ? __ldg(decay_cells + i) : decay);
    occ[i] = (d != d || t != t) ? NAN : fmaxf(d, t);          // torch.maximum propagates NaN
}

VN_API int vn_ngp_grid_update(float* occ, float* tmp, int32_t* winner, int64_t n_cells, const int64_t* indices,
                              const float* sigmas, int64_t M, float decay, const float* decay_cells, void* stream) {
    VN_REQUIRE(n_cells > 0 && M >= 0 && M < ((int64_t)1 << 31), "vn_ngp_grid_update: bad sizes");
    VN_REQUIRE(occ && tmp && winner && (M == 0 || (indices && sigmas)), "vn_ngp_grid_update: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (M > 0) {
        ngp_winner_kernel<<<vn_blocks(M, 256), 256, 0, st>>>(indices, M, n_cells, winner);
        VN_CHECK_LAUNCH("ngp_winner_kernel");
        ngp_scatter_kernel<<<vn_blocks(M, 256), 256, 0, st>>>(indices, sigmas, M, n_cells, winner, tmp);
        VN_CHECK_LAUNCH("ngp_scatter_kernel");
    }
    ngp_decay_max_kernel<<<vn_blocks(n_cells, 256), 256, 0, st>>>(occ, tmp, n_cells, decay, decay_cells);
    VN_CHECK_LAUNCH("ngp_decay_max_kernel");
    return VN_OK;
}

// ---- threshold = min(mean(grid[grid > 0]), density_threshold) and the bitfield, :154-163 ------
#define NGP_MEAN_BLOCKS 256
__global__ void __launch_bounds__(256) ngp_mean_stage1(const float* __restrict__ occ, int64_t n, double* __restrict__ partial_sum,
                                                       long long* __restrict__ partial_cnt) {
    __shared__ double ss[256];
    __shared__ long long sc[256];
    double acc = 0.0;
    long long cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)NGP_MEAN_BLOCKS * 256) {
        const float v = occ[i];
        if (v > 0.0f) { acc += (double)v; ++cnt; }
    }
    ss[threadIdx.x] = acc; sc[threadIdx.x] = cnt;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { ss[threadIdx.x] += ss[threadIdx.x + s]; sc[threadIdx.x] += sc[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial_sum[blockIdx.x] = ss[0]; partial_cnt[blockIdx.x] = sc[0]; }
}
__global__ void __launch_bounds__(256) ngp_mean_stage2(const double* __restrict__ partial_sum, const long long* __restrict__ partial_cnt,
                                                       float density_threshold, float* __restrict__ out2) {
    __shared__ double ss[256];
    __shared__ long long sc[256];
    ss[threadIdx.x] = partial_sum[threadIdx.x]; sc[threadIdx.x] = partial_cnt[threadIdx.x];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { ss[threadIdx.x] += ss[threadIdx.x + s]; sc[threadIdx.x] += sc[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float mean = sc[0] > 0 ? (float)(ss[0] / (double)sc[0]) : NAN;     // mean of an empty selection is nan
        out2[0] = mean;
        out2[1] = (density_threshold < mean) ? density_threshold : mean;       // python min(mean, thr): nan stays nan
    }
}
__global__ void __launch_bounds__(256) ngp_pack_kernel(const float4* __restrict__ grid, int64_t n_bytes,
                                                       const float* __restrict__ thr2, uint8_t* __restrict__ bitfield) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_bytes) return;
    const float thr = thr2[1];
    const float4 a = __ldg(grid + 2 * n), b = __ldg(grid + 2 * n + 1);
    uint32_t bits = 0;                                         // modules/utils.py:157-169, strict >
    bits |= (a.x > thr) ? 1u : 0u; bits |= (a.y > thr) ? 2u : 0u; bits |= (a.z > thr) ? 4u : 0u; bits |= (a.w > thr) ? 8u : 0u;
    bits |= (b.x > thr) ? 16u : 0u; bits |= (b.y > thr) ? 32u : 0u; bits |= (b.z > thr) ? 64u : 0u; bits |= (b.w > thr) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// scratch: 8-byte aligned, >= vn_ngp_threshold_tmp_bytes() bytes; thr_out [2] = (mean, threshold)
VN_API int64_t vn_ngp_threshold_tmp_bytes(void) { return (int64_t)NGP_MEAN_BLOCKS * 16; }

VN_API int vn_ngp_threshold_pack(const float* occ, int64_t n_cells_total, float density_threshold, void* scratch,
                                 float* thr_out, uint8_t* bitfield, void* stream) {
    VN_REQUIRE(n_cells_total > 0 && n_cells_total % 8 == 0, "vn_ngp_threshold_pack: cell count must be a positive multiple of 8");
    VN_REQUIRE(occ && scratch && thr_out && bitfield, "vn_ngp_threshold_pack: null pointer");
    VN_REQUIRE(vn_aligned(occ, 16) && vn_aligned(scratch, 8), "vn_ngp_threshold_pack: misaligned buffer");
    cudaStream_t st = (cudaStream_t)stream;
    double* ps = (double*)scratch;
    long long* pc = (long long*)(ps + NGP_MEAN_BLOCKS);
    ngp_mean_stage1<<<NGP_MEAN_BLOCKS, 256, 0, st>>>(occ, n_cells_total, ps, pc);
    VN_CHECK_LAUNCH("ngp_mean_stage1");
    ngp_mean_stage2<<<1, 256, 0, st>>>(ps, pc, density_threshold, thr_out);
    VN_CHECK_LAUNCH("ngp_mean_stage2");
    const int64_t nb = n_cells_total / 8;
    ngp_pack_kernel<<<vn_blocks(nb, 256), 256, 0, st>>>((const float4*)occ, nb, thr_out, bitfield);
    VN_CHECK_LAUNCH("ngp_pack_kernel");
    return VN_OK;
}
