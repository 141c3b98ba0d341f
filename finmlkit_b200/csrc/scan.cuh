// scan.cuh -- deterministic chunked device scan (reduce -> scan of tile sums -> apply), hand-written, no CUB.
// Used for: volume prefix sums (volume bars), the synthetic generator, footprint level offsets.
#pragma once
#include "common.cuh"

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T warp_incl_scan_sum(T x) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (unsigned)o) x = y + x;   // earlier lanes on the left: the operator may be non-commutative
    }
    return x;
}

// Block-wide exclusive scan of one value per thread (blockDim.x == SCAN_THREADS) for any associative `+` with
// identity T(0) (no subtraction: the operator need not be invertible).  Returns the exclusive prefix; total in *tot.
template <typename T>
__device__ __forceinline__ T block_excl_scan_sum(T x, T *tot) {
    __shared__ T warp_tot[SCAN_THREADS / 32];
    __shared__ T block_tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = warp_incl_scan_sum(x);
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
        T v = lane < SCAN_THREADS / 32 ? warp_tot[lane] : T(0);
        T s = warp_incl_scan_sum(v);
        T e = __shfl_up_sync(0xffffffffu, s, 1);
        if (lane == 0) e = T(0);
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = e;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) block_tot = s;
    }
    __syncthreads();
    T ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = T(0);
    T r = warp_tot[w] + ex;
    *tot = block_tot;
    __syncthreads();
    return r;
}


// Thread t of a block owns elements [8t, 8t+8) of its tile (blocked order keeps the scan operator's order).  Reading them
// straight from global memory makes every lane touch its own 64-byte segment; for small element types the tile is
// instead fetched with fully coalesced row loads and transposed through padded shared memory.
template <typename T, typename InF>
__device__ __forceinline__ void scan_load_items(InF &in, int64_t tile_base, int64_t n, T (&vals)[SCAN_ITEMS]) {
    if constexpr (sizeof(T) <= 8) {
        __shared__ __align__(16) unsigned char raw_in[(SCAN_TILE + SCAN_TILE / 8) * 8];
        T *tile = reinterpret_cast<T *>(raw_in);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const int e = threadIdx.x + k * SCAN_THREADS;
            const int64_t i = tile_base + e;
            tile[e + (e >> 3)] = (i < n) ? in(i) : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) vals[k] = tile[threadIdx.x * (SCAN_ITEMS + 1) + k];
        __syncthreads();
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const int64_t i = tile_base + (int64_t)threadIdx.x * SCAN_ITEMS + k;
            vals[k] = (i < n) ? in(i) : T(0);
        }
    }
}

template <typename T, typename InF>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(InF in, int64_t n, T *tile_sums) {
    T vals[SCAN_ITEMS];
    scan_load_items<T>(in, (int64_t)blockIdx.x * SCAN_TILE, n, vals);
    T s = T(0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s = s + vals[k];
    T tot;
    block_excl_scan_sum(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of tile sums; writes grand total to tile_sums[ntiles]
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(T *tile_sums, int64_t ntiles) {
    __shared__ T carry_s;
    if (threadIdx.x == 0) carry_s = T(0);
    __syncthreads();
    for (int64_t b = 0; b < ntiles; b += SCAN_TILE) {
        T vals[SCAN_ITEMS];
        T s = T(0);
        const int64_t base = b + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            vals[k] = (base + k < ntiles) ? tile_sums[base + k] : T(0);
            s = s + vals[k];
        }
        T tot;
        T ex = block_excl_scan_sum(s, &tot);
        T run = carry_s + ex;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < ntiles) tile_sums[base + k] = run;
            run = run + vals[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry_s + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry_s;
}

// out(i, inclusive_prefix_i)
template <typename T, typename InF, typename OutF>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(InF in, OutF out, int64_t n, const T *tile_offsets) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T vals[SCAN_ITEMS];
    scan_load_items<T>(in, (int64_t)blockIdx.x * SCAN_TILE, n, vals);
    T s = T(0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s = s + vals[k];
    T tot;
    T ex = block_excl_scan_sum(s, &tot);
    T run = tile_offsets[blockIdx.x] + ex;
    if constexpr (sizeof(T) <= 8) {
        // transpose the results back so that out() is called in coalesced order
        __shared__ __align__(16) unsigned char raw_out[(SCAN_TILE + SCAN_TILE / 8) * 8];
        T *otile = reinterpret_cast<T *>(raw_out);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            run = run + vals[k];
            otile[threadIdx.x * (SCAN_ITEMS + 1) + k] = run;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const int e = threadIdx.x + k * SCAN_THREADS;
            const int64_t i = (int64_t)blockIdx.x * SCAN_TILE + e;
            if (i < n) out(i, otile[e + (e >> 3)]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + k;
            run = run + vals[k];
            if (i < n) out(i, run);
        }
    }
}

// Inclusive scan: out(i, sum_{j<=i} in(j)). total_out (device pointer, may be null) receives the grand total.
template <typename T, typename InF, typename OutF>
static int device_inclusive_scan(fmk_ctx *ctx, InF in, OutF out, int64_t n, T *total_out_dev) {
    if (n <= 0) return FMK_OK;
    const int64_t ntiles = cdiv(n, SCAN_TILE);
    Scratch<T> sums(ctx);
    FMK_TRY(sums.alloc(ntiles + 1));
    FMK_LAUNCH(ctx, (k_scan_reduce<T, InF>), (unsigned)ntiles, SCAN_THREADS, 0, in, n, sums.p);
    FMK_LAUNCH(ctx, (k_scan_tiles<T>), 1, SCAN_THREADS, 0, sums.p, ntiles);
    FMK_LAUNCH(ctx, (k_scan_apply<T, InF, OutF>), (unsigned)ntiles, SCAN_THREADS, 0, in, out, n, sums.p);
    if (total_out_dev) {
        FMK_CUDA(ctx, cudaMemcpyAsync(total_out_dev, sums.p + ntiles, sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return FMK_OK;
}
