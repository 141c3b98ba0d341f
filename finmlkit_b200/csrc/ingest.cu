// ingest.cu -- the two scans that sit right before the bar path: the tick rule
// (comp_trade_side_vector, bar/utils.py:12-46) and split-trade merging (merge_split_trades, bar/utils.py:263-329).
#include <math.h>
#include <new>
#include "common.cuh"
#include "scan.cuh"

// ---- tick rule: side_i = sign(p_i - p_{i-1}) if |dp| > 1e-12 else side_{i-1}; side_0 = 0 -----------------------------
// A "last non-zero" forward fill: scan with the operator (a, b) -> b != 0 ? b : a.
struct SideC {
    int v;
    __device__ SideC() {}
    __device__ explicit SideC(int x) { v = x; }     // T(0) is the identity
};
__device__ __forceinline__ SideC operator+(const SideC &a, const SideC &b) { return b.v != 0 ? b : a; }
__device__ __forceinline__ SideC __shfl_up_sync(unsigned m, const SideC &x, int o) {
    SideC r;
    r.v = ::__shfl_up_sync(m, x.v, o);
    return r;
}
struct SideIn {
    const double *p;
    __device__ SideC operator()(int64_t i) const {
        if (i == 0) return SideC(0);
        const double dp = __dadd_rn(p[i], -p[i - 1]);
        if (!(fabs(dp) > 1e-12)) return SideC(0);           // also NaN: keeps the previous side
        return SideC(dp > 0.0 ? 1 : -1);
    }
};
struct SideOut {
    int8_t *s;
    __device__ void operator()(int64_t i, const SideC &c) const { s[i] = (int8_t)c.v; }
};

// tick rule on a device-resident price column (used by the imbalance / run bar indexers)
int fmk_tick_rule_device(fmk_ctx *ctx, const double *price_dev, int64_t n, int8_t *sides_dev) {
    if (n <= 0) return FMK_OK;
    return device_inclusive_scan<SideC>(ctx, SideIn{price_dev}, SideOut{sides_dev}, n, (SideC *)nullptr);
}

extern "C" int fmk_trade_side_vector(fmk_ctx *ctx, const double *prices, int64_t n, int8_t *sides_out) {
    FMK_ENTER(ctx);
    if (n <= 0) return FMK_OK;
    Scratch<double> p(ctx);
    Scratch<int8_t> s(ctx);
    FMK_TRY(p.alloc(n)); FMK_TRY(s.alloc(n));
    FMK_CUDA(ctx, cudaMemcpyAsync(p.p, prices, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_TRY((device_inclusive_scan<SideC>(ctx, SideIn{p.p}, SideOut{s.p}, n, (SideC *)nullptr)));
    FMK_CUDA(ctx, cudaMemcpyAsync(sides_out, s.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

// ---- merge_split_trades ------------------------------------------------------------------------------------------------
// A trade joins the current merged trade iff it has the HEAD's timestamp, a price within 1e-8 of the HEAD's price and
// the head's side.  A different timestamp always opens a new merged trade, so the stream splits into independent runs
// of equal timestamps; inside a run the head comparison is sequential (it is not transitive), and runs are short (the
// fills of one aggressive order).  M1: the first tick of every timestamp run walks its run and flags the group heads.
// M2: scan of the flags -> merged index of every tick.  M3: every group head writes ts / price / side and accumulates
// its members' amounts in float32, in arrival order, exactly like `merged_amounts[idx] += amounts[i]`.
__global__ void k_merge_flags(const int64_t *__restrict__ ts, const double *__restrict__ p, const uint8_t *__restrict__ ibm,
                              int64_t n, uint8_t *__restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t t0 = ts[i];
    if (i > 0 && ts[i - 1] == t0) return;          // not the first tick of its timestamp run
    double hp = p[i];
    int hs = ibm ? (ibm[i] != 0) : 0;
    head[i] = 1;
    for (int64_t j = i + 1; j < n && ts[j] == t0; j++) {
        bool same = fabs(__dadd_rn(p[j], -hp)) < 1e-8;
        if (ibm) same = same && ((ibm[j] != 0) == hs);
        if (same) head[j] = 0;
        else { head[j] = 1; hp = p[j]; hs = ibm ? (ibm[j] != 0) : 0; }
    }
}
struct HeadIn {
    const uint8_t *h;
    __device__ int64_t operator()(int64_t i) const { return h[i]; }
};
struct HeadOut {
    int64_t *gid;
    __device__ void operator()(int64_t i, int64_t incl) const { gid[i] = incl - 1; }
};
__global__ void k_merge_write(const int64_t *__restrict__ ts, const double *__restrict__ p, const float *__restrict__ a,
                              const uint8_t *__restrict__ ibm, const uint8_t *__restrict__ head,
                              const int64_t *__restrict__ gid, int64_t n, int64_t *__restrict__ ots,
                              double *__restrict__ op, float *__restrict__ oa, int8_t *__restrict__ oside) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const int64_t g = gid[i];
    float acc = a[i];
    for (int64_t j = i + 1; j < n && !head[j]; j++) acc = __fadd_rn(acc, a[j]);
    ots[g] = ts[i]; op[g] = p[i]; oa[g] = acc;
    if (ibm) oside[g] = ibm[i] ? -1 : 1;
}

extern "C" int fmk_merge_split_trades(fmk_ctx *ctx, const int64_t *ts, const double *prices, const float *amounts,
                                      const uint8_t *is_buyer_maker, int64_t n, int64_t *ts_out, double *prices_out,
                                      float *amounts_out, int8_t *sides_out, int64_t *n_out) {
    FMK_ENTER(ctx);
    *n_out = 0;
    if (n <= 0) return FMK_OK;
    Scratch<int64_t> dts(ctx), gid(ctx), ots(ctx), dtot(ctx);
    Scratch<double> dp(ctx), op(ctx);
    Scratch<float> da(ctx), oa(ctx);
    Scratch<uint8_t> dm(ctx), head(ctx);
    Scratch<int8_t> os(ctx);
    FMK_TRY(dts.alloc(n)); FMK_TRY(gid.alloc(n)); FMK_TRY(dp.alloc(n)); FMK_TRY(da.alloc(n)); FMK_TRY(head.alloc(n));
    FMK_TRY(dtot.alloc(1));
    FMK_CUDA(ctx, cudaMemcpyAsync(dts.p, ts, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(dp.p, prices, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(da.p, amounts, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (is_buyer_maker) {
        FMK_TRY(dm.alloc(n));
        FMK_CUDA(ctx, cudaMemcpyAsync(dm.p, is_buyer_maker, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    }
    const uint8_t *m = is_buyer_maker ? dm.p : nullptr;
    FMK_LAUNCH(ctx, k_merge_flags, (unsigned)cdiv(n, 256), 256, 0, (const int64_t *)dts.p, (const double *)dp.p, m, n, head.p);
    FMK_TRY((device_inclusive_scan<int64_t>(ctx, HeadIn{head.p}, HeadOut{gid.p}, n, dtot.p)));
    int64_t total = 0;
    FMK_CUDA(ctx, cudaMemcpyAsync(&total, dtot.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FMK_TRY(ots.alloc(total)); FMK_TRY(op.alloc(total)); FMK_TRY(oa.alloc(total)); FMK_TRY(os.alloc(total));
    FMK_LAUNCH(ctx, k_merge_write, (unsigned)cdiv(n, 256), 256, 0, (const int64_t *)dts.p, (const double *)dp.p,
               (const float *)da.p, m, (const uint8_t *)head.p, (const int64_t *)gid.p, n, ots.p, op.p, oa.p, os.p);
    FMK_CUDA(ctx, cudaMemcpyAsync(ts_out, ots.p, (size_t)total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(prices_out, op.p, (size_t)total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaMemcpyAsync(amounts_out, oa.p, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (is_buyer_maker && sides_out)
        FMK_CUDA(ctx, cudaMemcpyAsync(sides_out, os.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = total;
    return FMK_OK;
}
