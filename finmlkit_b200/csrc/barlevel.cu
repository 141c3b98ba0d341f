// barlevel.cu -- bar-level volatility / order-flow features: realized_vol (feature/core/volatility.py:
// 256-286), ewms (volatility.py:9-69), vpin (feature/core/volume.py:610-641), comp_flow_acceleration (volume.py:572-607).
// These run on n_bars-length arrays (1e3..1e6 elements): small kernels built on the generic scan (scan.cuh).
#include <math.h>
#include "common.cuh"
#include "scan.cuh"

__device__ __forceinline__ double bl_nan() { return __longlong_as_double(0x7ff8000000000000ll); }

// realized_vol: one thread per output; the window is summed in index order like np.nansum(r_window ** 2)
__global__ void k_realized_vol(const double *__restrict__ r, int64_t n, int64_t window, int is_sample,
                               double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o = bl_nan();
    if (window >= 1 && i >= window - 1) {
        int64_t valid = 0;
        double s = 0.0;
        for (int64_t j = i - window + 1; j <= i; j++) {
            const double x = r[j];
            if (x == x) { valid++; s = __dadd_rn(s, __dmul_rn(x, x)); }
        }
        if (valid > 1) o = __dsqrt_rn(__ddiv_rn(s, (double)(is_sample ? valid - 1 : valid)));
    }
    out[i] = o;
}

// ewms: S' = om*S + b  for (S_w, S_y, S_y2) and S_w2' = om^2*S_w2 + b  -> affine maps, scanned
struct EA {
    double A, A2, bw, bw2, by, by2;
    __device__ EA() {}
    __device__ explicit EA(int) { A = 1.0; A2 = 1.0; bw = bw2 = by = by2 = 0.0; }
};
__device__ __forceinline__ EA operator+(const EA &f, const EA &g) {   // f first, then g
    EA h;
    h.A = g.A * f.A; h.A2 = g.A2 * f.A2;
    h.bw = g.A * f.bw + g.bw; h.bw2 = g.A2 * f.bw2 + g.bw2;
    h.by = g.A * f.by + g.by; h.by2 = g.A * f.by2 + g.by2;
    return h;
}
__device__ __forceinline__ EA __shfl_up_sync(unsigned m, const EA &x, int o) {
    EA r;
    r.A = ::__shfl_up_sync(m, x.A, o); r.A2 = ::__shfl_up_sync(m, x.A2, o);
    r.bw = ::__shfl_up_sync(m, x.bw, o); r.bw2 = ::__shfl_up_sync(m, x.bw2, o);
    r.by = ::__shfl_up_sync(m, x.by, o); r.by2 = ::__shfl_up_sync(m, x.by2, o);
    return r;
}
struct EwmsIn {
    const double *y;
    double om;
    __device__ EA operator()(int64_t t) const {
        const double v = y[t];
        const bool nan = v != v;
        EA e;
        e.A = om; e.A2 = om * om;
        e.bw = nan ? 0.0 : 1.0; e.bw2 = e.bw;
        e.by = nan ? 0.0 : v; e.by2 = nan ? 0.0 : v * v;
        return e;
    }
};
struct EwmsOut {
    double *out;
    __device__ void operator()(int64_t t, const EA &s) const {   // state after tick t = offsets of the inclusive map
        double o = bl_nan();
        const double Sw = s.bw, Sw2 = s.bw2, Sy = s.by, Sy2 = s.by2;
        if (Sw > 0.0) {
            const double mean = __ddiv_rn(Sy, Sw);
            const double den = Sw - __ddiv_rn(Sw2, Sw);
            if (den > 0.0) {
                double var = __ddiv_rn((__ddiv_rn(Sy2, Sw) - mean * mean) * Sw, den);
                if (!(var > 0.0)) var = (var != var) ? var : 0.0;
                o = __dsqrt_rn(var);
            }
        }
        out[t] = o;
    }
};

// prefix sums for vpin / flow acceleration: P[i+1] = P[i] + x_i, P[0] = 0
struct PreIn {
    const double *a, *b;
    int mode;   // 0: a ; 1: b ; 2: |a-b| ; 3: nan flag ; values with a NaN partner count as 0 ; 4: a as is (NaN propagates)
    __device__ double operator()(int64_t i) const {
        const double x = a[i], y = b ? b[i] : 0.0;
        if (mode == 4) return x;
        const bool nan = (x != x) || (y != y);
        if (mode == 3) return nan ? 1.0 : 0.0;
        if (nan) return 0.0;
        return mode == 0 ? x : (mode == 1 ? y : fabs(x - y));
    }
};
struct PreOut {
    double *P;
    __device__ void operator()(int64_t i, double cs) const { P[i + 1] = cs; if (i == 0) P[0] = 0.0; }
};

__global__ void k_vpin_out(const double *__restrict__ bc, const double *__restrict__ sc, const double *__restrict__ ac,
                           const double *__restrict__ nf, int64_t n, int64_t window, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float o = __int_as_float(0x7fc00000);
    if (i >= window - 1 && nf[i + 1] - nf[i + 1 - window] == 0.0) {
        const double tot = (bc[i + 1] - bc[i + 1 - window]) + (sc[i + 1] - sc[i + 1 - window]);
        if (tot > 1e-9) o = (float)__ddiv_rn(ac[i + 1] - ac[i + 1 - window], tot);
    }
    out[i] = o;
}

__global__ void k_flow_out(const double *__restrict__ S, int64_t n, int64_t window, int64_t recent, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o = bl_nan();
    if (n >= window && recent < window && i >= window - 1) {
        const double rs = S[i + 1] - S[i + 1 - recent], ps = S[i + 1 - recent] - S[i + 1 - window];
        o = log(__ddiv_rn(rs + 1e-12, ps + 1e-12));
    }
    out[i] = o;
}

template <typename T>
static int up(fmk_ctx *ctx, T *dev, const T *host, int64_t n) {
    if (n > 0) FMK_CUDA(ctx, cudaMemcpyAsync(dev, host, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return FMK_OK;
}
template <typename T>
static int down(fmk_ctx *ctx, T *host, const T *dev, int64_t n) {
    if (n > 0) FMK_CUDA(ctx, cudaMemcpyAsync(host, dev, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    FMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMK_OK;
}

extern "C" int fmk_realized_vol(fmk_ctx *ctx, const double *r, int64_t n, int64_t window, int is_sample, double *out) {
    FMK_ENTER(ctx);
    Scratch<double> dr(ctx), dout(ctx);
    FMK_TRY(dr.alloc(n)); FMK_TRY(dout.alloc(n));
    FMK_TRY(up(ctx, dr.p, r, n));
    if (n > 0) FMK_LAUNCH(ctx, k_realized_vol, (unsigned)cdiv(n, 128), 128, 0, (const double *)dr.p, n, window, is_sample, dout.p);
    return down(ctx, out, dout.p, n);
}

extern "C" int fmk_ewms(fmk_ctx *ctx, const double *y, int64_t n, int64_t span, double *out) {
    FMK_ENTER(ctx);
    if (span <= 1) {   // volatility.py:27-30: all NaN
        for (int64_t i = 0; i < n; i++) out[i] = NAN;
        return FMK_OK;
    }
    Scratch<double> dy(ctx), dout(ctx);
    FMK_TRY(dy.alloc(n)); FMK_TRY(dout.alloc(n));
    FMK_TRY(up(ctx, dy.p, y, n));
    const double alpha = 2.0 / ((double)span + 1.0);
    FMK_TRY((device_inclusive_scan<EA>(ctx, EwmsIn{dy.p, 1.0 - alpha}, EwmsOut{dout.p}, n, (EA *)nullptr)));
    return down(ctx, out, dout.p, n);
}

extern "C" int fmk_vpin(fmk_ctx *ctx, const double *volume_buy, const double *volume_sell, int64_t n, int64_t window,
                        float *out) {
    FMK_ENTER(ctx);
    if (window < 1) return fmk_fail(ctx, FMK_ERR_ARG, "window must be positive");
    Scratch<double> db(ctx), ds(ctx), P(ctx);
    Scratch<float> dout(ctx);
    FMK_TRY(db.alloc(n)); FMK_TRY(ds.alloc(n)); FMK_TRY(P.alloc(4 * (n + 1))); FMK_TRY(dout.alloc(n));
    FMK_TRY(up(ctx, db.p, volume_buy, n)); FMK_TRY(up(ctx, ds.p, volume_sell, n));
    for (int m = 0; m < 4; m++)
        FMK_TRY((device_inclusive_scan<double>(ctx, PreIn{db.p, ds.p, m}, PreOut{P.p + m * (n + 1)}, n, (double *)nullptr)));
    if (n > 0)
        FMK_LAUNCH(ctx, k_vpin_out, (unsigned)cdiv(n, 128), 128, 0, (const double *)P.p, (const double *)(P.p + (n + 1)),
                   (const double *)(P.p + 2 * (n + 1)), (const double *)(P.p + 3 * (n + 1)), n, window, dout.p);
    return down(ctx, out, dout.p, n);
}

extern "C" int fmk_flow_acceleration(fmk_ctx *ctx, const double *volumes, int64_t n, int64_t window, int64_t recent,
                                     double *out) {
    FMK_ENTER(ctx);
    Scratch<double> dv(ctx), S(ctx), dout(ctx);
    FMK_TRY(dv.alloc(n)); FMK_TRY(S.alloc(n + 1)); FMK_TRY(dout.alloc(n));
    FMK_TRY(up(ctx, dv.p, volumes, n));
    // volume.py:590-593 builds S[i+1] = S[i] + volumes[i] with no NaN handling: a NaN volume poisons every later sum
    FMK_TRY((device_inclusive_scan<double>(ctx, PreIn{dv.p, nullptr, 4}, PreOut{S.p}, n, (double *)nullptr)));
    if (n > 0) FMK_LAUNCH(ctx, k_flow_out, (unsigned)cdiv(n, 128), 128, 0, (const double *)S.p, n, window, recent, dout.p);
    return down(ctx, out, dout.p, n);
}
