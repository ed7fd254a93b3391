// mcig_kernels.cuh — ahead-of-time sm_100a kernels of the engine that do not depend on user plugins:
//   K2  combine_walkers      (replaces MPIMCI::integrate's reduce, src/MPIMCI.cpp:85-92)
//   K4  est_uncorrelated     (MultiDim/OneDimUncorrelatedEstimator, src/Estimators.cpp:36-56, 125-155)
//   K3  mj_segments + mj_finish (MJBlocker::estimate, src/MJBlocker.cpp:46-154) — one streaming read of the series
//   K5  est_fcblocker        (FCBlockerEstimator, src/Estimators.cpp:82-122, 191-246) — single pass, 45 partitions at once
//   finalize_simple / noop   (SimpleAccumulator::_finalize src/SimpleAccumulator.cpp:22-28, NoopEstimator :288-293)
//   reduce_u64               (acceptance counters, src/MCIntegrator.cpp:587-592)
//   dfma_peak / imad_peak    (issue-rate microbenchmarks: the FP64 roofline denominator, measured live by bench.py)
//
// Sample storage layout everywhere: data[(i*nobs + j)*W + w]  (store index i, observable component j, walker w):
// a warp reads/writes 32 consecutive walkers = one 256 B segment.
//
// Estimator kernels avoid FMA contraction on purpose (explicit __dmul_rn/__dadd_rn): with one thread per chain they
// then reproduce the reference's left-to-right sums bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcig_k {

typedef long long i64;
typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------- small utilities
__global__ void reduce_u64_kernel(const u64 * __restrict__ in, i64 n, u64 * __restrict__ out)
{ // single block, deterministic
    __shared__ u64 sm[32];
    u64 s = 0;
    i64 i = threadIdx.x;
    for (; i + 7*(i64)blockDim.x < n; i += 8*(i64)blockDim.x) { // 8 loads in flight per thread (a single block: latency, not bandwidth)
        u64 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { v[u] = in[i + u*(i64)blockDim.x]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s += v[u]; }
    }
    for (; i < n; i += blockDim.x) { s += in[i]; }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); }
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x < 32) {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); }
        if (threadIdx.x == 0) { out[0] = s; }
    }
}

// Simple accumulator finalize + Noop estimator: avg_w = sum_w * (1./NAccu), err_w = 0
__global__ void finalize_simple_kernel(const double * __restrict__ sums, i64 n /*nobs*W*/, double normf, double * __restrict__ wavg,
                                       double * __restrict__ werr)
{
    const i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) {
        wavg[i] = __dmul_rn(sums[i], normf);
        werr[i] = 0.;
    }
}

// Noop estimator on stored data (Block/Full): "average" = first stored sample, error 0 (src/Estimators.cpp:288-293)
__global__ void noop_first_kernel(const double * __restrict__ data, i64 n /*nobs*W*/, double * __restrict__ wavg, double * __restrict__ werr)
{
    const i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) {
        wavg[i] = data[i];
        werr[i] = 0.;
    }
}

// ---------------------------------------------------------------------------------------------- K2 combine over walkers
// out[3*nobsdim]: sum_w avg_w | sum_w err_w^2 | sum_w avg_w^2 . One block per observable component, fixed tree => deterministic.
__global__ void combine_walkers_kernel(const double * __restrict__ wavg, const double * __restrict__ werr, i64 W, int nobsdim,
                                       double * __restrict__ out)
{
    const int j = blockIdx.x;
    __shared__ double sm[3][32];
    double a = 0., e = 0., q = 0.;
    i64 w = threadIdx.x;
    for (; w + 3*(i64)blockDim.x < W; w += 4*(i64)blockDim.x) { // 8 loads in flight per thread, added in a fixed order
        double v[4], r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = wavg[(i64)j*W + w + u*(i64)blockDim.x];
            r[u] = werr[(i64)j*W + w + u*(i64)blockDim.x];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a += v[u];
            e += r[u]*r[u];
            q += v[u]*v[u];
        }
    }
    for (; w < W; w += blockDim.x) {
        const double v = wavg[(i64)j*W + w], r = werr[(i64)j*W + w];
        a += v;
        e += r*r;
        q += v*v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        e += __shfl_down_sync(0xffffffffu, e, o);
        q += __shfl_down_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sm[0][threadIdx.x >> 5] = a;
        sm[1][threadIdx.x >> 5] = e;
        sm[2][threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const bool in = threadIdx.x < (blockDim.x >> 5);
        a = in ? sm[0][threadIdx.x] : 0.;
        e = in ? sm[1][threadIdx.x] : 0.;
        q = in ? sm[2][threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            e += __shfl_down_sync(0xffffffffu, e, o);
            q += __shfl_down_sync(0xffffffffu, q, o);
        }
        if (threadIdx.x == 0) {
            out[j] = a;
            out[nobsdim + j] = e;
            out[2*nobsdim + j] = q;
        }
    }
}

// ---------------------------------------------------------------------------------------------- K4 uncorrelated estimator
// Partial sums over a time segment: thread = (segment, column) with column = j*W + w. part[(seg*2 + {0,1})*ncol + col].
__global__ void uncorr_partial_kernel(const double * __restrict__ data, i64 n, i64 ncol, int nseg, double * __restrict__ part)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    const int seg = blockIdx.y;
    if (col >= ncol) { return; }
    const i64 i0 = n*seg/nseg, i1 = n*(seg + 1)/nseg;
    double s = 0., q = 0.;
    i64 i = i0;
    for (; i + 16 <= i1; i += 16) { // 16 independent loads in flight per thread (HBM latency), sums stay in sample order
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) { v[u] = __ldcs(data + (i + u)*ncol + col); }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            s = __dadd_rn(s, v[u]);
            q = __dadd_rn(q, __dmul_rn(v[u], v[u]));
        }
    }
    for (; i < i1; ++i) {
        const double v = __ldcs(data + i*ncol + col);
        s = __dadd_rn(s, v);
        q = __dadd_rn(q, __dmul_rn(v, v));
    }
    part[((i64)seg*2 + 0)*ncol + col] = s;
    part[((i64)seg*2 + 1)*ncol + col] = q;
}

// column sums from the per-segment partials (sum part only), segments added in order
__global__ void sum_partials_kernel(const double * __restrict__ part, i64 ncol, int nseg, double * __restrict__ sums)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    double s = 0.;
    for (int seg = 0; seg < nseg; ++seg) { s = __dadd_rn(s, part[((i64)seg*2 + 0)*ncol + col]); }
    sums[col] = s;
}

// nobs_is_one selects the 1-D formula sqrt(var/(n-1.)) vs the N-D one sqrt(var*(1./(n-1.))) (Estimators.cpp:49-50 vs :147-150)
__global__ void uncorr_finish_kernel(const double * __restrict__ part, i64 n, i64 ncol, int nseg, int nobs_is_one, double * __restrict__ wavg,
                                     double * __restrict__ werr)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    double s = 0., q = 0.;
    for (int seg = 0; seg < nseg; ++seg) {
        s = __dadd_rn(s, part[((i64)seg*2 + 0)*ncol + col]);
        q = __dadd_rn(q, part[((i64)seg*2 + 1)*ncol + col]);
    }
    const double norm = 1./(double)n;
    const double avg = __dmul_rn(s, norm);
    double er = __dadd_rn(__dmul_rn(q, norm), -__dmul_rn(avg, avg));
    if (er > 1.e-300) { er = nobs_is_one ? sqrt(er/((double)n - 1.)) : sqrt(__dmul_rn(er, 1./((double)n - 1.))); }
    else { er = 0.; }
    wavg[col] = avg;
    werr[col] = er;
}

// The same estimator when the walk kernel already accumulated sum x and sum x^2 of the stored values in its registers (FUSE in
// device/mcig_device.cuh): one segment, the reference's left-to-right order, nothing read from the series.
__global__ void uncorr_from_sums_kernel(const double * __restrict__ sums, const double * __restrict__ sqs, i64 n, i64 ncol, int nobs_is_one,
                                        double * __restrict__ wavg, double * __restrict__ werr)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    const double norm = 1./(double)n;
    const double avg = __dmul_rn(sums[col], norm);
    double er = __dadd_rn(__dmul_rn(sqs[col], norm), -__dmul_rn(avg, avg));
    if (er > 1.e-300) { er = nobs_is_one ? sqrt(er/((double)n - 1.)) : sqrt(__dmul_rn(er, 1./((double)n - 1.))); }
    else { er = 0.; }
    wavg[col] = avg;
    werr[col] = er;
}

// ---------------------------------------------------------------------------------------------- K3 MJBlocker
// Streaming blocking pyramid over one aligned segment of L = 2^m samples of one chain (thread = (segment, column)).
// For every level k < m it accumulates sum X^2 and sum X_i X_{i+1} inside the segment and remembers the first and
// last centred element (for the products that straddle segment borders); the segment's level-m element (its mean)
// goes to `top`. All levels come from ONE read of the series; the reference makes ~13 passes (src/MJBlocker.cpp:127-154).
#define MCIG_MJ_MAXLEV 40
struct MJLevel {
    double pend;   // first element of an incomplete pair (uncentred)
    double prevX;  // previous centred element of this level
    double sq;     // sum X^2
    double cr;     // sum X_i X_{i+1}
    double firstX; // first centred element
};

__device__ __forceinline__ void mj_push(MJLevel * lv, unsigned long long & havemask, unsigned long long & cntmask, int m, double x0, double mean,
                                        double & top)
{ // insert one level-0 sample; carries propagate upwards like a binary counter
    double x = x0;
    for (int k = 0; k <= m; ++k) {
        if (k == m) { top = x; return; }
        const double X = __dadd_rn(x, -mean);
        MJLevel & L = lv[k];
        L.sq = __dadd_rn(L.sq, __dmul_rn(X, X));
        if (havemask >> k & 1ULL) { L.cr = __dadd_rn(L.cr, __dmul_rn(L.prevX, X)); }
        else { L.firstX = X; havemask |= 1ULL << k; }
        L.prevX = X;
        if (cntmask >> k & 1ULL) { // completes a pair
            cntmask &= ~(1ULL << k);
            x = __dmul_rn(0.5, __dadd_rn(L.pend, x));
        }
        else {
            cntmask |= 1ULL << k;
            L.pend = x;
            return;
        }
    }
}

// seg_out layout: [(seg*m + k)*4 + {sq,cr,firstX,lastX}]*ncol + col ; top[seg*ncol + col]
__global__ void mj_segments_kernel(const double * __restrict__ data, i64 ncol, i64 L, int m, const double * __restrict__ mean /*[ncol]*/,
                                   double * __restrict__ seg_out, double * __restrict__ top)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    const i64 seg = blockIdx.y;
    if (col >= ncol) { return; }
    MJLevel lv[MCIG_MJ_MAXLEV];
    for (int k = 0; k < m; ++k) { lv[k].sq = 0.; lv[k].cr = 0.; lv[k].firstX = 0.; lv[k].prevX = 0.; lv[k].pend = 0.; }
    unsigned long long have = 0, cnt = 0;
    const double mu = mean[col];
    double t = 0.;
    const double * src = data + seg*L*ncol + col;
    for (i64 i = 0; i < L; ++i) { mj_push(lv, have, cnt, m, __ldcs(src + i*ncol), mu, t); }
    for (int k = 0; k < m; ++k) {
        double * o = seg_out + ((seg*m + k)*4)*ncol + col;
        o[0] = lv[k].sq;
        o[ncol] = lv[k].cr;
        o[2*ncol] = lv[k].firstX;
        o[3*ncol] = lv[k].prevX;
    }
    top[seg*ncol + col] = t;
}


// Register-tiled variant (used when the segment has at least 16 samples): 16 consecutive samples are loaded with 16
// independent strided loads (memory-level parallelism) and levels 0..3 are reduced as a fully unrolled tree in registers;
// only the level-4 element (1/16 of the samples) enters the dynamic pyramid. Accumulation order per level is ascending in
// the sample index, exactly as in the generic kernel and in the reference.
#define MCIG_MJ_TILE_LOG 4
#ifndef MCIG_MJ_MID
#define MCIG_MJ_MID 2 // levels T .. T+MID-1 also live in registers (driven by the bits of the tile index); the dynamic pyramid sees 1/2^(T+MID) of the samples
#endif
__global__ void __launch_bounds__(128) mj_segments_tiled_kernel(const double * __restrict__ data, i64 ncol, i64 L, int m, const double * __restrict__ mean,
                                                               double * __restrict__ seg_out, double * __restrict__ top)
{
    constexpr int T = MCIG_MJ_TILE_LOG, TS = 1 << T, S = MCIG_MJ_MID;
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    const i64 seg = blockIdx.y;
    if (col >= ncol) { return; }
    const double mu = mean[col];
    double sq[T], cr[T], prevX[T];
#pragma unroll
    for (int k = 0; k < T; ++k) { sq[k] = 0.; cr[k] = 0.; prevX[k] = 0.; }
    // middle levels: the level-T element of tile t reaches level T+k iff the low k bits of t are all ones, and completes a pair there iff bit k
    // is set too: warp-uniform branches on the tile index, statically named registers (the dynamically indexed pyramid below lives in local memory:
    // with every tile pushing into it the kernel moved as many bytes through L1 as it streamed from HBM)
    const int ms = (m - T < S) ? (m - T) : S; // middle levels in use
    double msq[S > 0 ? S : 1], mcr[S > 0 ? S : 1], mprev[S > 0 ? S : 1], mpend[S > 0 ? S : 1];
#pragma unroll
    for (int k = 0; k < S; ++k) { msq[k] = 0.; mcr[k] = 0.; mprev[k] = 0.; mpend[k] = 0.; }
    MJLevel lv[MCIG_MJ_MAXLEV];
    const int mu_lev = m - T - ms; // levels handled by the dynamic pyramid
    for (int k = 0; k < mu_lev; ++k) { lv[k].sq = 0.; lv[k].cr = 0.; lv[k].firstX = 0.; lv[k].prevX = 0.; lv[k].pend = 0.; }
    unsigned long long have = 0, cnt = 0;
    double t = 0.;
    const double * src = data + seg*L*ncol + col;
    double * const o_first = seg_out + (seg*m*4 + 2)*ncol + col; // firstX of level k at o_first[k*4*ncol]: written when it occurs, not kept in registers
    const i64 ntiles = L >> T;
    double nxt[TS]; // software prefetch: the loads of tile t+1 are in flight while tile t runs through its ~150 dependent FP64 operations
#pragma unroll
    for (int i = 0; i < TS; ++i) { nxt[i] = __ldcs(src + (i64)i*ncol); }
    for (i64 tile = 0; tile < ntiles; ++tile) {
        double v[TS];
#pragma unroll
        for (int i = 0; i < TS; ++i) { v[i] = nxt[i]; }
        if (tile + 1 < ntiles) {
#pragma unroll
            for (int i = 0; i < TS; ++i) { nxt[i] = __ldcs(src + ((tile + 1)*TS + i)*ncol); }
        }
#pragma unroll
        for (int k = 0; k < T; ++k) {
            // level k holds TS >> k elements in v[0..)
            // sums of squares / lagged products with fused multiply-adds: one rounding per term instead of two (the reference's
            // separately rounded products differ by <= 1 ulp per term, far inside the 1e-9 estimator tolerance) and 40 % fewer
            // FP64 instructions per tile
            double X0 = __dadd_rn(v[0], -mu);
            if (tile == 0) { o_first[(i64)k*4*ncol] = X0; }
            else { cr[k] = fma(prevX[k], X0, cr[k]); }
            sq[k] = fma(X0, X0, sq[k]);
            double Xp = X0;
#pragma unroll
            for (int i = 1; i < (TS >> k); ++i) {
                const double X = __dadd_rn(v[i], -mu);
                sq[k] = fma(X, X, sq[k]);
                cr[k] = fma(Xp, X, cr[k]);
                Xp = X;
            }
            prevX[k] = Xp;
#pragma unroll
            for (int i = 0; i < (TS >> (k + 1)); ++i) { v[i] = __dmul_rn(0.5, __dadd_rn(v[2*i], v[2*i + 1])); }
        }
        double x = v[0];
        bool up = true; // x is an element of the next level
#pragma unroll
        for (int k = 0; k < S; ++k) {
            if (up && k < ms) {
                const double X = __dadd_rn(x, -mu);
                msq[k] = fma(X, X, msq[k]);
                if (tile >= ((i64)1 << k)) { mcr[k] = fma(mprev[k], X, mcr[k]); }
                else { o_first[(i64)(T + k)*4*ncol] = X; }
                mprev[k] = X;
                if ((tile >> k) & 1) { x = __dmul_rn(0.5, __dadd_rn(mpend[k], x)); }
                else {
                    mpend[k] = x;
                    up = false;
                }
            }
        }
        if (up) { mj_push(lv, have, cnt, mu_lev, x, mu, t); }
    }
#pragma unroll
    for (int k = 0; k < T; ++k) {
        double * o = seg_out + ((seg*m + k)*4)*ncol + col;
        o[0] = sq[k];
        o[ncol] = cr[k];
        o[3*ncol] = prevX[k];
    }
#pragma unroll
    for (int k = 0; k < S; ++k) {
        if (k < ms) {
            double * o = seg_out + ((seg*m + T + k)*4)*ncol + col;
            o[0] = msq[k];
            o[ncol] = mcr[k];
            o[3*ncol] = mprev[k];
        }
    }
    for (int k = 0; k < mu_lev; ++k) {
        double * o = seg_out + ((seg*m + T + ms + k)*4)*ncol + col;
        o[0] = lv[k].sq;
        o[ncol] = lv[k].cr;
        o[2*ncol] = lv[k].firstX;
        o[3*ncol] = lv[k].prevX;
    }
    top[seg*ncol + col] = t;
}

__constant__ double c_mj_quantile[64] = {
    3.841459, 5.991465, 7.814728, 9.487729, 11.070498, 12.591587, 14.067140, 15.507313, 16.918978, 18.307038, 19.675138,
    21.026070, 22.362032, 23.684791, 24.995790, 26.296228, 27.587112, 28.869299, 30.143527, 31.410433, 32.670573, 33.924438,
    35.172462, 36.415029, 37.652484, 38.885139, 40.113272, 41.337138, 42.556968, 43.772972, 44.985343, 46.194260, 47.399884,
    48.602367, 49.801850, 50.998460, 52.192320, 53.383541, 54.572228, 55.758479, 56.942387, 58.124038, 59.303512, 60.480887,
    61.656233, 62.829620, 64.001112, 65.170769, 66.338649, 67.504807, 68.669294, 69.832160, 70.993453, 72.153216, 73.311493,
    74.468324, 75.623748, 76.777803, 77.930524, 79.081944, 80.232098, 81.381015, 82.528727, 83.675261};

// One thread per (chain, level < m): merges the per-segment statistics of that level in segment order -> lvl[(2k + {0,1})*ncol + col] = var_k, gamma_k
__global__ void mj_merge_kernel(const double * __restrict__ seg_out, i64 ncol, i64 n, i64 nseg, int m, double * __restrict__ lvl)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    const int k = (int)blockIdx.y;
    double sq = 0., cr = 0., lastX = 0.;
    for (i64 s = 0; s < nseg; ++s) {
        const double * o = seg_out + ((s*m + k)*4)*ncol + col;
        const double o0 = __ldcs(o), o1 = __ldcs(o + ncol), o2 = __ldcs(o + 2*ncol), o3 = __ldcs(o + 3*ncol);
        if (s > 0) { cr = __dadd_rn(cr, __dmul_rn(lastX, o2)); } // product across the segment border
        sq = __dadd_rn(sq, o0);
        cr = __dadd_rn(cr, o1);
        lastX = o3;
    }
    const double nred = (double)(n >> k);
    lvl[(i64)(2*k)*ncol + col] = sq/nred;
    lvl[(i64)(2*k + 1)*ncol + col] = cr/nred;
}

// One thread per chain: levels < m from mj_merge_kernel, the pyramid over the nseg segment tops (levels >= m), then the test
// statistic / level choice of src/MJBlocker.cpp:106-152.
__global__ void mj_finish_kernel(const double * __restrict__ lvl, const double * __restrict__ top, i64 ncol, i64 n, i64 nseg, int m, int npow,
                                 const double * __restrict__ mean, double * __restrict__ wavg, double * __restrict__ werr)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    double var[MCIG_MJ_MAXLEV], gam[MCIG_MJ_MAXLEV];
    const double mu = mean[col];
    for (int k = 0; k < m; ++k) {
        var[k] = lvl[(i64)(2*k)*ncol + col];
        gam[k] = lvl[(i64)(2*k + 1)*ncol + col];
    }
    {
        MJLevel lv[MCIG_MJ_MAXLEV];
        const int mt = npow - m; // levels held by the tops
        for (int k = 0; k < mt; ++k) { lv[k].sq = 0.; lv[k].cr = 0.; lv[k].firstX = 0.; lv[k].prevX = 0.; lv[k].pend = 0.; }
        unsigned long long have = 0, cnt = 0;
        double t = 0.;
        for (i64 s = 0; s < nseg; ++s) { mj_push(lv, have, cnt, mt, top[s*ncol + col], mu, t); }
        for (int k = 0; k < mt; ++k) {
            const double nred = (double)(n >> (m + k));
            var[m + k] = lv[k].sq/nred;
            gam[m + k] = lv[k].cr/nred;
        }
    }
    // _generateM + cumulative sums (src/MJBlocker.cpp:106-123)
    double M[MCIG_MJ_MAXLEV];
    for (int i = 0; i < npow; ++i) {
        const double q = gam[i]/var[i];
        M[npow - i - 1] = __dmul_rn(__dmul_rn(q, q), exp2((double)(npow - i)));
    }
    int kk = -1;
    {
        double Msum[MCIG_MJ_MAXLEV];
        double run = 0.;
        for (int i = 0; i < npow; ++i) {
            run = __dadd_rn(run, M[i]);
            Msum[i] = run;
        }
        for (kk = npow - 1; kk >= 0; --kk) {
            if (Msum[kk] < c_mj_quantile[kk]) { break; }
        }
    }
    const int k = npow - (kk + 1);
    wavg[col] = mu;
    // no level passes only when the variance is 0 (NaN statistics): the reference then reads out of bounds
    // (SURVEY.md Appendix C #12); the error of a constant series is defined as 0 here
    werr[col] = (k >= npow) ? 0. : sqrt(var[k]/exp2((double)(npow - k)));
}

// ---- chunked staging (a series longer than HBM is sampled, staged and folded chunk by chunk; include/mci/FullAccumulator.hpp:11-13 warns
// about exactly this memory need). The blocker centres every level on the GLOBAL mean (src/MJBlocker.cpp:70-103), which is only known after the
// last chunk; the chunks are therefore centred on a provisional mean mu0 (the mean of the first chunk) and the level statistics corrected
// exactly at the end: with X = x - mu0, Y = x - mu = X - d (d = mu - mu0) and sum_i X_i = n_k d on every level,
//     sum Y_i^2       = sum X_i^2 - n_k d^2
//     sum Y_i Y_{i+1} = sum X_i X_{i+1} - (n_k + 1) d^2 + d (X_first + X_last).
// d is of the order sigma/sqrt(chunk length): the correction terms are far below the rounding of the sums (errors stay inside 1e-9).

// One thread per (chain, level < m): folds the segments of ONE chunk into the running statistics of that level, in time order.
// acc[(k*4 + {sq, cr, firstX (of the whole series), lastX})*ncol + col]
__global__ void mj_merge_accum_kernel(const double * __restrict__ seg_out, i64 ncol, i64 nseg, int m, int first_chunk, double * __restrict__ acc)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    const int k = (int)blockIdx.y;
    double * a = acc + (i64)(k*4)*ncol + col;
    double sq = 0., cr = 0., firstX = 0., lastX = 0.;
    if (!first_chunk) {
        sq = a[0];
        cr = a[ncol];
        firstX = a[2*ncol];
        lastX = a[3*ncol];
    }
    for (i64 s = 0; s < nseg; ++s) {
        const double * o = seg_out + ((s*m + k)*4)*ncol + col;
        const double o0 = __ldcs(o), o1 = __ldcs(o + ncol), o2 = __ldcs(o + 2*ncol), o3 = __ldcs(o + 3*ncol);
        if (first_chunk && s == 0) { firstX = o2; }
        else { cr = __dadd_rn(cr, __dmul_rn(lastX, o2)); } // product across the segment / chunk border
        sq = __dadd_rn(sq, o0);
        cr = __dadd_rn(cr, o1);
        lastX = o3;
    }
    a[0] = sq;
    a[ncol] = cr;
    a[2*ncol] = firstX;
    a[3*ncol] = lastX;
}

// One thread per chain: levels < m from the folded statistics (with the mean correction above), levels >= m from the pyramid over all
// segment tops of all chunks (uncentred segment means: centred on the true mean here), then the level choice of src/MJBlocker.cpp:106-152.
__global__ void mj_finish_shifted_kernel(const double * __restrict__ acc, const double * __restrict__ top, i64 ncol, i64 n, i64 nseg, int m, int npow,
                                         const double * __restrict__ mu0, const double * __restrict__ sums, double * __restrict__ wavg, double * __restrict__ werr)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    double var[MCIG_MJ_MAXLEV], gam[MCIG_MJ_MAXLEV];
    const double mu = sums[col]/(double)n; // src/MJBlocker.cpp:46-55: input-order sum, then one division
    const double d = __dadd_rn(mu, -mu0[col]);
    const double d2 = __dmul_rn(d, d);
    for (int k = 0; k < m; ++k) {
        const double * a = acc + (i64)(k*4)*ncol + col;
        const double nred = (double)(n >> k);
        var[k] = __dadd_rn(a[0], -__dmul_rn(nred, d2))/nred;
        gam[k] = __dadd_rn(__dadd_rn(a[ncol], -__dmul_rn(nred + 1., d2)), __dmul_rn(d, __dadd_rn(a[2*ncol], a[3*ncol])))/nred;
    }
    {
        MJLevel lv[MCIG_MJ_MAXLEV];
        const int mt = npow - m;
        for (int k = 0; k < mt; ++k) { lv[k].sq = 0.; lv[k].cr = 0.; lv[k].firstX = 0.; lv[k].prevX = 0.; lv[k].pend = 0.; }
        unsigned long long have = 0, cnt = 0;
        double t = 0.;
        for (i64 s = 0; s < nseg; ++s) { mj_push(lv, have, cnt, mt, __ldcs(top + s*ncol + col), mu, t); }
        for (int k = 0; k < mt; ++k) {
            const double nred = (double)(n >> (m + k));
            var[m + k] = lv[k].sq/nred;
            gam[m + k] = lv[k].cr/nred;
        }
    }
    double M[MCIG_MJ_MAXLEV];
    for (int i = 0; i < npow; ++i) {
        const double q = gam[i]/var[i];
        M[npow - i - 1] = __dmul_rn(__dmul_rn(q, q), exp2((double)(npow - i)));
    }
    int kk = -1;
    {
        double Msum[MCIG_MJ_MAXLEV];
        double run = 0.;
        for (int i = 0; i < npow; ++i) {
            run = __dadd_rn(run, M[i]);
            Msum[i] = run;
        }
        for (kk = npow - 1; kk >= 0; --kk) {
            if (Msum[kk] < c_mj_quantile[kk]) { break; }
        }
    }
    const int k = npow - (kk + 1);
    wavg[col] = mu;
    werr[col] = (k >= npow) ? 0. : sqrt(var[k]/exp2((double)(npow - k)));
}

// running [sum x | sum x^2] of a chain over the chunks, from one chunk's per-segment partials (uncorrelated estimator, many components)
__global__ void uncorr_accum_kernel(const double * __restrict__ part, i64 ncol, int nseg, int first_chunk, double * __restrict__ acc /*[2][ncol]*/)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    double s = first_chunk ? 0. : acc[col], q = first_chunk ? 0. : acc[ncol + col];
    for (int seg = 0; seg < nseg; ++seg) {
        s = __dadd_rn(s, part[((i64)seg*2 + 0)*ncol + col]);
        q = __dadd_rn(q, part[((i64)seg*2 + 1)*ncol + col]);
    }
    acc[col] = s;
    acc[ncol + col] = q;
}

// mean of stored data per column from the in-kernel running sums: mean = sum / n  (src/MJBlocker.cpp:46-55 divides)
__global__ void mean_from_sum_kernel(const double * __restrict__ sums, i64 ncol, double n, double * __restrict__ mean)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col < ncol) { mean[col] = sums[col]/n; }
}

// Fixed-block estimator, first half (One/MultiDimBlockEstimator, src/Estimators.cpp:59-80, 158-185): block means
// av[b][col] = (sum of the nper samples of block b, left to right) * (1./nper); the uncorrelated estimator then runs over av.
// One thread per (column, block): every block sum has the reference's order.
__global__ void block_means_kernel(const double * __restrict__ data, i64 nper, i64 ncol, i64 nblocks, double * __restrict__ av)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    const i64 b = (i64)blockIdx.y;
    if (col >= ncol || b >= nblocks) { return; }
    const double * src = data + b*nper*ncol + col;
    double s = 0.;
    i64 i = 0;
    for (; i + 8 <= nper; i += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { v[u] = __ldcs(src + (i + u)*ncol); }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s = __dadd_rn(s, v[u]); }
    }
    for (; i < nper; ++i) { s = __dadd_rn(s, __ldcs(src + i*ncol)); }
    av[b*ncol + col] = __dmul_rn(s, 1./(double)nper);
}

// ---------------------------------------------------------------------------------------------- K5 FCBlocker
// One thread per chain, ONE pass over the series feeding all 45 block partitions (6..50 blocks) at once; block means are
// pushed into the uncorrelated estimator's running sums in block order, so every sum has the reference's order.
template <class E>
__device__ __forceinline__ double fc_err_delta(int mode, E e /* e(k): element at offset k from the centre */)
{ // calcErrDelta, src/Estimators.cpp:9-31
    switch (mode) {
    case 1: return (-0.5*e(-1) + 0.5*e(1));
    case 2: return ((1./12.)*e(-2) - (2./3.)*e(-1) + (2./3.)*e(1) - (1./12.)*e(2));
    case 3: return (-(1./60.)*e(-3) + (3./20.)*e(-2) - 0.75*e(-1) + 0.75*e(1) - (3./20.)*e(2) + (1./60.)*e(3));
    default:
        return ((1./280.)*e(-4) - (4./105.)*e(-3) + 0.2*e(-2) - 0.8*e(-1) + 0.8*e(1) - 0.2*e(2) + (4./105.)*e(3) - (1./280.)*e(4));
    }
}

#define MCIG_FC_EXACT_MAX 4096
__device__ __forceinline__ double fc_tree16(const double * v)
{ // pairwise sum: dependency depth 4 instead of 16 (the running sum is a latency chain: one thread per chain, few warps per SM)
    const double a0 = __dadd_rn(v[0], v[1]), a1 = __dadd_rn(v[2], v[3]), a2 = __dadd_rn(v[4], v[5]), a3 = __dadd_rn(v[6], v[7]);
    const double a4 = __dadd_rn(v[8], v[9]), a5 = __dadd_rn(v[10], v[11]), a6 = __dadd_rn(v[12], v[13]), a7 = __dadd_rn(v[14], v[15]);
    return __dadd_rn(__dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3)), __dadd_rn(__dadd_rn(a4, a5), __dadd_rn(a6, a7)));
}

// Stream rows [r0, r1) of one chain into the running sum, 16 loads in flight whether or not a block border falls inside the batch
// (short series have a border every ~n/1260 samples: a loop that stops at every border would issue one load at a time).
// on_border(run) is called after the sample that completes the prefix [0, next) and must advance `next`.
template <class OnBorder>
__device__ __forceinline__ void fc_stream(const double * __restrict__ src, i64 ncol, i64 r0, i64 r1, i64 & next, double & run, OnBorder && on_border)
{
    i64 i = r0;
    if (i + 16 <= r1) {
        double v[16], vn[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) { v[u] = __ldcs(src + (i + u)*ncol); }
        for (; i + 16 <= r1; i += 16) {
            // the next batch is requested before this one is consumed: 32 loads in flight per thread while the adds below run
            const bool more = (i + 32 <= r1);
            if (more) {
#pragma unroll
                for (int u = 0; u < 16; ++u) { vn[u] = __ldcs(src + (i + 16 + u)*ncol); }
            }
            if (next > i + 16) { run = __dadd_rn(run, fc_tree16(v)); }
            else {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    run = __dadd_rn(run, v[u]);
                    if (i + u + 1 == next) { on_border(run); }
                }
            }
            if (more) {
#pragma unroll
                for (int u = 0; u < 16; ++u) { v[u] = vn[u]; }
            }
        }
    }
    for (; i < r1; ++i) {
        run = __dadd_rn(run, __ldcs(src + i*ncol));
        if (i + 1 == next) { on_border(run); }
    }
}

// Statistics of the 45 partitions -> plateau search -> 5-point averages (src/Estimators.cpp:82-122, 191-246). s1(a) / s2(a) give read-write access
// to the partition sums and are overwritten with the partition means / errors (in place: the shared-memory variant keeps no per-thread arrays, and
// the plateau score is tracked as a running minimum instead of an array)
template <class A1, class A2>
__device__ __forceinline__ void fc_finish_inplace(A1 s1, A2 s2, int nobs_is_one, double & out_avg, double & out_err)
{
    constexpr int MINB = 6, MAXB = 50, NAV = MAXB - MINB + 1, MPA = 4, NACCD = NAV - 2*MPA;
    for (int a = 0; a < NAV; ++a) {
        const double nb = (double)(a + MINB);
        const double norm = 1./nb;
        const double mean = __dmul_rn(s1(a), norm);
        double er = __dadd_rn(__dmul_rn(s2(a), norm), -__dmul_rn(mean, mean));
        if (er > 1.e-300) { er = nobs_is_one ? sqrt(er/(nb - 1.)) : sqrt(__dmul_rn(er, 1./(nb - 1.))); }
        else { er = 0.; }
        s1(a) = mean;
        s2(a) = er;
    }
    int imin = 0;
    double best = 0.;
    for (int i2 = MPA; i2 < NACCD + MPA; ++i2) {
        double acc = 0.;
        for (int i1 = 1; i1 <= MPA; ++i1) { acc = __dadd_rn(acc, fc_err_delta(i1, [&](int k) { return (double)s2(i2 + k); })); }
        if (i2 == MPA || fabs(acc) < best) { // first minimum wins, as in the reference's strict comparison
            best = fabs(acc);
            imin = i2;
        }
    }
    out_avg = 0.2*(s1(imin - 2) + s1(imin - 1) + s1(imin) + s1(imin + 1) + s1(imin + 2));
    out_err = 0.2*(s2(imin - 2) + s2(imin - 1) + s2(imin) + s2(imin + 1) + s2(imin + 2));
}

__device__ __forceinline__ void fc_finish(double * s1, double * s2, int nobs_is_one, double & out_avg, double & out_err)
{
    fc_finish_inplace([&](int a) -> double & { return s1[a]; }, [&](int a) -> double & { return s2[a]; }, nobs_is_one, out_avg, out_err);
}

// Exact short-series path. The reference makes one pass per partition (45 passes, src/Estimators.cpp:207-213); a single pass feeding
// all 45 partitions needs dynamically indexed per-partition state in local memory (measured 5x slower).
// Partitions nb_lo..nb_hi share the block length nper = n/nb (for n = 100 the 45 partitions have 13 distinct block lengths): their block means are the
// same sequence, partition nb uses its first nb terms. One pass per distinct block length therefore serves the whole group: the running sums
// (t1, t2) are handed to out(partition index, t1, t2) after nb blocks for every nb of the group. Every block sum and both running sums are accumulated in
// the reference's order (src/Estimators.cpp:59-76: blocks left to right, samples left to right), so the results are the reference's bit for bit;
// four blocks are summed side by side because a single chain of dependent adds would leave the FP64 pipe idle.
template <class LD, class OUT>
__device__ __forceinline__ void fc_exact_group(i64 n, LD ld, int nb_lo, int nb_hi, OUT out)
{
    constexpr int MINB = 6;
    const i64 nper = n/nb_lo;
    const double rnper = 1./(double)nper;
    double t1 = 0., t2 = 0.;
    int j = 0;
    for (; j + 4 <= nb_hi; j += 4) {
        double b0 = 0., b1 = 0., b2 = 0., b3 = 0.;
        const i64 o0 = (i64)j*nper, o1 = o0 + nper, o2 = o1 + nper, o3 = o2 + nper;
#pragma unroll 2
        for (i64 i = 0; i < nper; ++i) {
            b0 = __dadd_rn(b0, ld(o0 + i));
            b1 = __dadd_rn(b1, ld(o1 + i));
            b2 = __dadd_rn(b2, ld(o2 + i));
            b3 = __dadd_rn(b3, ld(o3 + i));
        }
        const double av[4] = {__dmul_rn(b0, rnper), __dmul_rn(b1, rnper), __dmul_rn(b2, rnper), __dmul_rn(b3, rnper)};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            t1 = __dadd_rn(t1, av[u]);
            t2 = __dadd_rn(t2, __dmul_rn(av[u], av[u]));
            if (j + u + 1 >= nb_lo) { out(j + u + 1 - MINB, t1, t2); }
        }
    }
    for (; j < nb_hi; ++j) {
        double bsum = 0.;
        const i64 o = (i64)j*nper;
        for (i64 i = 0; i < nper; ++i) { bsum = __dadd_rn(bsum, ld(o + i)); }
        const double av = __dmul_rn(bsum, rnper);
        t1 = __dadd_rn(t1, av);
        t2 = __dadd_rn(t2, __dmul_rn(av, av));
        if (j + 1 >= nb_lo) { out(j + 1 - MINB, t1, t2); }
    }
}

// last partition (block count) that shares the block length of partition nb: n/nb' == n/nb  <=>  nb' <= n/(n/nb)
__device__ __forceinline__ int fc_group_end(i64 n, int nb)
{
    const i64 hi = n/(n/nb);
    return (hi < 50) ? (int)hi : 50;
}

template <class LD>
__device__ __forceinline__ void fc_exact_passes(i64 n, LD ld, double * s1, double * s2)
{
    for (int nb = 6; nb <= 50;) {
        const int nb_hi = fc_group_end(n, nb);
        fc_exact_group(n, ld, nb, nb_hi, [&](int a, double t1, double t2) { s1[a] = t1; s2[a] = t2; });
        nb = nb_hi + 1;
    }
}

// series read from global memory 45 times (L2-resident for the sizes this serves: n <= 4096)
__global__ void fcblocker_kernel(const double * __restrict__ data, i64 n, i64 ncol, int nobs_is_one, double * __restrict__ wavg,
                                 double * __restrict__ werr)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    double s1[45], s2[45];
    const double * src = data + col;
    fc_exact_passes(n, [&](i64 i) { return __ldg(src + i*ncol); }, s1, s2);
    fc_finish(s1, s2, nobs_is_one, wavg[col], werr[col]);
}

// short series (the 100-sample chunks of the decorrelation loop): one warp stages its 32 chains in shared memory [n][32] once
// stop: flag of a device-resident control loop (CalibCtl::done), or nullptr -- iterations enqueued behind the loop's end leave the previous results in place
__global__ void __launch_bounds__(32) fcblocker_smem_kernel(const double * __restrict__ data, i64 n, i64 ncol, int nobs_is_one, double * __restrict__ wavg,
                                                            double * __restrict__ werr, const int * __restrict__ stop)
{
    extern __shared__ double fc_tile[];
    if (stop != nullptr && *stop != 0) { return; }
    const i64 col = (i64)blockIdx.x*32 + threadIdx.x;
    const bool live = col < ncol;
    for (i64 i = 0; i < n; ++i) { fc_tile[i*32 + threadIdx.x] = live ? __ldcs(data + i*ncol + col) : 0.; }
    if (!live) { return; } // no block-level synchronisation: every thread reads only what it wrote
    double s1[45], s2[45];
    const double * src = fc_tile + threadIdx.x;
    fc_exact_passes(n, [&](i64 i) { return src[i*32]; }, s1, s2);
    fc_finish(s1, s2, nobs_is_one, wavg[col], werr[col]);
}

// The same with the partition groups of a chain spread over NW warps (the 100-sample chunks of the decorrelation loop: 0.8 ms per iteration
// at 262144 chains when one thread per chain made the reference's 45 passes): every warp of the block handles some groups of the same 32
// chains from the shared tile and warp 0 runs the plateau search. Every block sum is still accumulated in the reference's order.
template <int NW>
__global__ void __launch_bounds__(32*NW, 3) fcblocker_smem_par_kernel(const double * __restrict__ data, i64 n, i64 ncol, int nobs_is_one,
                                                                  double * __restrict__ wavg, double * __restrict__ werr, const int * __restrict__ stop)
{
    extern __shared__ double fc_tile[];
    if (stop != nullptr && *stop != 0) { return; } // (block-uniform: before the first barrier)
    double * const st = fc_tile + n*32; // [90][32] partition statistics
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const i64 col = (i64)blockIdx.x*32 + lane;
    const bool live = col < ncol;
    for (i64 i = wq; i < n; i += NW) { fc_tile[i*32 + lane] = live ? __ldcs(data + i*ncol + col) : 0.; }
    __syncthreads();
    const double * src = fc_tile + lane;
    int g = 0;
    for (int nb = 6; nb <= 50; ++g) { // groups of partitions with a common block length, dealt round-robin to the warps (a group costs ~n adds whatever its size)
        const int nb_hi = fc_group_end(n, nb);
        if (g%NW == wq) {
            fc_exact_group(n, [&](i64 i) { return src[i*32]; }, nb, nb_hi, [&](int a, double t1, double t2) {
                st[(2*a)*32 + lane] = t1;
                st[(2*a + 1)*32 + lane] = t2;
            });
        }
        nb = nb_hi + 1;
    }
    __syncthreads();
    if (wq == 0 && live) { // statistics stay in shared memory (per-thread arrays would be local memory, and the large carve-out leaves little L1)
        fc_finish_inplace([&](int a) -> double & { return st[(2*a)*32 + lane]; }, [&](int a) -> double & { return st[(2*a + 1)*32 + lane]; },
                          nobs_is_one, wavg[col], werr[col]);
    }
}

// Event-driven FCBlocker for long series: ONE streaming pass with a single running sum per chain. The (position, partition)
// pairs at which some partition completes a block are precomputed on the host and sorted by position (<= 1260 events);
// a block sum is the difference of the running sum at its two borders. Per sample the thread does one add and one compare,
// so the kernel streams at HBM speed; the running sum at every event is written out (1260 doubles per chain) and the 45
// partitions' statistics are formed from them by a second, partition-parallel pass.
// Block sums obtained as prefix differences differ from the reference's direct sums by O(1e-16 * |prefix|/|block sum|)
// relative — far inside the 1e-9 estimator tolerance; series up to MCIG_FC_EXACT_MAX samples use the exact kernel above.
// Chains are split into nseg time segments until the grid fills the machine. Pass 1 streams segment (seg, col) and records the
// LOCAL running sum at every block border inside it plus the segment total; pass 2 uses global prefix = (sum of earlier segment
// totals) + local prefix.
__global__ void fc_split_prefix_kernel(const double * __restrict__ data, i64 n, i64 ncol, int nseg, const i64 * __restrict__ ev_pos, const int * __restrict__ ev_slot,
                                       int nev, double * __restrict__ ev_prefix /*[nev][ncol], row = ev_slot[event]: partition-major*/,
                                       double * __restrict__ seg_total /*[nseg][ncol]*/)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    const int seg = (int)blockIdx.y;
    if (col >= ncol) { return; }
    const i64 len = (n + nseg - 1)/nseg;
    const i64 r0 = (i64)seg*len;
    const i64 r1 = (r0 + len < n) ? r0 + len : n;
    // first event with position > r0 (events at r0 belong to the previous segment: their prefix excludes row r0)
    int lo = 0, hi = nev;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ev_pos[mid] > r0) { hi = mid; }
        else { lo = mid + 1; }
    }
    int e = lo;
    i64 next = (e < nev) ? ev_pos[e] : n + 1;
    double run = 0.;
    fc_stream(data + col, ncol, r0, r1, next, run, [&](double r) {
        const i64 pos = next;
        while (next == pos) {
            __stcs(ev_prefix + (i64)ev_slot[e]*ncol + col, r);
            ++e;
            next = (e < nev) ? ev_pos[e] : n + 1;
        }
    });
    seg_total[(i64)seg*ncol + col] = run;
}

// exclusive scan of the segment totals per chain, in place: seg_total[s][col] becomes the sum of the segments before s
__global__ void fc_seg_scan_kernel(i64 ncol, int nseg, double * __restrict__ seg_total)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    double base = 0.;
    for (int sg = 0; sg < nseg; ++sg) {
        const double t = seg_total[(i64)sg*ncol + col];
        seg_total[(i64)sg*ncol + col] = base;
        base = __dadd_rn(base, t);
    }
}

// Pass 2, one thread per (chain, partition): the prefixes at the block borders of partition a are the rows off(a) .. off(a) + nb - 1
// of ev_prefix (the streaming pass writes them partition-major); a block sum is the difference of the global prefix = segment base
// + local prefix at its two borders. Scalar state only (a single thread walking all 1260 events of a chain needs 45 dynamically
// indexed accumulator triples, i.e. local memory: measured 1.0 GB of extra DRAM writes on an 8.6 GB series).
__global__ void fc_part_stats_kernel(i64 n, i64 ncol, const int * __restrict__ seg_of_slot, const double * __restrict__ ev_prefix,
                                     const double * __restrict__ seg_base, double * __restrict__ stats /*[45][2][ncol]*/)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    constexpr int MINB = 6;
    const int a = (int)blockIdx.y;
    const int nb = a + MINB;
    const int off = MINB*a + a*(a - 1)/2; // sum of the block counts of the partitions before a
    const double rnper = 1./(double)(n/nb);
    double last = 0., s1 = 0., s2 = 0.;
    int j = 0;
    for (; j + 4 <= nb; j += 4) { // loads of four borders in flight; the adds keep the block order
        double run[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            run[u] = __dadd_rn(seg_base[(i64)seg_of_slot[off + j + u]*ncol + col], __ldcs(ev_prefix + (i64)(off + j + u)*ncol + col));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double av = __dmul_rn(__dadd_rn(run[u], -last), rnper);
            last = run[u];
            s1 = __dadd_rn(s1, av);
            s2 = __dadd_rn(s2, __dmul_rn(av, av));
        }
    }
    for (; j < nb; ++j) {
        const double run = __dadd_rn(seg_base[(i64)seg_of_slot[off + j]*ncol + col], __ldcs(ev_prefix + (i64)(off + j)*ncol + col));
        const double av = __dmul_rn(__dadd_rn(run, -last), rnper);
        last = run;
        s1 = __dadd_rn(s1, av);
        s2 = __dadd_rn(s2, __dmul_rn(av, av));
    }
    stats[(i64)(2*a)*ncol + col] = s1;
    stats[(i64)(2*a + 1)*ncol + col] = s2;
}

__global__ void fc_final_kernel(i64 ncol, int nobs_is_one, const double * __restrict__ stats, double * __restrict__ wavg, double * __restrict__ werr)
{
    const i64 col = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (col >= ncol) { return; }
    constexpr int NAV = 45;
    double s1[NAV], s2[NAV];
    for (int a = 0; a < NAV; ++a) {
        s1[a] = stats[(i64)(2*a)*ncol + col];
        s2[a] = stats[(i64)(2*a + 1)*ncol + col];
    }
    fc_finish(s1, s2, nobs_is_one, wavg[col], werr[col]);
}

// ---------------------------------------------------------------------------------------------- device-resident calibration
// Host mirror of mcig::CalibCtl (device/mcig_device.cuh) — keep both in sync.
#define MCIG_CALIB_MAXTYPES 8
struct CalibCtl {
    double steps[MCIG_CALIB_MAXTYPES];
    u64 group;
    u64 acc;
    int done, cons_count, counter, executed;
};

struct CalibArgs { // constants of one findMRT2Step call
    double target;
    double half_dim_size[MCIG_CALIB_MAXTYPES]; // min over the coordinates of a type of 0.5*domain size (upper clamp)
    i64 steps_per_iter;  // MIN_STAT
    i64 walkers;         // local walkers (acceptance counts are local: no cross-process sum on this path)
    int ntypes;
    int N;               // _NfindMRT2Iterations
    int groups_per_step;
};

// acceptance counters -> ctl->acc (skipped once calibration is done)
__global__ void calib_reduce_kernel(const u64 * __restrict__ nacc, i64 n, CalibCtl * __restrict__ ctl)
{
    if (ctl->done != 0) { return; }
    __shared__ u64 sm[32];
    u64 s = 0;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) { s += nacc[i]; }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); }
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x < 32) {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); }
        if (threadIdx.x == 0) { ctl->acc = s; }
    }
}

// The feedback rule of MCI::findMRT2Step (src/MCIntegrator.cpp:124-167), one iteration, one thread.
__global__ void calib_controller_kernel(CalibCtl * __restrict__ ctl, const CalibArgs a)
{
    if (ctl->done != 0) { return; }
    const double acc = (double)ctl->acc, rej = (double)(a.walkers*a.steps_per_iter) - acc;
    const double rate = (ctl->acc > 0) ? acc/(acc + rej) : 0.; // getAcceptanceRate :587-592
    int cons = ctl->cons_count;
    if (fabs(rate - a.target) < 0.05) { ++cons; } else { cons = 0; } // TOLERANCE
    const double fact = fmin(2., fmax(0.5, rate/a.target));
    for (int t = 0; t < a.ntypes; ++t) {
        double st = ctl->steps[t]*fact;
        if (st > a.half_dim_size[t]) { st = a.half_dim_size[t]; }
        if (st < 1.17549435082228750797e-38) { st = 1.17549435082228750797e-38; } // FLT_MIN
        ctl->steps[t] = st;
    }
    const int counter = ctl->counter + 1;
    ctl->cons_count = cons;
    ctl->counter = counter;
    ctl->executed += 1;
    ctl->group += (u64)a.steps_per_iter*(u64)a.groups_per_step;
    const bool again = ((a.N < 0 && cons < 5) || counter < a.N) && !(a.N < 0 && counter >= -a.N);
    ctl->done = again ? 0 : 1;
}

// The stopping rule of MCI::initialDecorrelation (src/MCIntegrator.cpp:193-241), one iteration, one block. comb = the combined estimates of the
// chunk just sampled ([sum_w avg | sum_w err^2], already summed over the ranks of a sharded job), R = total number of walkers ("ranks").
// st = [oldestimate[nod] | olderr[nod] | steps counted so far | warning flag]. Iteration 0 only records the first estimate (:199-202); iteration
// it >= 1 adds MIN_NMC to the step count, stops with the warning when the maximum is reached (before estimating, as the reference does), else
// compares with the previous chunk on 2 sigma and stops when every component agrees.
struct DecorrArgs {
    double R;
    i64 min_nmc, max_steps;
    int nod, groups_per_step;
};
__global__ void decorr_controller_kernel(CalibCtl * __restrict__ ctl, const double * __restrict__ comb, double * __restrict__ st, const DecorrArgs a)
{
    if (ctl->done != 0) { return; }
    const int it = ctl->counter;
    double * old = st, * olderr = st + a.nod;
    const bool at_max = (it > 0) && ((i64)st[2*a.nod] + a.min_nmc >= a.max_steps);
    int differs = 0;
    if (!at_max) {
        for (int i = threadIdx.x; i < a.nod; i += blockDim.x) {
            const double nw = comb[i]/a.R, ne = sqrt(comb[a.nod + i])/a.R; // MPIMCI combination: src/MPIMCI.cpp:89-92
            if (it > 0 && fabs(old[i] - nw) > 2*sqrt(__dadd_rn(__dmul_rn(olderr[i], olderr[i]), __dmul_rn(ne, ne)))) { differs = 1; }
            old[i] = nw;
            olderr[i] = ne;
        }
    }
    const int any = __syncthreads_or(differs);
    if (threadIdx.x == 0) {
        ctl->group += (u64)a.min_nmc*(u64)a.groups_per_step;
        ctl->executed += 1;
        ctl->counter = it + 1;
        if (it > 0) {
            st[2*a.nod] += (double)a.min_nmc;
            if (at_max) {
                st[2*a.nod + 1] = 1.;
                ctl->done = 1;
            }
            else if (!any) { ctl->done = 1; }
        }
    }
}

// ready-queue of the dynamically scheduled walk kernel: chunk 0 of every walker block is ready, the rest is produced at run time
__global__ void dyn_init_kernel(int * __restrict__ queue, int * __restrict__ ctrl, int nblocks, int total)
{
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < total) { queue[i] = (i < nblocks) ? i : -1; }
    if (i == 0) {
        ctrl[0] = 0;       // head ticket
        ctrl[1] = nblocks; // tail ticket
        ctrl[2] = 0;       // error flag
    }
}

// ---------------------------------------------------------------------------------------------- layout helpers
// host-order [W][n] -> device-order [n][W] (replay draws, start positions)
__global__ void transpose_kernel(const double * __restrict__ in, i64 W, i64 n, double * __restrict__ out)
{
    const i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < W*n) {
        const i64 w = i%W, k = i/W;
        out[i] = in[w*n + k];
    }
}

// one walker's column of a stored series: out[r] = data[r*W + w], r < nrows (read-back of a chain: a strided 8-byte cudaMemcpy2D moves one row per
// DMA descriptor, minutes for the 2^27 rows of a BASELINE configs[3] chain block)
__global__ void gather_column_kernel(const double * __restrict__ data, i64 nrows, i64 W, i64 w, double * __restrict__ out)
{
    for (i64 r = (i64)blockIdx.x*blockDim.x + threadIdx.x; r < nrows; r += (i64)gridDim.x*blockDim.x) { out[r] = __ldcs(data + r*W + w); }
}

// ---------------------------------------------------------------------------------------------- issue-rate microbenchmarks
template <int ILP>
__global__ void dfma_peak_kernel(double * out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { v[j] = threadIdx.x*1e-9 + j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) { v[j] = fma(v[j], a, b); }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s += v[j]; }
    if (s == 123.456) { out[0] = s; }
}

template <int ILP>
__global__ void imad_peak_kernel(unsigned * out, int iters, unsigned a, unsigned b)
{
    unsigned v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { v[j] = threadIdx.x + j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) { v[j] = v[j]*a + b; }
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s += v[j]; }
    if (s == 0x12345678u) { out[0] = s; }
}

// Philox4x32-10 blocks back to back (2 independent counters per thread): the issue rate of the RNG alone. One round = 2 IMAD.WIDE.U32 + 2 LOP3;
// on sm_100a the 32x32->64 multiply issues at a quarter of the FP32 rate, so a block costs ~86 scheduler cycles per warp whatever else the loop does.
__global__ void philox_peak_kernel(unsigned * out, int iters, unsigned k0, unsigned k1)
{
    unsigned c[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; c[j][2] = blockIdx.x; c[j][3] = 7u; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            unsigned ka = k0, kb = k1;
#pragma unroll
            for (int r = 0; r < 10; ++r) {
                const unsigned long long p0 = (unsigned long long)c[j][0]*0xD2511F53ull, p1 = (unsigned long long)c[j][2]*0xCD9E8D57ull;
                const unsigned n0 = (unsigned)(p1 >> 32) ^ c[j][1] ^ ka, n2 = (unsigned)(p0 >> 32) ^ c[j][3] ^ kb;
                c[j][1] = (unsigned)p1;
                c[j][3] = (unsigned)p0;
                c[j][0] = n0;
                c[j][2] = n2;
                ka += 0x9E3779B9u;
                kb += 0xBB67AE85u;
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) { s += c[j][0] ^ c[j][1] ^ c[j][2] ^ c[j][3]; }
    if (s == 0x12345678u) { out[0] = s; }
}

} // namespace mcig_k
