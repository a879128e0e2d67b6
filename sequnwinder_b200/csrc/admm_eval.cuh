// admm_eval.cuh — one (objective, gradient) evaluation of the x-update sub-problem's data term.
//
// Replaces LOne.OptObject.objectiveFunction / evaluateGradient (framework/LOne.java:1007-1029,
// 941-979): s = X_b W^T, max-shifted log-sum-exp, nll -= w_i (s_y - max - log sum exp),
// G = (diag(w)(P - Y))^T X_b. The reference makes three scalar passes over the block; here a CTA
// owns a contiguous range of rows of its ADMM block and makes two tensor-pipe passes:
//
//   pass 1  S[8 rows x Cp] += X[8 rows x 4] * W^T[4 x 8]   per warp with DMMA (mma.sync m8n8k4
//           f64), X fragments straight from global memory as 16-byte loads (each lane's two
//           doubles feed two k-steps; the k order inside a fragment is a free permutation), W
//           fragments from shared memory; softmax over the 4 lanes that share a row; the residual
//           R = w_i (p - onehot) goes to a small scratch matrix, the NLL terms to a CTA sum;
//   pass 2  G[Cp x 16 cols] += R^T[8 x 4 rows] * X[4 rows x 16 cols] per warp and column group,
//           accumulators in registers over all rows of the CTA, then one store of the CTA's
//           partial gradient (reduced across the CTAs of the block in the solver's next phase).
//
// fp64 peak on B200 is the same for DFMA and DMMA (measured 34 / 37 TFLOP/s, tools/microbench),
// but DMMA needs 1/8 of the issue slots and no cross-lane reduction of the logits.
#pragma once

#include "admm_types.cuh"

namespace sqw {

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// wall-clock nanoseconds (profiling counters only)
__device__ __forceinline__ long long sqw_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return (long long)t;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA (fixed order); result valid in every thread. `red` holds >= kWarps doubles.
template <int NW = kWarps>
__device__ __forceinline__ double block_sum(double v, double *red)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i)
        s += red[i];
    return s;
}

// Rows [r0, r1) of the block handled by CTA `cta` of its group: whole 8-row tiles.
__device__ __forceinline__ void cta_row_range(const AdmmDev &P, int cta, int G, int &r0, int &r1)
{
    const long long tiles = (P.nb + 7) / 8;
    const long long t0 = tiles * cta / G, t1 = tiles * (cta + 1) / G;
    r0 = (int)(t0 * 8);
    r1 = (int)min((long long)P.nb, t1 * 8);
}

// NCT = Cp / 8 class tiles. W_SMEM: W staged in shared memory (row stride P.wstride), otherwise
// read from the global trial point (row stride P.dimp; rows C..Cp-1 of that buffer are zero).
// Returns the CTA's partial of the weighted NLL (valid in all threads) and writes the CTA's
// partial gradient to P.gpart.
template <int NCT, bool W_SMEM>
__device__ double eval_data_term(const AdmmDev &P, int blk, int cta, int G, const double *xcur,
                                 double *smW, double *red)
{
    constexpr int NBG = (NCT == 1) ? 6 : (NCT == 2 ? 4 : 3);  // column groups per warp per batch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane & 3, gid = lane >> 2;
    const int dimp = P.dimp, nb = P.nb, Cp = P.Cp;
    int r0, r1;
    cta_row_range(P, cta, G, r0, r1);

    const double *Xb = P.X + (size_t)blk * nb * dimp;
    const int *yb = P.y + (size_t)blk * nb;
    const double *wb = P.w + (size_t)blk * nb;
    double *Rb = P.R + (size_t)blk * nb * Cp;

    int wstride = dimp;
    const double *Wsrc = xcur;
    if (W_SMEM) {
        wstride = P.wstride;
        for (int i = threadIdx.x; i < Cp * dimp; i += kThreads) {
            const int c = i / dimp, j = i - c * dimp;
            smW[c * wstride + j] = (c < P.C) ? __ldcg(xcur + (size_t)c * dimp + j) : 0.0;
        }
        __syncthreads();
        Wsrc = smW;
    }

    // ---------------- pass 1: logits, softmax, NLL, residual ----------------
    double facc = 0.0;
    const int nk = dimp >> 3;  // 8 columns (two k-steps) per iteration
    for (int tile = (r0 >> 3) + warp; tile * 8 < r1; tile += kWarps) {
        const int row = tile * 8 + gid;
        const bool rv = row < r1;
        const int rowc = rv ? row : r0;
        const double2 *xp = reinterpret_cast<const double2 *>(Xb + (size_t)rowc * dimp) + q;
        const double2 *wp[NCT];
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct)
            wp[ct] = reinterpret_cast<const double2 *>(Wsrc + (size_t)(ct * 8 + gid) * wstride) + q;
        double acc[NCT][2];
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct)
            acc[ct][0] = acc[ct][1] = 0.0;
#pragma unroll 8
        for (int kk = 0; kk < nk; ++kk) {
            const double2 xa = __ldg(xp + 4 * kk);
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                const double2 wv = W_SMEM ? wp[ct][4 * kk] : __ldcg(wp[ct] + 4 * kk);
                dmma884(acc[ct], xa.x, wv.x);
                dmma884(acc[ct], xa.y, wv.y);
            }
        }
        // lane holds S[row gid][class ct*8 + 2q + {0,1}]
        double mx = -INFINITY;
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (ct * 8 + 2 * q + j < P.C)
                    mx = fmax(mx, acc[ct][j]);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        double ex[NCT][2], den = 0.0;
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                ex[ct][j] = (ct * 8 + 2 * q + j < P.C) ? exp(acc[ct][j] - mx) : 0.0;
                den += ex[ct][j];
            }
        den += __shfl_xor_sync(0xffffffffu, den, 1);
        den += __shfl_xor_sync(0xffffffffu, den, 2);
        const int cls = yb[rowc];
        const double wi = wb[rowc];
        const double logden = log(den);
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) {
            double2 rr;
            double *rp = &rr.x;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = ct * 8 + 2 * q + j;
                double r = wi * (ex[ct][j] / den);
                if (c == cls) {
                    r -= wi;
                    if (rv)
                        facc -= wi * ((acc[ct][j] - mx) - logden);
                }
                rp[j] = r;
            }
            if (rv)
                *reinterpret_cast<double2 *>(Rb + (size_t)row * Cp + ct * 8 + 2 * q) = rr;
        }
    }
    const double fsum = block_sum(facc, red);  // also orders the R stores before pass 2

    // ---------------- pass 2: G = R^T X over the CTA's rows ----------------
    const int ngroups = (dimp + 15) >> 4;  // the last group is half wide when dimp = 8 (mod 16)
    const int nrq = (r1 - r0 + 3) >> 2;
    double *gp = P.gpart + ((size_t)blk * P.Gmax + cta) * P.npad;
    for (int gb = 0; gb < ngroups; gb += kWarps * NBG) {
        double acc[NBG][2][NCT][2];
#pragma unroll
        for (int i = 0; i < NBG; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
                    acc[i][h][ct][0] = acc[i][h][ct][1] = 0.0;
#pragma unroll 4
        for (int rq = 0; rq < nrq; ++rq) {
            const int row = r0 + 4 * rq + q;
            const bool rv = row < r1;
            const int rowc = rv ? row : r0;
            double a[NCT];
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                const double v = __ldcg(Rb + (size_t)rowc * Cp + ct * 8 + gid);
                a[ct] = rv ? v : 0.0;
            }
            const double2 *xr = reinterpret_cast<const double2 *>(Xb + (size_t)rowc * dimp) + gid;
#pragma unroll
            for (int i = 0; i < NBG; ++i) {
                const int grp = gb + warp + kWarps * i;
                if (grp < ngroups) {
                    double2 xb = make_double2(0.0, 0.0);
                    if (16 * grp + 2 * gid < dimp)
                        xb = __ldg(xr + 8 * grp);
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        dmma884(acc[i][0][ct], a[ct], xb.x);
                        dmma884(acc[i][1][ct], a[ct], xb.y);
                    }
                }
            }
        }
        // lane holds G[class ct*8+gid][16*grp + 4q + {0,1,2,3}] = {even.c0, odd.c0, even.c1, odd.c1}
#pragma unroll
        for (int i = 0; i < NBG; ++i) {
            const int grp = gb + warp + kWarps * i;
            if (grp < ngroups && 16 * grp + 4 * q < dimp) {
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    const int c = ct * 8 + gid;
                    if (c < P.C) {
                        double2 *dst =
                            reinterpret_cast<double2 *>(gp + (size_t)c * dimp + 16 * grp + 4 * q);
                        __stcg(dst, make_double2(acc[i][0][ct][0], acc[i][1][ct][0]));
                        __stcg(dst + 1, make_double2(acc[i][0][ct][1], acc[i][1][ct][1]));
                    }
                }
            }
        }
    }
    return threadIdx.x == 0 ? fsum : 0.0;  // per-thread partial whose CTA sum is fsum
}

}  // namespace sqw
