// admm_eval_packed_chunk.cuh — (objective, gradient) evaluation on the PACKED count matrix for
// shapes whose rows do not stay in shared memory (MODE 10): rows of any length, up to 24 classes,
// blocks of any size — BASELINE configs[2] / [3] (200 000 x 10 921, C = 12; 2 M x 2 729, C = 17),
// and count matrices trained with a CTA budget (concurrent sweeps).
//
// Same mathematics as admm_eval_packed.cuh (X = (n - mean) / sd is never built; LOne.java:936-1049):
//     s[i][c] = sum_j n[i][j] V[c][j] - o[c],  V = W / sd,  o[c] = sum_j mean[j] V[c][j]
//     G[c][j] = (sum_i R[i][c] n[i][j] - mean[j] sum_i R[i][c]) / sd[j]
// Same schedule as the fp64 column-chunked mode 6 (admm_eval_tma.cuh): two passes over the CTA's
// rows, KC columns at a time, logits / residuals of the rows in a global scratch matrix (L2) — but
// the rows cross HBM as one BYTE per element (8x less than the fp64 matrix, which made mode 6 HBM
// bound), so what is left is the DMMA time:
//   pass 1  per column chunk: V fragments of the chunk in registers (constant for all tiles),
//           stages of 32 rows x KC bytes fetched by the three helper warps with cp.async
//           (16 B per lane, three stages deep), DMMA warps: one LDS.32 = 4 k-steps of one row,
//           I2F.F64, DMMA for every class tile; partial logits of the 8 K-slices to shared memory,
//           the helper warps add them into the global logits one stage behind;
//   softmax one row per thread: logits - o -> residuals, NLL, class sums of the residuals;
//   pass 2  per column chunk: G[:, chunk] += R^T n with the accumulators in registers, the
//           residual rows of a stage staged next to its counts; mean / sd correction; one store.
#pragma once

#include "admm_eval_packed.cuh"

namespace sqw {

constexpr int kPcRows = 32;    // rows per stage (four 8-row DMMA tiles)
constexpr int kPcStages = 3;

template <int NCT>
struct PackedChunkCfg {
    static constexpr int NBG = (NCT == 1) ? 6 : (NCT == 2 ? 4 : 3);  // 16-column groups per DMMA warp
    static constexpr int KC = 8 * NBG * 16;                           // 768 / 512 / 384 columns per chunk
    static constexpr int KCS = KC + 16;                               // stage row stride in bytes (16 mod 128)
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid)
{
    const unsigned bytes = valid ? 16u : 0u;   // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared-memory carve-up (TileSmemLayout re-used): w = the solver's slice cache (one class tile
// only), stage = kPcStages x (32 rows x KCS bytes | 32 x Cp residual doubles), plog = [2][8 warps]
// [32 rows][Cp] partial logits, rs = 8 x Cp offset partials + Cp offsets + Cp class sums.
inline TileSmemLayout packed_chunk_smem_layout(int Cp, size_t slice_bytes)
{
    TileSmemLayout L{};
    const int nct = Cp / 8;
    const int KCS = nct == 1 ? PackedChunkCfg<1>::KCS : (nct == 2 ? PackedChunkCfg<2>::KCS : PackedChunkCfg<3>::KCS);
    L.w_off = 0;
    L.w_bytes = (slice_bytes + 127) & ~(size_t)127;
    L.stage_off = L.w_bytes;
    L.stage_bytes = (size_t)kPcRows * KCS + (size_t)kPcRows * Cp * sizeof(double);
    L.plog_off = L.stage_off + (size_t)kPcStages * L.stage_bytes;
    L.rs_off = L.plog_off + (size_t)2 * kDmmaWarps * kPcRows * Cp * sizeof(double);
    L.bar_off = L.rs_off + (size_t)(kDmmaWarps * Cp + 2 * Cp) * sizeof(double);
    L.total = L.bar_off + 128;
    L.NS = kPcStages;
    return L;
}

template <int NCT>
struct PackedChunkEval {
    static constexpr int NBG = PackedChunkCfg<NCT>::NBG, KC = PackedChunkCfg<NCT>::KC,
                         KCS = PackedChunkCfg<NCT>::KCS, CP = NCT * 8;
    const AdmmDev &P;
    int blk, cta, G;
    unsigned char *stage0;
    double *plog, *scr;
    double *red;
    size_t stage_bytes;
    int r0, r1, nst, nch, rsq;

    __device__ void init(unsigned char *smem, const TileSmemLayout &L, double *red_)
    {
        stage0 = smem + L.stage_off;
        stage_bytes = L.stage_bytes;
        plog = reinterpret_cast<double *>(smem + L.plog_off);
        scr = reinterpret_cast<double *>(smem + L.rs_off);
        red = red_;
        rsq = P.rsq;
        cta_row_range(P, cta, G, r0, r1);
        nst = max(0, (r1 - r0 + kPcRows - 1) / kPcRows);
        nch = (P.dimp + KC - 1) / KC;
        __syncthreads();
    }
    __device__ void set_profile(long long *) {}
    __device__ void drain() {}

    __device__ __forceinline__ unsigned char *stage_rows(int s) const { return stage0 + (size_t)s * stage_bytes; }
    __device__ __forceinline__ double *stage_res(int s) const
    {
        return reinterpret_cast<double *>(stage0 + (size_t)s * stage_bytes + (size_t)kPcRows * KCS);
    }

    // helper warps: fetch stage `st` of column chunk `c` (and, in pass 2, its residual rows)
    __device__ void fetch(int c, int st, bool with_res, const double *Lb)
    {
        const int ht = (int)threadIdx.x - kDmmaWarps * 32;          // 0..95
        const int nht = kThreads - kDmmaWarps * 32;
        const int s = st % kPcStages;
        const int wc = min(KC, P.dimp - c * KC);                      // bytes of this chunk (multiple of 16)
        const int pieces = wc >> 4;
        const int row0 = r0 + st * kPcRows;
        unsigned char *dst = stage_rows(s);
        const unsigned char *src = P.Xq + ((size_t)blk * P.nb + row0) * rsq + (size_t)c * KC;
        for (int i = ht; i < kPcRows * pieces; i += nht) {
            const int r = i / pieces, q = i - r * pieces;
            const bool valid = row0 + r < r1;
            cp_async16(dst + (size_t)r * KCS + 16 * q, src + (valid ? (size_t)r * rsq + 16 * q : 0), valid);
        }
        if (with_res) {
            double *rd = stage_res(s);
            const double *rsrc = Lb + (size_t)row0 * CP;
            constexpr int per_row = CP * 8 / 16;
            for (int i = ht; i < kPcRows * per_row; i += nht) {
                const int r = i / per_row;
                const bool valid = row0 + r < r1;
                cp_async16(reinterpret_cast<unsigned char *>(rd) + 16 * i,
                           reinterpret_cast<const unsigned char *>(rsrc) + (valid ? 16 * (size_t)i : 0), valid);
            }
        }
        cp_async_commit();
    }

    __device__ double eval(const double *xcur)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int qd = lane & 3, gid = lane >> 2;
        const int dimp = P.dimp;
        const bool dmma_warp = warp < kDmmaWarps;
        double *Lb = P.R + (size_t)blk * P.nb * CP;   // logits, then residuals, of this block's rows
        double *so = scr;                               // [8 warps][CP] offset partials
        double *o_all = scr + kDmmaWarps * CP;          // [CP] offsets o[c]
        double *sc_all = o_all + CP;                    // [CP] class sums of the residuals
        if (dmma_warp && qd == 0)
            for (int ct = 0; ct < NCT; ++ct)
                so[warp * CP + ct * 8 + gid] = 0.0;

        // ---------------- pass 1: logits, chunk by chunk ----------------
        for (int c = 0; c < nch; ++c) {
            const int wc = min(KC, dimp - c * KC);
            const int ngroups = wc >> 4;
            double vreg[NBG * 4][NCT];
            if (dmma_warp) {
                // V fragments of this warp's column groups (constant for all tiles of the chunk)
                double pb[NCT];
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
                    pb[ct] = 0.0;
#pragma unroll
                for (int i = 0; i < NBG; ++i) {
                    const int g = warp + kDmmaWarps * i;
                    const int col = c * KC + 16 * g + 4 * qd;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const bool ok = g < ngroups;
                        const double is = ok ? __ldg(P.cinvsd + col + t) : 0.0;
                        const double mu = ok ? __ldg(P.cmean + col + t) : 0.0;
#pragma unroll
                        for (int ct = 0; ct < NCT; ++ct) {
                            const int cls = ct * 8 + gid;
                            const double w = (ok && cls < P.C) ? __ldcg(xcur + (size_t)cls * dimp + col + t) : 0.0;
                            vreg[4 * i + t][ct] = w * is;
                            pb[ct] += mu * vreg[4 * i + t][ct];
                        }
                    }
                }
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    double v = pb[ct];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (qd == 0)
                        so[warp * CP + ct * 8 + gid] += v;   // this warp's own slot, summed over the chunks
                }
            } else {
                if (nst > 0)
                    fetch(c, 0, false, Lb);
                if (nst > 1)
                    fetch(c, 1, false, Lb);
                else
                    cp_async_commit();
                cp_async_wait<1>();
            }
            __syncthreads();
            const int ngw = dmma_warp ? (ngroups - warp + kDmmaWarps - 1) / kDmmaWarps : 0;
            for (int st = 0; st <= nst; ++st) {
                if (dmma_warp) {
                    if (st < nst) {
                        const unsigned char *rows = stage_rows(st % kPcStages);
                        double *pl = plog + ((size_t)(st & 1) * kDmmaWarps + warp) * kPcRows * CP;
#pragma unroll 1
                        for (int tt = 0; tt < kPcRows / 8; ++tt) {
                            const unsigned char *rp = rows + (size_t)(8 * tt + gid) * KCS + 4 * qd;
                            unsigned wv[NBG];
#pragma unroll
                            for (int i = 0; i < NBG; ++i)
                                wv[i] = (i < ngw) ? *reinterpret_cast<const unsigned *>(rp + 16 * (warp + kDmmaWarps * i)) : 0u;
                            double pa[NCT][2][2];
#pragma unroll
                            for (int ct = 0; ct < NCT; ++ct)
                                pa[ct][0][0] = pa[ct][0][1] = pa[ct][1][0] = pa[ct][1][1] = 0.0;
#pragma unroll
                            for (int i = 0; i < NBG; ++i) {
                                if (i >= ngw)  // warp-uniform
                                    break;
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const double xa = cvt_u8((wv[i] >> (8 * t)) & 0xffu);
#pragma unroll
                                    for (int ct = 0; ct < NCT; ++ct)
                                        dmma884(pa[ct][t & 1], xa, vreg[4 * i + t][ct]);
                                }
                            }
                            // lane holds the partial S[row gid][class ct*8 + 2 qd + {0, 1}]
#pragma unroll
                            for (int ct = 0; ct < NCT; ++ct)
                                *reinterpret_cast<double2 *>(pl + (size_t)(8 * tt + gid) * CP + ct * 8 + 2 * qd) =
                                    make_double2(pa[ct][0][0] + pa[ct][1][0], pa[ct][0][1] + pa[ct][1][1]);
                        }
                    }
                } else {
                    // helper warps: next fetch, then the logits of the PREVIOUS stage += its 8 K-slices
                    if (st + 2 < nst)
                        fetch(c, st + 2, false, Lb);
                    else
                        cp_async_commit();
                    if (st >= 1) {
                        const int ht = (int)threadIdx.x - kDmmaWarps * 32;
                        const int row0 = r0 + (st - 1) * kPcRows;
                        const double *pl = plog + (size_t)((st - 1) & 1) * kDmmaWarps * kPcRows * CP;
                        for (int i = ht; i < kPcRows * CP / 2; i += kThreads - kDmmaWarps * 32) {
                            const int r = (2 * i) / CP, cc = 2 * i - r * CP;
                            if (row0 + r < r1) {
                                double2 v = *reinterpret_cast<const double2 *>(pl + (size_t)r * CP + cc);
#pragma unroll
                                for (int w8 = 1; w8 < kDmmaWarps; ++w8) {
                                    const double2 pv = *reinterpret_cast<const double2 *>(
                                        pl + ((size_t)w8 * kPcRows + r) * CP + cc);
                                    v.x += pv.x;
                                    v.y += pv.y;
                                }
                                double2 *lp = reinterpret_cast<double2 *>(Lb + (size_t)(row0 + r) * CP + cc);
                                if (c > 0) {
                                    const double2 old = *lp;
                                    v.x += old.x;
                                    v.y += old.y;
                                }
                                *lp = v;
                            }
                        }
                    }
                    cp_async_wait<1>();
                }
                __syncthreads();
            }
        }
        // offsets o[c] = sum over the 8 warps' slots (fixed order)
        if (threadIdx.x < CP) {
            double s = 0.0;
            for (int w8 = 0; w8 < kDmmaWarps; ++w8)
                s += so[w8 * CP + threadIdx.x];
            o_all[threadIdx.x] = s;
        }
        __syncthreads();

        // ---------------- softmax over my rows: logits -> residuals, NLL, class sums ----------------
        double facc = 0.0;
        double rsum[CP];
#pragma unroll
        for (int cc = 0; cc < CP; ++cc)
            rsum[cc] = 0.0;
        {
            const int *yb = P.y + (size_t)blk * P.nb;
            const double *wb = P.w + (size_t)blk * P.nb;
            for (int row = r0 + threadIdx.x; row < r1; row += kThreads) {
                double *lp = Lb + (size_t)row * CP;
                double lv[CP];
                double mx = -INFINITY;
#pragma unroll
                for (int cc = 0; cc < CP; ++cc) {
                    lv[cc] = lp[cc] - o_all[cc];
                    if (cc < P.C)
                        mx = fmax(mx, lv[cc]);
                }
                double den = 0.0;
#pragma unroll
                for (int cc = 0; cc < CP; ++cc) {
                    lv[cc] = (cc < P.C) ? exp(lv[cc] - mx) : 0.0;
                    den += lv[cc];
                }
                const int cls = yb[row];
                const double wi = wb[row];
                facc -= wi * ((lp[cls] - o_all[cls] - mx) - log(den));
#pragma unroll
                for (int cc = 0; cc < CP; ++cc) {
                    double r = 0.0;
                    if (cc < P.C) {
                        r = wi * (lv[cc] / den);
                        if (cc == cls)
                            r -= wi;
                    }
                    lp[cc] = r;
                    rsum[cc] += r;
                }
            }
        }
#pragma unroll
        for (int cc = 0; cc < CP; ++cc) {
            const double s = block_sum(rsum[cc], red);   // (also orders the residual stores before pass 2)
            if (threadIdx.x == 0)
                sc_all[cc] = s;
        }
        __syncthreads();

        // ---------------- pass 2: gradient, chunk by chunk ----------------
        double *gp = P.gpart + ((size_t)blk * P.Gmax + cta) * P.npad;
        for (int c = 0; c < nch; ++c) {
            const int wc = min(KC, dimp - c * KC);
            const int ngroups = wc >> 4;
            const int ngw = dmma_warp ? (ngroups - warp + kDmmaWarps - 1) / kDmmaWarps : 0;
            double acc[NBG][2][NCT][2];
#pragma unroll
            for (int i = 0; i < NBG; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct)
                        acc[i][h][ct][0] = acc[i][h][ct][1] = 0.0;
            if (!dmma_warp) {
                if (nst > 0)
                    fetch(c, 0, true, Lb);
                if (nst > 1)
                    fetch(c, 1, true, Lb);
                else
                    cp_async_commit();
                cp_async_wait<1>();
            }
            __syncthreads();
            for (int st = 0; st < nst; ++st) {
                if (dmma_warp) {
                    const unsigned char *rows = stage_rows(st % kPcStages);
                    const double *res = stage_res(st % kPcStages);
#pragma unroll 1
                    for (int tt = 0; tt < kPcRows / 8; ++tt) {
                        double a0[NCT], a1[NCT];
#pragma unroll
                        for (int ct = 0; ct < NCT; ++ct) {
                            a0[ct] = res[(size_t)(8 * tt + qd) * CP + ct * 8 + gid];       // R[row qd][class]
                            a1[ct] = res[(size_t)(8 * tt + 4 + qd) * CP + ct * 8 + gid];   // R[row 4+qd][class]
                        }
                        const unsigned char *rp0 = rows + (size_t)(8 * tt + qd) * KCS + 2 * gid;
                        const unsigned char *rp1 = rp0 + 4 * (size_t)KCS;
                        unsigned x0[NBG], x1[NBG];
#pragma unroll
                        for (int i = 0; i < NBG; ++i) {
                            const int off = 16 * (warp + kDmmaWarps * i);
                            x0[i] = (i < ngw) ? *reinterpret_cast<const unsigned short *>(rp0 + off) : 0u;
                            x1[i] = (i < ngw) ? *reinterpret_cast<const unsigned short *>(rp1 + off) : 0u;
                        }
#pragma unroll
                        for (int i = 0; i < NBG; ++i) {
                            if (i >= ngw)
                                break;
                            const double b00 = cvt_u8(x0[i] & 0xffu), b01 = cvt_u8(x0[i] >> 8);
                            const double b10 = cvt_u8(x1[i] & 0xffu), b11 = cvt_u8(x1[i] >> 8);
                            // (an accumulator is re-used at a distance of 2 NCT DMMAs)
#pragma unroll
                            for (int ct = 0; ct < NCT; ++ct)
                                dmma884(acc[i][0][ct], a0[ct], b00);
#pragma unroll
                            for (int ct = 0; ct < NCT; ++ct)
                                dmma884(acc[i][1][ct], a0[ct], b01);
#pragma unroll
                            for (int ct = 0; ct < NCT; ++ct)
                                dmma884(acc[i][0][ct], a1[ct], b10);
#pragma unroll
                            for (int ct = 0; ct < NCT; ++ct)
                                dmma884(acc[i][1][ct], a1[ct], b11);
                        }
                    }
                } else {
                    if (st + 2 < nst)
                        fetch(c, st + 2, true, Lb);
                    else
                        cp_async_commit();
                    cp_async_wait<1>();
                }
                __syncthreads();
            }
            // mean / sd correction and store of this chunk's columns: lane holds
            // sum_i R[i][class ct*8+gid] n[i][c KC + 16 g + 4 qd + 2 e + h] in acc[i][h][ct][e]
            if (dmma_warp) {
#pragma unroll
                for (int i = 0; i < NBG; ++i) {
                    const int g = warp + kDmmaWarps * i;
                    const int col = c * KC + 16 * g + 4 * qd;
                    if (g < ngroups) {
                        const double2 s01 = __ldg(reinterpret_cast<const double2 *>(P.cinvsd + col));
                        const double2 s23 = __ldg(reinterpret_cast<const double2 *>(P.cinvsd + col + 2));
                        const double2 m01 = __ldg(reinterpret_cast<const double2 *>(P.cmean + col));
                        const double2 m23 = __ldg(reinterpret_cast<const double2 *>(P.cmean + col + 2));
#pragma unroll
                        for (int ct = 0; ct < NCT; ++ct) {
                            const int cls = ct * 8 + gid;
                            if (cls < P.C) {
                                const double sc = sc_all[cls];
                                double2 *dst = reinterpret_cast<double2 *>(gp + (size_t)cls * dimp + col);
                                __stcg(dst, make_double2((acc[i][0][ct][0] - m01.x * sc) * s01.x,
                                                         (acc[i][1][ct][0] - m01.y * sc) * s01.y));
                                __stcg(dst + 1, make_double2((acc[i][0][ct][1] - m23.x * sc) * s23.x,
                                                             (acc[i][1][ct][1] - m23.y * sc) * s23.y));
                            }
                        }
                    }
                }
            }
        }
        const double fs = block_sum(facc, red);
        return threadIdx.x == 0 ? fs : 0.0;  // per-thread partial whose CTA sum is fs
    }
};

}  // namespace sqw
