// admm_eval_packed.cuh — (objective, gradient) evaluation on the PACKED count matrix (MODE 8).
//
// SeqUnwinder's training matrix is a matrix of k-mer counts (loaddata/MakeArff.java:347-368: every
// feature is an integer 0..L-k+1) that Classifier.buildClassifier standardises column by column
// (framework/Classifier.java:362-381). The fp64 modes stream the standardised doubles from HBM on
// every evaluation (8 bytes per element, 104 MB per evaluation round at cfg-2). Here the CTA keeps
// its rows of the block as the ORIGINAL counts, one byte per element, in shared memory for the
// whole launch (cfg-2: 144 rows x 656 B = 94 KB) and evaluates LOne.OptObject
// (framework/LOne.java:936-1049) on X[i][j] = (n[i][j] - mean[j]) / sd[j] without materialising X:
//
//     s[i][c]  = sum_j n[i][j] V[c][j] - o[c],   V[c][j] = W[c][j] / sd[j],  o[c] = sum_j mean[j] V[c][j]
//     G[c][j]  = (sum_i R[i][c] n[i][j] - mean[j] sum_i R[i][c]) / sd[j],    R = diag(w)(P - Y)
//
// (column 0 of the packed matrix is the constant 1 of Classifier.java:340 with mean 0, sd 1, so the
// intercept is an ordinary column and sum_i R[i][c] is the column-0 entry of the first product).
// All arithmetic is fp64; counts convert exactly. The products are the same DMMA m8n8k4 passes as
// the fp64 modes, but nothing crosses HBM after the launch prologue and nothing waits on a tile:
//
//   pass 1   partial logits per DMMA warp (K split over the 8 warps by 16-column groups: one LDS.32
//            = 4 k-steps of one row) -> shared memory
//   softmax  all 11 warps: sum of the 8 K-slices, max-shifted softmax, residuals R, NLL
//   pass 2   G += R^T n with the accumulators in registers (one LDS.U16 = 2 columns)
//   epilogue the mean / sd correction of the CTA's partial gradient, one store
//
// three __syncthreads per evaluation instead of two named-barrier hand-offs per 8-row tile.
// Relative to the fp64 modes the result differs by rounding only (X is never rounded element by
// element); against the oracle one evaluation agrees to ~1e-15 relative.
#pragma once

#include "admm_eval_tma.cuh"

namespace sqw {

constexpr int kPackedMaxGroups = 48;                 // 16-column groups per row
constexpr int kPackedMaxDim = kPackedMaxGroups * 16;  // 768 columns incl. the intercept
constexpr int kPackedSliceCap2 = 160;                // slice-cache coordinates per vector, two CTAs per SM

// no L2 hint: the packed matrix (13 MB at cfg-2) stays in L2 between the launches of an iteration
__device__ __forceinline__ void bulk_g2s_plain(void *dst, const void *src, unsigned bytes,
                                               unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Shared-memory carve-up (TileSmemLayout re-used): head = the solver's slice cache, stage = the
// CTA's rows (rows_cap8 x rsq bytes), plog = [8 warps][rows_cap8][8] partial logits (slot 0 is
// re-used for the residuals), rs = 8 x 8 offset partials + 8 class sums, bar = one mbarrier.
inline TileSmemLayout packed_smem_layout(int rsq, int rows_cap8, size_t slice_bytes, int kDmmaWarps = 8)
{
    TileSmemLayout L{};
    L.w_off = 0;
    L.w_bytes = (slice_bytes + 127) & ~(size_t)127;
    L.stage_off = L.w_bytes;
    L.stage_bytes = ((size_t)rows_cap8 * rsq + 127) & ~(size_t)127;
    L.plog_off = L.stage_off + L.stage_bytes;
    L.rs_off = L.plog_off + (size_t)kDmmaWarps * rows_cap8 * 8 * sizeof(double);
    // scratch: 8 x 8 offset partials + 8 class sums | xMean, 1/xSD [rsq each] | w [rows], y [rows]
    L.bar_off = L.rs_off + (size_t)(kDmmaWarps * 8 + 8 + 2 * rsq + rows_cap8) * sizeof(double) +
                (((size_t)rows_cap8 * sizeof(int) + 15) & ~(size_t)15);
    L.total = L.bar_off + 128;
    L.NS = rows_cap8 / kTileRows;  // tiles the resident buffer holds
    return L;
}

// Exact conversion of a byte count (I2F.F64 on the XU pipe). Measured alternatives at cfg-2, us per
// evaluation round (profiles/README.md): I2F 52.3 | 2^52 + v built with one integer op and the 2^52
// removed by a DADD 55.7 (the DADD shares the fp64 lanes the DMMA runs on) | a 256-entry
// byte -> double table in shared memory (LDS.64) 58.9 | I2F with the conversions of column group
// i+1 placed ahead of the DMMAs of group i in the source (volatile asm) 54.3 - ptxas keeps its own
// order: one conversion register, I2F -> DMMA -> I2F -> DMMA, ~40 cycles per DMMA and warp, which
// two warps per SM sub-partition overlap to ~21 cycles per DMMA (the pipe's 16 = 76 % busy).
__device__ __forceinline__ double cvt_u8(unsigned v)
{
    return __uint2double_rn(v);
}
// byte t of a 32-bit word as a double: ONE I2F.F64.U8 with a byte selector (the shift folds into the
// operand), where (w >> 8t) & 0xff through __uint2double_rn costs SHF + LOP3 + I2F.F64.U32
__device__ __forceinline__ double cvt_byte(unsigned w, int t)
{
    double d;
    asm("cvt.rn.f64.u8 %0, %1;" : "=d"(d) : "r"(w >> (8 * t)));
    return d;
}
// Shared-memory operands of the DMMA loops by 32-bit shared address: ptxas folds a constant addend
// into the instruction's immediate, so a tile needs one address per lane; the generic-pointer form
// with a predicate per load re-derived every address from %tid (ncu: half of the issue slots of
// the streamed kernel's pass 2 were address arithmetic).
__device__ __forceinline__ unsigned lds_u32(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned a)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_v2f64(unsigned a, double x, double y)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

// Pass 1 over `nt` 8-row tiles of byte rows at shared address rows_s (row stride rsq bytes, tile
// stride tstride): this warp's partial logits of every tile -> shared memory at plw (one tile =
// 8 rows x 8 classes). One shared address per lane and tile, immediate offsets for the warp's
// groups; the loads are unconditional (an absent group reads bytes of the next row or, past the
// last row, of the partial-logit buffer: multiplied by a zero V fragment) and the next tile's
// words are in flight while this tile's DMMAs issue.
template <int DW, int NBG>
__device__ __forceinline__ void packed_pass1_tiles(const double (&wreg)[NBG * 4], unsigned rows_s, unsigned plw,
                                                   unsigned tstride, int nt, int ngw)
{
    unsigned cur[NBG], nxt[NBG];
#pragma unroll
    for (int i = 0; i < NBG; ++i)
        cur[i] = nxt[i] = (nt > 0) ? lds_u32(rows_s + 16u * DW * i) : 0u;
#pragma unroll 2
    for (int j = 0; j < nt; ++j) {
        const unsigned nb_ = rows_s + (unsigned)min(j + 1, nt - 1) * tstride;
#pragma unroll
        for (int i = 0; i < NBG; ++i)
            nxt[i] = lds_u32(nb_ + 16u * DW * i);
        double pa[4][2];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            pa[c][0] = pa[c][1] = 0.0;
#pragma unroll
        for (int i = 0; i < NBG; ++i) {
            if (i >= NBG - 2 && i >= ngw)  // warp-uniform; an absent earlier group multiplies zeros
                break;
            dmma884(pa[0], cvt_byte(cur[i], 0), wreg[4 * i + 0]);
            dmma884(pa[1], cvt_byte(cur[i], 1), wreg[4 * i + 1]);
            dmma884(pa[2], cvt_byte(cur[i], 2), wreg[4 * i + 2]);
            dmma884(pa[3], cvt_byte(cur[i], 3), wreg[4 * i + 3]);
        }
        // lane holds the partial S[row gid][class 2 qd + {0, 1}]
        sts_v2f64(plw + (unsigned)(kTileRows * 8 * 8) * j, (pa[0][0] + pa[1][0]) + (pa[2][0] + pa[3][0]),
                  (pa[0][1] + pa[1][1]) + (pa[2][1] + pa[3][1]));
#pragma unroll
        for (int i = 0; i < NBG; ++i)
            cur[i] = nxt[i];
    }
}

// Pass 2 over the same tiles: acc += R^T n for HG column groups of this warp (group index gbase + i
// of the warp's NBG), residual rows at rs_ (8 x 8 doubles per tile), byte rows at xs_.
template <int DW, int NBG, int HG>
__device__ __forceinline__ void packed_pass2_tiles(double (&acc)[HG][2][2], unsigned rs_, unsigned xs_, unsigned half,
                                                   unsigned tstride, int nt, int gbase, int ngw)
{
    double a0 = lds_f64(rs_), a1 = lds_f64(rs_ + 4 * 8 * 8), na0 = a0, na1 = a1;
    unsigned x0[HG], x1[HG], nx0[HG], nx1[HG];
#pragma unroll
    for (int i = 0; i < HG; ++i) {
        x0[i] = nx0[i] = lds_u16(xs_ + 16u * DW * i);
        x1[i] = nx1[i] = lds_u16(xs_ + half + 16u * DW * i);
    }
#pragma unroll 2
    for (int j = 0; j < nt; ++j) {
        // operands of the next tile (clamped on the last) ahead of this tile's DMMAs
        const int jn = min(j + 1, nt - 1);
        const unsigned rn = rs_ + (unsigned)(kTileRows * 8 * 8) * jn, xn = xs_ + tstride * jn;
        na0 = lds_f64(rn);                  // R[row qd][class gid]
        na1 = lds_f64(rn + 4 * 8 * 8);      // R[row 4+qd][class gid]
#pragma unroll
        for (int i = 0; i < HG; ++i) {
            nx0[i] = lds_u16(xn + 16u * DW * i);
            nx1[i] = lds_u16(xn + half + 16u * DW * i);
        }
        // rows 0..3 of every group first, then rows 4..7: an accumulator is re-used at a
        // distance of 2 HG DMMAs (latency 26 cycles, issue 16)
#pragma unroll
        for (int i = 0; i < HG; ++i) {
            if (gbase + i >= NBG - 2 && gbase + i >= ngw)  // warp-uniform
                break;
            dmma884(acc[i][0], a0, cvt_byte(x0[i], 0));
            dmma884(acc[i][1], a0, cvt_byte(x0[i], 1));
        }
#pragma unroll
        for (int i = 0; i < HG; ++i) {
            if (gbase + i >= NBG - 2 && gbase + i >= ngw)
                break;
            dmma884(acc[i][0], a1, cvt_byte(x1[i], 0));
            dmma884(acc[i][1], a1, cvt_byte(x1[i], 1));
        }
        a0 = na0;
        a1 = na1;
#pragma unroll
        for (int i = 0; i < HG; ++i) {
            x0[i] = nx0[i];
            x1[i] = nx1[i];
        }
    }
}

// DW = DMMA warps per CTA (8: one CTA per SM, 352 threads; 4: two CTAs per SM, 192 threads), NT = threads.
// STREAM (MODE 12): the CTA owns more rows than its shared memory holds (a CTA budget below the
// whole GPU: concurrent sweeps and cross-validation folds; blocks of many short rows). The rows
// then pass through the same buffer PANEL by panel - cap8 rows per bulk load from L2 / HBM, the
// three phases on the panel, the gradient accumulators staying in registers across the panels - so
// the inner loops are those of the resident form and every byte is fetched once per evaluation
// (the column-chunked MODE 10 fetches twice and pays per 32-row stage). The first panel of the
// next evaluation is fetched behind the solver's exchange chain; one panel = the resident form.
template <int NCT, int DW = 8, int NT = kThreads, bool STREAM = false>
struct PackedEval {
    static_assert(NCT == 1, "packed evaluation: one class tile");
    static constexpr int NBG = kPackedMaxGroups / DW;   // 16-column groups per DMMA warp
    static constexpr int kDmmaWarps = DW, kWarps = NT / 32, kThreads = NT;
    static constexpr int HG = 6, NH = NBG / HG;         // pass 2: column groups per sweep, sweeps
    static_assert(NBG % HG == 0, "column groups per warp");
    static_assert(!STREAM || NH == 1, "panels keep one sweep's accumulators in registers");
    const AdmmDev &P;
    int blk, cta, G;
    unsigned char *rows;
    double *plog, *scr, *smean, *sinvsd, *sw;
    int *sy;
    unsigned long long *full;
    double *red;
    int cap8, r0, r1, ntiles, rsq;
    bool loaded;
    long long *prof;
    // STREAM: panels of cap8 rows; `have` = the panel in (or on its way into) the buffer, `inflight` =
    // its load has not been waited for yet, `par` = mbarrier phase parity of that load
    int npan, have;
    bool inflight;
    unsigned par;

    // thread 0: bulk-load panel p of the CTA's rows (contiguous in the block-major matrix)
    __device__ __forceinline__ void load_panel(int p)
    {
        const int pr0 = r0 + p * cap8;
        const unsigned total = (unsigned)min(cap8, r1 - pr0) * (unsigned)rsq;
        const unsigned char *src = P.Xq + ((size_t)blk * P.nb + pr0) * rsq;
        // the buffer was read through the generic proxy (LDS) until the last __syncthreads
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[0], total);
        constexpr unsigned kChunk = 16384;
        for (unsigned o = 0; o < total; o += kChunk)
            bulk_g2s_plain(rows + o, src + o, min(kChunk, total - o), &full[0]);
    }

    __device__ void init(unsigned char *smem, const TileSmemLayout &L, double *red_)
    {
        rows = smem + L.stage_off;
        plog = reinterpret_cast<double *>(smem + L.plog_off);
        scr = reinterpret_cast<double *>(smem + L.rs_off);
        full = reinterpret_cast<unsigned long long *>(smem + L.bar_off);
        red = red_;
        cap8 = L.NS * kTileRows;
        rsq = P.rsq;
        smean = scr + kDmmaWarps * 8 + 8;
        sinvsd = smean + rsq;
        sw = sinvsd + rsq;
        sy = reinterpret_cast<int *>(sw + cap8);
        cta_row_range(P, cta, G, r0, r1);
        ntiles = max(0, (r1 - r0 + kTileRows - 1) / kTileRows);
        loaded = false;
        prof = nullptr;
        npan = STREAM ? (max(0, r1 - r0) + cap8 - 1) / cap8 : 1;
        have = 0;
        inflight = STREAM && npan > 0;
        par = 0u;
        // constant for the launch: the column statistics and this CTA's instance weights / classes
        for (int i = threadIdx.x; i < rsq; i += kThreads) {
            smean[i] = __ldg(P.cmean + i);
            sinvsd[i] = __ldg(P.cinvsd + i);
        }
        if (!STREAM)
            for (int i = threadIdx.x; i < ntiles * kTileRows; i += kThreads) {
                const bool rv = r0 + i < r1;
                sw[i] = rv ? __ldg(P.w + (size_t)blk * P.nb + r0 + i) : 0.0;
                sy[i] = rv ? __ldg(P.y + (size_t)blk * P.nb + r0 + i) : -1;
            }
        if (threadIdx.x == 0) {
            mbar_init(&full[0], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (STREAM) {
            if (threadIdx.x == 0 && npan > 0)
                load_panel(0);
        } else if (threadIdx.x == 0 && ntiles > 0) {
            // the CTA's rows are contiguous in the block-major matrix: a few large bulk copies
            const unsigned total = (unsigned)(r1 - r0) * (unsigned)rsq;
            const unsigned char *src = P.Xq + ((size_t)blk * P.nb + r0) * rsq;
            mbar_expect_tx(&full[0], total);
            constexpr unsigned kChunk = 16384;
            for (unsigned o = 0; o < total; o += kChunk)
                bulk_g2s_plain(rows + o, src + o, min(kChunk, total - o), &full[0]);
        }
    }

    __device__ void set_profile(long long *p) { prof = p; }

    __device__ void drain()
    {
        if (STREAM) {
            if (inflight)
                mbar_wait(&full[0], par);
            inflight = false;
        } else if (!loaded && ntiles > 0) {
            mbar_wait(&full[0], 0);
        }
    }

    // V[class gid][16 g + 4 qd + t] = W / sd for this warp's column groups, and the warp's part of
    // the offset o[class] = sum_j mean[j] V[class][j] -> scr
    __device__ __forceinline__ void load_v(const double *xcur, double (&wreg)[NBG * 4], int warp, int qd, int gid,
                                           int dimp, int ngroups)
    {
        double pb = 0.0;
#pragma unroll
        for (int i = 0; i < NBG; ++i) {
            const int g = warp + kDmmaWarps * i;
            const int col = 16 * g + 4 * qd;
            double2 w01 = make_double2(0.0, 0.0), w23 = w01, s01 = w01, s23 = w01, m01 = w01, m23 = w01;
            if (g < ngroups) {
                s01 = *reinterpret_cast<const double2 *>(sinvsd + col);
                s23 = *reinterpret_cast<const double2 *>(sinvsd + col + 2);
                m01 = *reinterpret_cast<const double2 *>(smean + col);
                m23 = *reinterpret_cast<const double2 *>(smean + col + 2);
                if (gid < P.C) {
                    const double *wg = xcur + (size_t)gid * dimp + col;
                    w01 = __ldcg(reinterpret_cast<const double2 *>(wg));
                    w23 = __ldcg(reinterpret_cast<const double2 *>(wg + 2));
                }
            }
            wreg[4 * i + 0] = w01.x * s01.x;
            wreg[4 * i + 1] = w01.y * s01.y;
            wreg[4 * i + 2] = w23.x * s23.x;
            wreg[4 * i + 3] = w23.y * s23.y;
            pb += m01.x * wreg[4 * i + 0];
            pb += m01.y * wreg[4 * i + 1];
            pb += m23.x * wreg[4 * i + 2];
            pb += m23.y * wreg[4 * i + 3];
        }
        pb += __shfl_xor_sync(0xffffffffu, pb, 1);
        pb += __shfl_xor_sync(0xffffffffu, pb, 2);
        if (qd == 0)
            scr[warp * 8 + gid] = pb;
    }

    // softmax of the rows of `nt` tiles (all warps; the CTA's rows row0 .. of the block): residuals
    // over K-slice 0 of the partial logits, NLL into facc
    __device__ __forceinline__ void softmax_tiles(int nt, int row0, int warp, int qd, int gid, double &facc)
    {
        double *rres = plog;
        double o0 = 0.0, o1 = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < kDmmaWarps; ++w8) {
            o0 += scr[w8 * 8 + 2 * qd];
            o1 += scr[w8 * 8 + 2 * qd + 1];
        }
#pragma unroll 2
        for (int j = warp; j < nt; j += kWarps) {
            const int lr = kTileRows * j + gid;
            const bool rv = row0 + lr < r1;
            const int cls = sy[lr];      // -1 and weight 0 beyond the CTA's last row
            const double wi = sw[lr];
            double2 v = *reinterpret_cast<const double2 *>(plog + (size_t)lr * 8 + 2 * qd);
#pragma unroll
            for (int w8 = 1; w8 < kDmmaWarps; ++w8) {
                const double2 pv =
                    *reinterpret_cast<const double2 *>(plog + ((size_t)w8 * cap8 + lr) * 8 + 2 * qd);
                v.x += pv.x;
                v.y += pv.y;
            }
            v.x -= o0;
            v.y -= o1;
            double mx = -INFINITY;
            if (2 * qd < P.C)
                mx = fmax(mx, v.x);
            if (2 * qd + 1 < P.C)
                mx = fmax(mx, v.y);
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const double e0 = (2 * qd < P.C) ? exp(v.x - mx) : 0.0;
            const double e1 = (2 * qd + 1 < P.C) ? exp(v.y - mx) : 0.0;
            double den = e0 + e1;
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            double rr0 = wi * (e0 / den), rr1 = wi * (e1 / den);
            if (2 * qd == cls)
                rr0 -= wi;
            if (2 * qd + 1 == cls)
                rr1 -= wi;
            *reinterpret_cast<double2 *>(rres + (size_t)lr * 8 + 2 * qd) =
                rv ? make_double2(rr0, rr1) : make_double2(0.0, 0.0);
            const double logden = log(den);
            if (rv && 2 * qd == cls)
                facc -= wi * ((v.x - mx) - logden);
            if (rv && 2 * qd + 1 == cls)
                facc -= wi * ((v.y - mx) - logden);
        }
    }

    // mean / sd correction of HG column groups' accumulators (groups gbase .. of this warp) and the
    // store of this CTA's partial gradient
    __device__ __forceinline__ void store_groups(const double (&acc)[HG][2][2], int gbase, int warp, int qd, int gid,
                                                 int dimp, int ngroups)
    {
        const double sc = scr[kDmmaWarps * 8 + gid];
        double *gp = P.gpart + ((size_t)blk * P.Gmax + cta) * P.npad + (size_t)gid * dimp;
#pragma unroll
        for (int i = 0; i < HG; ++i) {
            const int g = warp + kDmmaWarps * (gbase + i);
            const int col = 16 * g + 4 * qd;
            if (g < ngroups) {
                const double2 s01 = *reinterpret_cast<const double2 *>(sinvsd + col);
                const double2 s23 = *reinterpret_cast<const double2 *>(sinvsd + col + 2);
                const double2 m01 = *reinterpret_cast<const double2 *>(smean + col);
                const double2 m23 = *reinterpret_cast<const double2 *>(smean + col + 2);
                double2 *dst = reinterpret_cast<double2 *>(gp + col);
                __stcg(dst, make_double2((acc[i][0][0] - m01.x * sc) * s01.x,
                                         (acc[i][1][0] - m01.y * sc) * s01.y));
                __stcg(dst + 1, make_double2((acc[i][0][1] - m23.x * sc) * s23.x,
                                             (acc[i][1][1] - m23.y * sc) * s23.y));
            }
        }
    }

    __device__ double eval(const double *xcur)
    {
        if (STREAM)
            return eval_panels(xcur);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int qd = lane & 3, gid = lane >> 2;
        const int dimp = P.dimp;
        const int ngroups = dimp >> 4;
        long long tk0 = 0, tk1 = 0;
        const bool do_prof = prof != nullptr && threadIdx.x == 0;
        if (do_prof)
            tk0 = sqw_now_ns();
#define SQW_EPROF(slot)          \
    if (do_prof) {               \
        tk1 = sqw_now_ns();      \
        prof[slot] += tk1 - tk0; \
        tk0 = tk1;               \
    }
        double *rres = plog;  // residuals overwrite K-slice 0 of the partial logits, row by row
        const bool dmma_warp = warp < kDmmaWarps;
        // (Splitting the tiles into two halves so that the three warps that issue no DMMA compute the
        // softmax of one half while the DMMA warps work on the other was measured SLOWER - 61.3 vs
        // 52.1 us per evaluation round at cfg-2: exp / div / log are DFMA chains on the same fp64 pipe
        // the DMMAs saturate, so there is nothing to overlap with.)

        // ---------------- pass 1: partial logits of every tile ----------------
        if (dmma_warp) {
            double wreg[NBG * 4];
            load_v(xcur, wreg, warp, qd, gid, dimp, ngroups);
            SQW_EPROF(0)
            if (!loaded && ntiles > 0)
                mbar_wait(&full[0], 0);
            SQW_EPROF(1)
            const int ngw = (ngroups - warp + kDmmaWarps - 1) / kDmmaWarps;  // groups of this warp
            const unsigned rows_s = smem_u32(rows) + (unsigned)(gid * rsq + 4 * qd + 16 * warp);
            const unsigned plw = smem_u32(plog) + (unsigned)((((size_t)warp * cap8 + gid) * 8 + 2 * qd) * 8);
            packed_pass1_tiles<kDmmaWarps, NBG>(wreg, rows_s, plw, (unsigned)(kTileRows * rsq), ntiles, ngw);
        }
        loaded = true;
        __syncthreads();
        SQW_EPROF(2)

        // ---------------- softmax of every row (all 11 warps): residuals, NLL ----------------
        double facc = 0.0;
        softmax_tiles(ntiles, r0, warp, qd, gid, facc);
        __syncthreads();
        SQW_EPROF(3)

        // ---------------- pass 2: G += R^T n over every tile, then the mean / sd correction and
        // the store of the CTA's partial gradient. The accumulators of HG = 6 column groups per
        // warp live in registers; a warp with 12 groups (two CTAs per SM) makes two sweeps.
        const int ngw = dmma_warp ? (ngroups - warp + kDmmaWarps - 1) / kDmmaWarps : 0;
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
            double acc[HG][2][2];
#pragma unroll
            for (int i = 0; i < HG; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    acc[i][h][0] = acc[i][h][1] = 0.0;
            if (dmma_warp && hh * HG < ngw) {
                const unsigned rs_ = smem_u32(rres) + (unsigned)((qd * 8 + gid) * 8);
                const unsigned xs_ = smem_u32(rows) +
                                     (unsigned)(qd * rsq + 2 * gid + 16 * warp + 16 * kDmmaWarps * hh * HG);
                packed_pass2_tiles<kDmmaWarps, NBG, HG>(acc, rs_, xs_, (unsigned)(4 * rsq),
                                                        (unsigned)(kTileRows * rsq), ntiles, hh * HG, ngw);
            }
            if (hh == 0) {
                // lane holds sum_i R[i][class gid] n[i][16 g + 4 qd + 2 e + h] in acc[i][h][e]; column 0
                // (g = 0: warp 0, qd = 0, h = e = 0) is the class sum of the residuals
                if (warp == 0 && qd == 0)
                    scr[kDmmaWarps * 8 + gid] = acc[0][0][0];
                __syncthreads();
                SQW_EPROF(5)
            }
            // mean / sd correction, partial gradient of this CTA
            if (dmma_warp && gid < P.C)
                store_groups(acc, hh * HG, warp, qd, gid, dimp, ngroups);
        }
        SQW_EPROF(7)
#undef SQW_EPROF
        return facc;  // per-thread NLL partial
    }

    // STREAM: the three phases panel by panel; accumulators and V fragments stay in registers
    __device__ double eval_panels(const double *xcur)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int qd = lane & 3, gid = lane >> 2;
        const int dimp = P.dimp;
        const int ngroups = dimp >> 4;
        const bool dmma_warp = warp < kDmmaWarps;
        const int ngw = dmma_warp ? (ngroups - warp + kDmmaWarps - 1) / kDmmaWarps : 0;
        double wreg[NBG * 4];
        double acc[HG][2][2];
#pragma unroll
        for (int i = 0; i < HG; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                acc[i][h][0] = acc[i][h][1] = 0.0;
        if (dmma_warp)
            load_v(xcur, wreg, warp, qd, gid, dimp, ngroups);
        double facc = 0.0;
        const unsigned tstride = (unsigned)(kTileRows * rsq);
        for (int p = 0; p < npan; ++p) {
            const int pr0 = r0 + p * cap8;
            const int nt = (min(cap8, r1 - pr0) + kTileRows - 1) / kTileRows;
            // (the buffer, the partial logits and sw / sy are free: every warp passed the barrier that
            // ended the previous panel's - or the previous evaluation's - pass 2)
            if (have != p) {
                if (threadIdx.x == 0)
                    load_panel(p);
                have = p;
                inflight = true;
            }
            for (int i = threadIdx.x; i < nt * kTileRows; i += kThreads) {
                const bool rv = pr0 + i < r1;
                sw[i] = rv ? __ldg(P.w + (size_t)blk * P.nb + pr0 + i) : 0.0;
                sy[i] = rv ? __ldg(P.y + (size_t)blk * P.nb + pr0 + i) : -1;
            }
            if (dmma_warp) {
                if (inflight)
                    mbar_wait(&full[0], par);
                const unsigned rows_s = smem_u32(rows) + (unsigned)(gid * rsq + 4 * qd + 16 * warp);
                const unsigned plw = smem_u32(plog) + (unsigned)((((size_t)warp * cap8 + gid) * 8 + 2 * qd) * 8);
                packed_pass1_tiles<kDmmaWarps, NBG>(wreg, rows_s, plw, tstride, nt, ngw);
            }
            if (inflight)
                par ^= 1u;
            inflight = false;
            __syncthreads();
            softmax_tiles(nt, pr0, warp, qd, gid, facc);
            __syncthreads();
            if (dmma_warp && ngw > 0) {
                const unsigned rs_ = smem_u32(plog) + (unsigned)((qd * 8 + gid) * 8);
                const unsigned xs_ = smem_u32(rows) + (unsigned)(qd * rsq + 2 * gid + 16 * warp);
                packed_pass2_tiles<kDmmaWarps, NBG, HG>(acc, rs_, xs_, (unsigned)(4 * rsq), tstride, nt, 0, ngw);
            }
            if (p + 1 < npan)
                __syncthreads();
        }
        // column 0 of the accumulated product is the class sum of the residuals over ALL panels
        if (warp == 0 && qd == 0)
            scr[kDmmaWarps * 8 + gid] = acc[0][0][0];
        __syncthreads();
        if (npan > 1) {   // the next evaluation starts with panel 0: fetch it behind the solver's chain
            if (threadIdx.x == 0)
                load_panel(0);
            have = 0;
            inflight = true;
        }
        if (dmma_warp && gid < P.C)
            store_groups(acc, 0, warp, qd, gid, dimp, ngroups);
        return facc;
    }
};

}  // namespace sqw
