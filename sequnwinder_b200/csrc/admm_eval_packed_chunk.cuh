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
//           the helper warps add them into the global logits one stage behind (tree sum of the
//           eight slices, RED.ADD.F64 into L2 from the second chunk on: no load to wait for);
//   softmax one row per thread: logits - o -> residuals, NLL, class sums of the residuals;
//   pass 2  per column chunk: G[:, chunk] += R^T n with the accumulators in registers, the
//           residual rows of a stage staged next to its counts; mean / sd correction; one store.
// Inner loops (ncu of the first version: DMMA pipe 50 % busy, half of pass 2's issue slots were
// address arithmetic - every predicated LDS.U16 recomputed its generic address from %tid): all
// shared-memory operands go through ONE 32-bit shared address per lane and stage plus immediate
// offsets, the loads are unconditional (a stage row always holds KC bytes; absent groups are
// skipped at the DMMAs), the four tiles of a stage are unrolled with the next tile's operands
// loaded ahead of the current tile's DMMAs, and full chunks run a branch-free instantiation.
//
// EX (MODE 11): C = 8 NCT + 1 classes. SeqUnwinder's label designs often are 2^k labels plus one
// outgroup / "Random" class (BASELINE configs[3]: 16 + 1 = 17), and an fp64 MMA tile is 8 classes
// wide, so a third class tile would issue 24 classes' worth of DMMAs for 17. The last class instead
// rides along as plain DFMAs on operands the DMMA path has already converted: in pass 1 every lane
// multiplies its count by the class's V entry of the same column (one DFMA per conversion, then a
// 4-lane shuffle sum per row), in pass 2 by the class's residual of the same row (accumulators per
// column, shuffle sum at the end of the chunk). One DFMA warp instruction costs the fp64 pipe 2
// cycles against 16 for a DMMA: 17 classes cost 2 x 16 + 2 instead of 3 x 16 cycles per conversion.
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

__device__ __forceinline__ void cp_async16(unsigned dst_s, const void *src, bool valid)
{
    const unsigned bytes = valid ? 16u : 0u;   // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_s), "l"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared-memory carve-up (TileSmemLayout re-used): w = the solver's slice cache (one class tile
// only), stage = kPcStages x (32 rows x KCS bytes | 32 residual rows of CL doubles at a stride that
// makes pass 2's fragment loads conflict-free), plog = [2][8 warps][32 rows][CL] partial logits,
// rs = 8 x CL offset partials + CL offsets + CL class sums. CL = 8 nct (+ 2 with the extra class).
inline TileSmemLayout packed_chunk_smem_layout(int nct, bool extra, size_t slice_bytes)
{
    TileSmemLayout L{};
    const int KCS = nct == 1 ? PackedChunkCfg<1>::KCS : (nct == 2 ? PackedChunkCfg<2>::KCS : PackedChunkCfg<3>::KCS);
    const int CL = 8 * nct + (extra ? 2 : 0), RS = CL + (extra ? 2 : 4);
    L.w_off = 0;
    L.w_bytes = (slice_bytes + 127) & ~(size_t)127;
    L.stage_off = L.w_bytes;
    L.stage_bytes = (size_t)kPcRows * KCS + (size_t)kPcRows * RS * sizeof(double);
    L.plog_off = L.stage_off + (size_t)kPcStages * L.stage_bytes;
    L.rs_off = L.plog_off + (size_t)2 * kDmmaWarps * kPcRows * CL * sizeof(double);
    L.bar_off = L.rs_off + (size_t)(kDmmaWarps * CL + 2 * CL) * sizeof(double);
    L.total = L.bar_off + 128;
    L.NS = kPcStages;
    return L;
}

template <int NCT, bool EX = false>
struct PackedChunkEval {
    static constexpr int NBG = PackedChunkCfg<NCT>::NBG, KC = PackedChunkCfg<NCT>::KC,
                         KCS = PackedChunkCfg<NCT>::KCS, CP = NCT * 8;
    static constexpr int CL = CP + (EX ? 2 : 0);               // doubles per logits / residual row; the
                                                               // extra class is entry CP, entry CP + 1 is 0
    static constexpr int RS = CL + (EX ? 2 : 4);               // residual row stride in a stage (4 or 12 mod 16)
    static constexpr int kHelpers = kThreads - kDmmaWarps * 32;  // 96 threads of the three helper warps
    static constexpr int kTiles = kPcRows / 8;
    static constexpr unsigned kPlogHalf = (unsigned)kDmmaWarps * kPcRows * CL * 8;  // bytes of one plog buffer
    const AdmmDev &P;
    int blk, cta, G;
    unsigned stage_s, plog_s;   // 32-bit shared addresses of stage 0 / the partial logits
    double *scr;
    double *red;
    unsigned stage_bytes;
    int r0, r1, nst, nch, rsq;

    __device__ void init(unsigned char *smem, const TileSmemLayout &L, double *red_)
    {
        stage_s = smem_u32(smem + L.stage_off);
        stage_bytes = (unsigned)L.stage_bytes;
        plog_s = smem_u32(smem + L.plog_off);
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

    __device__ __forceinline__ unsigned stage_addr(int st) const
    {
        return stage_s + (unsigned)(st % kPcStages) * stage_bytes;
    }

    // helper warps: fetch stage `st` of column chunk `c` (and, in pass 2, its residual rows)
    __device__ void fetch(int c, int st, bool with_res, const double *Lb)
    {
        const int ht = (int)threadIdx.x - kDmmaWarps * 32;          // 0..95
        const unsigned sb = stage_addr(st);
        const int wc = min(KC, P.dimp - c * KC);                      // bytes of this chunk (multiple of 16)
        const int pieces = wc >> 4;
        const int row0 = r0 + st * kPcRows;
        const unsigned char *src = P.Xq + ((size_t)blk * P.nb + row0) * rsq + (size_t)c * KC;
        // copy i = ht + 96 k is piece q of row r: one division per call, then (r, q) advance by steps
        int r = ht / pieces, q = ht - r * pieces;
        const int dr = kHelpers / pieces, dq = kHelpers - dr * pieces;
        while (r < kPcRows) {
            const bool valid = row0 + r < r1;
            cp_async16(sb + (unsigned)(r * KCS + 16 * q), src + (valid ? (size_t)r * rsq + 16 * q : 0), valid);
            q += dq;
            r += dr;
            if (q >= pieces) {
                q -= pieces;
                ++r;
            }
        }
        if (with_res) {
            const unsigned rd = sb + (unsigned)(kPcRows * KCS);
            const unsigned char *rsrc = reinterpret_cast<const unsigned char *>(Lb + (size_t)row0 * CL);
            constexpr int per_row = CL / 2;   // 16-byte pieces of a residual row
#pragma unroll
            for (int k = 0; k < (kPcRows * per_row + kHelpers - 1) / kHelpers; ++k) {
                const int i = ht + k * kHelpers;
                if (i < kPcRows * per_row) {
                    const int rr = i / per_row, qq = i - rr * per_row;
                    const bool valid = row0 + rr < r1;
                    cp_async16(rd + (unsigned)(rr * RS * 8 + 16 * qq), rsrc + (valid ? 16 * (size_t)i : 0), valid);
                }
            }
        }
        cp_async_commit();
    }

    // helper warps: global logits of stage `st` (+)= tree sum of its 8 K-slices. First chunk: store;
    // later chunks: RED.ADD.F64 into L2 (same thread, same address as the store: ordered).
    __device__ __forceinline__ void accumulate(int c, int st, double *Lb) const
    {
        const int ht = (int)threadIdx.x - kDmmaWarps * 32;
        const int row0 = r0 + st * kPcRows;
        const unsigned pl = plog_s + (unsigned)(st & 1) * kPlogHalf;
#pragma unroll
        for (int k = 0; k < (kPcRows * CL / 2 + kHelpers - 1) / kHelpers; ++k) {
            const int i = ht + k * kHelpers;
            if (i < kPcRows * CL / 2) {
                const int r = (2 * i) / CL, cc = 2 * i - r * CL;
                if (row0 + r < r1) {
                    double2 p[kDmmaWarps];
#pragma unroll
                    for (int w8 = 0; w8 < kDmmaWarps; ++w8)
                        p[w8] = lds_v2f64(pl + (unsigned)(((w8 * kPcRows + r) * CL + cc) * 8));
                    double2 v;
                    v.x = ((p[0].x + p[1].x) + (p[2].x + p[3].x)) + ((p[4].x + p[5].x) + (p[6].x + p[7].x));
                    v.y = ((p[0].y + p[1].y) + (p[2].y + p[3].y)) + ((p[4].y + p[5].y) + (p[6].y + p[7].y));
                    double *lp = Lb + (size_t)(row0 + r) * CL + cc;
                    if (c == 0) {
                        __stcg(reinterpret_cast<double2 *>(lp), v);
                    } else {
                        atomicAdd(lp, v.x);
                        atomicAdd(lp + 1, v.y);
                    }
                }
            }
        }
    }

    // DMMA warps, pass 1, one stage: partial logits of the four tiles over this warp's column groups.
    // sb = stage + lane offset (row gid, byte 4 qd + 16 warp); pw = this warp's plog slice + lane offset.
    // EX: vex = the extra class's V entries of this lane's columns; its partial logit of row gid is
    // the 4-lane sum of the lanes' DFMA chains, written by lane qd = 0 to entry CP of the row.
    template <bool FULL>
    __device__ __forceinline__ void pass1_stage(unsigned sb, unsigned pw, const double (&vreg)[NBG * 4][NCT],
                                                const double (&vex)[NBG * 4], int ngw) const
    {
        unsigned cur[NBG], nxt[NBG];
#pragma unroll
        for (int i = 0; i < NBG; ++i)
            cur[i] = nxt[i] = lds_u32(sb + 128u * i);
#pragma unroll
        for (int tt = 0; tt < kTiles; ++tt) {
            if (tt + 1 < kTiles) {
#pragma unroll
                for (int i = 0; i < NBG; ++i)
                    nxt[i] = lds_u32(sb + (unsigned)((tt + 1) * 8 * KCS) + 128u * i);
            }
            double pa[NCT][2][2];
            double ex0 = 0.0, ex1 = 0.0;
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct)
                pa[ct][0][0] = pa[ct][0][1] = pa[ct][1][0] = pa[ct][1][1] = 0.0;
#pragma unroll
            for (int i = 0; i < NBG; ++i) {
                if (!FULL && i >= ngw)  // warp-uniform
                    break;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double xa = cvt_byte(cur[i], t);
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct)
                        dmma884(pa[ct][t & 1], xa, vreg[4 * i + t][ct]);
                    if (EX) {
                        if (t & 1)
                            ex1 = fma(xa, vex[4 * i + t], ex1);
                        else
                            ex0 = fma(xa, vex[4 * i + t], ex0);
                    }
                }
            }
            // lane holds the partial S[row gid][class ct*8 + 2 qd + {0, 1}]
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct)
                sts_v2f64(pw + (unsigned)((8 * tt * CL + ct * 8) * 8), pa[ct][0][0] + pa[ct][1][0],
                          pa[ct][0][1] + pa[ct][1][1]);
            if (EX) {
                double e = ex0 + ex1;
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                e += __shfl_xor_sync(0xffffffffu, e, 2);
                if ((threadIdx.x & 3) == 0)   // pw points at classes 2 qd = 0 of row gid
                    sts_v2f64(pw + (unsigned)((8 * tt * CL + CP) * 8), e, 0.0);
            }
#pragma unroll
            for (int i = 0; i < NBG; ++i)
                cur[i] = nxt[i];
        }
    }

    // DMMA warps, pass 2, one stage: acc += R^T n over the four tiles.
    // sb = stage + lane offset (row qd, byte 2 gid + 16 warp); rb = residual rows + lane offset (row qd, class gid).
    // EX: rx = residual rows + (row qd, entry CP); aex[i][h] accumulates the extra class's
    // sum over this lane's rows of R[row][CP] n[row][16 g + 2 gid + h].
    template <bool FULL>
    __device__ __forceinline__ void pass2_stage(unsigned sb, unsigned rb, unsigned rx,
                                                double (&acc)[NBG][2][NCT][2], double (&aex)[NBG][2],
                                                int ngw) const
    {
        double a0[NCT], a1[NCT], na0[NCT], na1[NCT];
        double e0 = 0.0, e1 = 0.0, ne0 = 0.0, ne1 = 0.0;
        unsigned x0[NBG], x1[NBG], nx0[NBG], nx1[NBG];
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) {
            a0[ct] = na0[ct] = lds_f64(rb + (unsigned)(ct * 64));                  // R[row qd][class]
            a1[ct] = na1[ct] = lds_f64(rb + (unsigned)((4 * RS + ct * 8) * 8));    // R[row 4+qd][class]
        }
        if (EX) {
            e0 = ne0 = lds_f64(rx);
            e1 = ne1 = lds_f64(rx + (unsigned)(4 * RS * 8));
        }
#pragma unroll
        for (int i = 0; i < NBG; ++i) {
            x0[i] = nx0[i] = lds_u16(sb + 128u * i);
            x1[i] = nx1[i] = lds_u16(sb + (unsigned)(4 * KCS) + 128u * i);
        }
#pragma unroll
        for (int tt = 0; tt < kTiles; ++tt) {
            if (tt + 1 < kTiles) {
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    na0[ct] = lds_f64(rb + (unsigned)(((8 * tt + 8) * RS + ct * 8) * 8));
                    na1[ct] = lds_f64(rb + (unsigned)(((8 * tt + 12) * RS + ct * 8) * 8));
                }
                if (EX) {
                    ne0 = lds_f64(rx + (unsigned)((8 * tt + 8) * RS * 8));
                    ne1 = lds_f64(rx + (unsigned)((8 * tt + 12) * RS * 8));
                }
#pragma unroll
                for (int i = 0; i < NBG; ++i) {
                    nx0[i] = lds_u16(sb + (unsigned)((8 * tt + 8) * KCS) + 128u * i);
                    nx1[i] = lds_u16(sb + (unsigned)((8 * tt + 12) * KCS) + 128u * i);
                }
            }
#pragma unroll
            for (int i = 0; i < NBG; ++i) {
                if (!FULL && i >= ngw)
                    break;
                const double b00 = cvt_byte(x0[i], 0), b01 = cvt_byte(x0[i], 1);
                const double b10 = cvt_byte(x1[i], 0), b11 = cvt_byte(x1[i], 1);
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
                if (EX) {
                    aex[i][0] = fma(e0, b00, aex[i][0]);
                    aex[i][1] = fma(e0, b01, aex[i][1]);
                    aex[i][0] = fma(e1, b10, aex[i][0]);
                    aex[i][1] = fma(e1, b11, aex[i][1]);
                }
            }
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                a0[ct] = na0[ct];
                a1[ct] = na1[ct];
            }
            e0 = ne0;
            e1 = ne1;
#pragma unroll
            for (int i = 0; i < NBG; ++i) {
                x0[i] = nx0[i];
                x1[i] = nx1[i];
            }
        }
    }

    __device__ double eval(const double *xcur)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int qd = lane & 3, gid = lane >> 2;
        const int dimp = P.dimp;
        const bool dmma_warp = warp < kDmmaWarps;
        double *Lb = P.R + (size_t)blk * P.nb * CL;   // logits, then residuals, of this block's rows
        double *so = scr;                               // [8 warps][CL] offset partials
        double *o_all = scr + kDmmaWarps * CL;          // [CL] offsets o[c]
        double *sc_all = o_all + CL;                    // [CL] class sums of the residuals
        // per-lane byte offsets into a stage / a plog slice (see pass1_stage / pass2_stage)
        const unsigned off1 = (unsigned)(gid * KCS + 4 * qd + 16 * warp);
        const unsigned off2 = (unsigned)(qd * KCS + 2 * gid + 16 * warp);
        const unsigned offr = (unsigned)(kPcRows * KCS + (qd * RS + gid) * 8);
        const unsigned offx = (unsigned)(kPcRows * KCS + (qd * RS + CP) * 8);
        const unsigned offp = (unsigned)(((warp * kPcRows + gid) * CL + 2 * qd) * 8);
        if (dmma_warp && qd == 0) {
            for (int ct = 0; ct < NCT; ++ct)
                so[warp * CL + ct * 8 + gid] = 0.0;
            if (EX && gid < 2)
                so[warp * CL + CP + gid] = 0.0;
        }

        // ---------------- pass 1: logits, chunk by chunk ----------------
        for (int c = 0; c < nch; ++c) {
            const int wc = min(KC, dimp - c * KC);
            const int ngroups = wc >> 4;
            const int ngw = dmma_warp ? (ngroups - warp + kDmmaWarps - 1) / kDmmaWarps : 0;
            double vreg[NBG * 4][NCT];
            double vex[NBG * 4];
            if (dmma_warp) {
                // V fragments of this warp's column groups (constant for all tiles of the chunk)
                double pb[NCT], pbx = 0.0;
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
                    pb[ct] = 0.0;
#pragma unroll
                for (int i = 0; i < NBG; ++i) {
                    const int g = warp + kDmmaWarps * i;
                    const int col = c * KC + 16 * g + 4 * qd;
                    const bool ok = g < ngroups;
                    double2 s01 = make_double2(0.0, 0.0), s23 = s01, m01 = s01, m23 = s01;
                    if (ok) {
                        s01 = __ldg(reinterpret_cast<const double2 *>(P.cinvsd + col));
                        s23 = __ldg(reinterpret_cast<const double2 *>(P.cinvsd + col + 2));
                        m01 = __ldg(reinterpret_cast<const double2 *>(P.cmean + col));
                        m23 = __ldg(reinterpret_cast<const double2 *>(P.cmean + col + 2));
                    }
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        const int cls = ct * 8 + gid;
                        double2 w01 = make_double2(0.0, 0.0), w23 = w01;
                        if (ok && cls < P.C) {
                            const double *wg = xcur + (size_t)cls * dimp + col;
                            w01 = __ldcg(reinterpret_cast<const double2 *>(wg));
                            w23 = __ldcg(reinterpret_cast<const double2 *>(wg + 2));
                        }
                        vreg[4 * i + 0][ct] = w01.x * s01.x;
                        vreg[4 * i + 1][ct] = w01.y * s01.y;
                        vreg[4 * i + 2][ct] = w23.x * s23.x;
                        vreg[4 * i + 3][ct] = w23.y * s23.y;
                        pb[ct] += m01.x * vreg[4 * i + 0][ct];
                        pb[ct] += m01.y * vreg[4 * i + 1][ct];
                        pb[ct] += m23.x * vreg[4 * i + 2][ct];
                        pb[ct] += m23.y * vreg[4 * i + 3][ct];
                    }
                    if (EX) {
                        // the extra class (index CP) at this lane's four columns: the same for all gid
                        double2 w01 = make_double2(0.0, 0.0), w23 = w01;
                        if (ok) {
                            const double *wg = xcur + (size_t)CP * dimp + col;
                            w01 = __ldcg(reinterpret_cast<const double2 *>(wg));
                            w23 = __ldcg(reinterpret_cast<const double2 *>(wg + 2));
                        }
                        vex[4 * i + 0] = w01.x * s01.x;
                        vex[4 * i + 1] = w01.y * s01.y;
                        vex[4 * i + 2] = w23.x * s23.x;
                        vex[4 * i + 3] = w23.y * s23.y;
                        pbx += m01.x * vex[4 * i + 0];
                        pbx += m01.y * vex[4 * i + 1];
                        pbx += m23.x * vex[4 * i + 2];
                        pbx += m23.y * vex[4 * i + 3];
                    } else {
                        vex[4 * i + 0] = vex[4 * i + 1] = vex[4 * i + 2] = vex[4 * i + 3] = 0.0;
                    }
                }
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    double v = pb[ct];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (qd == 0)
                        so[warp * CL + ct * 8 + gid] += v;   // this warp's own slot, summed over the chunks
                }
                if (EX) {
                    pbx += __shfl_xor_sync(0xffffffffu, pbx, 1);
                    pbx += __shfl_xor_sync(0xffffffffu, pbx, 2);
                    if (lane == 0)
                        so[warp * CL + CP] += pbx;
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
            const bool full = ngw == NBG;
            for (int st = 0; st <= nst; ++st) {
                if (dmma_warp) {
                    if (st < nst) {
                        const unsigned sb = stage_addr(st) + off1;
                        const unsigned pw = plog_s + (unsigned)(st & 1) * kPlogHalf + offp;
                        if (full)
                            pass1_stage<true>(sb, pw, vreg, vex, ngw);
                        else
                            pass1_stage<false>(sb, pw, vreg, vex, ngw);
                    }
                } else {
                    // helper warps: next fetch, then the logits of the PREVIOUS stage += its 8 K-slices
                    if (st + 2 < nst)
                        fetch(c, st + 2, false, Lb);
                    else
                        cp_async_commit();
                    if (st >= 1)
                        accumulate(c, st - 1, Lb);
                    cp_async_wait<1>();
                }
                __syncthreads();
            }
        }
        // offsets o[c] = sum over the 8 warps' slots (fixed order)
        if (threadIdx.x < CL) {
            double s = 0.0;
            for (int w8 = 0; w8 < kDmmaWarps; ++w8)
                s += so[w8 * CL + threadIdx.x];
            o_all[threadIdx.x] = s;
        }
        __syncthreads();

        // ---------------- softmax over my rows: logits -> residuals, NLL, class sums ----------------
        // (the logits were accumulated in L2 by RED: read them past L1)
        double facc = 0.0;
        double rsum[CL];
#pragma unroll
        for (int cc = 0; cc < CL; ++cc)
            rsum[cc] = 0.0;
        {
            const int *yb = P.y + (size_t)blk * P.nb;
            const double *wb = P.w + (size_t)blk * P.nb;
            for (int row = r0 + threadIdx.x; row < r1; row += kThreads) {
                double *lp = Lb + (size_t)row * CL;
                double lv[CL];
                double mx = -INFINITY;
#pragma unroll
                for (int cc = 0; cc < CL; cc += 2) {
                    const double2 l2 = __ldcg(reinterpret_cast<const double2 *>(lp + cc));
                    lv[cc] = l2.x - o_all[cc];
                    lv[cc + 1] = l2.y - o_all[cc + 1];
                }
#pragma unroll
                for (int cc = 0; cc < CL; ++cc)
                    if (cc < P.C)
                        mx = fmax(mx, lv[cc]);
                const int cls = yb[row];
                const double wi = wb[row];
                const double lcls = __ldcg(lp + cls) - o_all[cls];
                double den = 0.0;
#pragma unroll
                for (int cc = 0; cc < CL; ++cc) {
                    lv[cc] = (cc < P.C) ? exp(lv[cc] - mx) : 0.0;
                    den += lv[cc];
                }
                facc -= wi * ((lcls - mx) - log(den));
#pragma unroll
                for (int cc = 0; cc < CL; cc += 2) {
                    double ra = 0.0, rb = 0.0;
                    if (cc < P.C) {
                        ra = wi * (lv[cc] / den);
                        if (cc == cls)
                            ra -= wi;
                    }
                    if (cc + 1 < P.C) {
                        rb = wi * (lv[cc + 1] / den);
                        if (cc + 1 == cls)
                            rb -= wi;
                    }
                    *reinterpret_cast<double2 *>(lp + cc) = make_double2(ra, rb);
                    rsum[cc] += ra;
                    rsum[cc + 1] += rb;
                }
            }
        }
#pragma unroll
        for (int cc = 0; cc < CL; ++cc) {
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
            double aex[NBG][2];
#pragma unroll
            for (int i = 0; i < NBG; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    aex[i][h] = 0.0;
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct)
                        acc[i][h][ct][0] = acc[i][h][ct][1] = 0.0;
                }
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
            const bool full = ngw == NBG;
            for (int st = 0; st < nst; ++st) {
                if (dmma_warp) {
                    const unsigned sa = stage_addr(st);
                    if (full)
                        pass2_stage<true>(sa + off2, sa + offr, sa + offx, acc, aex, ngw);
                    else
                        pass2_stage<false>(sa + off2, sa + offr, sa + offx, acc, aex, ngw);
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
                    if (EX) {
                        // extra class: the lanes of a quad hold the partial sums of rows qd, 4 + qd of
                        // every tile for columns 16 g + 2 gid + {0, 1}; lane qd = 0 stores the total
                        double v0 = aex[i][0], v1 = aex[i][1];
                        v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
                        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                        v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
                        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
                        if (g < ngroups && qd == 0) {
                            const int cx = c * KC + 16 * g + 2 * gid;
                            const double2 sx = __ldg(reinterpret_cast<const double2 *>(P.cinvsd + cx));
                            const double2 mx2 = __ldg(reinterpret_cast<const double2 *>(P.cmean + cx));
                            const double sc = sc_all[CP];
                            __stcg(reinterpret_cast<double2 *>(gp + (size_t)CP * dimp + cx),
                                   make_double2((v0 - mx2.x * sc) * sx.x, (v1 - mx2.y * sc) * sx.y));
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
