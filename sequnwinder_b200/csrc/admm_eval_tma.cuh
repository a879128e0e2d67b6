// admm_eval_tma.cuh — fused single-pass (objective, gradient) evaluation for row lengths that fit
// shared memory: the fast path of the x-update kernel (MODE 4).
//
// Same mathematics as admm_eval.cuh (LOne.java:1007-1029, 941-979); different data movement:
//   * the CTA's rows travel HBM -> shared memory exactly once per evaluation as 8-row tiles moved
//     by the TMA engine (cp.async.bulk, one bulk copy per row, eight on one mbarrier) into an
//     NS-deep ring; no thread ever waits on a global load;
//   * both passes read the tile from shared memory; the ring keeps running across evaluations:
//     while the solver's reduce / decide phases run, the first NS tiles of the next evaluation
//     are already in flight; if the CTA's rows fit the ring entirely they are loaded once per
//     kernel and every later evaluation runs from shared memory.
// Row stride: dimp = 4 (mod 8) doubles, i.e. 32 or 96 bytes (mod 128). Pass 1 reads 8 rows x 4
// consecutive doubles per LDS.64 (half-warp = 4 rows x 32 B, offsets 0/32/64/96: conflict free),
// pass 2 reads 4 rows x 8 x 16 B per LDS.128 (quarter-warp = 4 rows x 2 x 16 B: conflict free).
// (A stride of 64 mod 128 serves pass 1 with LDS.128 but gives 2-way conflicts in pass 2:
// measured 1.6k instead of 0.7k cycles per tile.)
#pragma once

#include "admm_eval.cuh"

namespace sqw {

constexpr int kTileRows = 8;

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (TMA engine; SASS UBLKCP)
#ifndef SQW_X_EVICT_FIRST
#define SQW_X_EVICT_FIRST 1
#endif
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes,
                                         unsigned long long *bar)
{
#if SQW_X_EVICT_FIRST
    // the training matrix is streamed once per evaluation and never re-used from L2 (104 MB per
    // pass at cfg-2): mark its lines evict-first so that the group's exchange buffers (partial
    // gradients, scalars, trial point) are not pushed out to HBM between a write and its read
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
    return;
#endif
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Shared-memory carve-up of the x-update kernel (fast path); offsets in bytes.
struct TileSmemLayout {
    size_t w_off, w_bytes;         // W [Cp][dimp] (C > 8 only); with one class tile: the solver's slice cache
    size_t stage_off, stage_bytes; // NS stages of kTileRows x dimp doubles
    size_t plog_off;               // [kWarps][NCT][32] double2 partial logits
    size_t rs_off;                 // [kWarps][NCT][64] residual fragments
    size_t bar_off;                // NS mbarriers
    size_t total;
    int NS;
};

inline TileSmemLayout tile_smem_layout(int Cp, int dimp, size_t budget, size_t min_w_bytes)
{
    TileSmemLayout L{};
    const int nct = Cp / 8;
    L.w_off = 0;
    L.w_bytes = std::max(nct > 1 ? (size_t)Cp * dimp * 8 : (size_t)0, min_w_bytes);
    L.w_bytes = (L.w_bytes + 127) & ~(size_t)127;
    L.stage_bytes = (size_t)kTileRows * dimp * 8;
    const size_t plog_bytes = (size_t)2 * 8 * nct * 32 * 16;  // [2][kDmmaWarps][nct][32] double2
    const size_t rs_bytes = (size_t)2 * nct * 64 * 8;          // [2][nct][64] residual fragments
    const size_t fixed = L.w_bytes + plog_bytes + rs_bytes + 128;
    int ns = 0;
    if (budget > fixed)
        ns = (int)std::min<size_t>(4, (budget - fixed) / L.stage_bytes);
    L.NS = ns;
    L.stage_off = L.w_bytes;
    L.plog_off = L.stage_off + (size_t)ns * L.stage_bytes;
    L.rs_off = L.plog_off + plog_bytes;
    L.bar_off = L.rs_off + rs_bytes;
    L.total = L.bar_off + 128;
    return L;
}

}  // namespace sqw

// ----------------------------------------------------------------------------------------------
// Warp-specialised variant (MODE 4): 8 DMMA warps + 2 softmax warps + 1 TMA producer warp.
//
// Measured on B200 with the variants above (SQW_ADMM_PROFILE, cfg-2): of 72k cycles per
// evaluation only 25k are DMMA issue; 15k sit in the softmax dependency chain (LDS -> max ->
// exp -> sum -> div, fp64, ~1.7k cycles per tile, executed by every warp), 14k in waits for a
// single 42 KB bulk copy per tile, the rest in barriers. Here
//   * warps 0..7 only ever run DMMA sections: pass 1 of tile j, then pass 2 of tile j-1;
//   * warps 8 / 9 turn the partial logits of the even / odd tiles into residual fragments while
//     the DMMA warps are busy with the neighbouring tiles (double-buffered scratch, named
//     barriers bar.arrive / bar.sync between the roles) and own the NLL sums; one softmax warp
//     alone (~2k cycles of dependent fp64 exp / div / log per tile) would pace the whole pipeline;
//   * every tile is fetched as 8 per-row bulk copies on one mbarrier, four tiles deep.
namespace sqw {

constexpr int kDmmaWarps = 8;     // warps 0..7 issue DMMA only: two per SM sub-partition (the DMMA
                                  // pipe is per sub-partition, 16 cycles per m8n8k4)
constexpr int kSoftmaxWarps = 2;  // warps 8, 9: softmax of the even / odd tiles
constexpr int kProducerWarp = 10; // lane 0 re-fills released stages (TMA issue costs ~140 cycles
                                  // per bulk copy: kept off the DMMA warps)

__device__ __forceinline__ void nbar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int threads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct WsSmemLayout {
    static constexpr int kBarPl = 1, kBarRr = 3, kBarFree = 5;  // kBarFree + stage (<= 4 stages)
};

template <int NCT, int KSW_ = 24>
struct TileEvalWS {
    static constexpr int NBG = 6;  // 16-column groups per DMMA warp: rows up to 8*6*16 = 768 columns
    const AdmmDev &P;
    int blk, cta, G;
    double *smW, *stages, *rs;
    double2 *plog;
    unsigned long long *full;
    double *red;
    int NS, r0, r1, ntiles;
    bool resident;
    int cs, cph;   // streaming mode: stage / mbarrier parity of the next tile to consume
    int pt;        // streaming mode (producer lane): next tile index to fetch
    int evals_done;
    const double *Xb;
    long long *prof;

    // fetch tile `tile` of this CTA into stage s: one bulk copy per row, all on full[s]
    __device__ void issue(int tile, int s)
    {
        const int row = r0 + tile * kTileRows;
        const int rows = min(kTileRows, r1 - row);
        const unsigned row_bytes = (unsigned)P.dimp * 8u;
        mbar_expect_tx(&full[s], (unsigned)rows * row_bytes);
        double *dst = stages + (size_t)s * kTileRows * P.dimp;
        const double *src = Xb + (size_t)row * P.dimp;
        for (int r = 0; r < rows; ++r)
            bulk_g2s(dst + (size_t)r * P.dimp, src + (size_t)r * P.dimp, row_bytes, &full[s]);
    }

    __device__ void init(unsigned char *smem, const TileSmemLayout &L, double *red_)
    {
        smW = reinterpret_cast<double *>(smem + L.w_off);
        stages = reinterpret_cast<double *>(smem + L.stage_off);
        plog = reinterpret_cast<double2 *>(smem + L.plog_off);   // [2][7][NCT][32]
        rs = reinterpret_cast<double *>(smem + L.rs_off);        // [2][NCT][64]
        full = reinterpret_cast<unsigned long long *>(smem + L.bar_off);
        red = red_;
        NS = L.NS;
        cta_row_range(P, cta, G, r0, r1);
        ntiles = max(0, (r1 - r0 + kTileRows - 1) / kTileRows);
        resident = ntiles <= NS;
        cs = cph = 0;
        pt = 0;
        evals_done = 0;
        prof = nullptr;
        Xb = P.X + (size_t)blk * P.nb * P.dimp;
        for (size_t i = threadIdx.x; i < (size_t)NS * kTileRows * P.dimp; i += kThreads)
            stages[i] = 0.0;
        if (threadIdx.x == 0) {
            for (int s = 0; s < NS; ++s)
                mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == kProducerWarp * 32 && ntiles > 0) {
            const int first = resident ? ntiles : NS;
            for (int i = 0; i < first; ++i)
                issue(i, i);     // not resident: ntiles > NS, so tile i exists
            pt = resident ? 0 : (NS == ntiles ? 0 : NS);
        }
    }

    __device__ void set_profile(long long *p) { prof = p; }

    __device__ void drain()
    {
        if (ntiles == 0)
            return;
        if (resident) {
            if (evals_done == 0)
                for (int s = 0; s < ntiles; ++s)
                    mbar_wait(&full[s], 0);
            return;
        }
        int s = cs, ph = cph;
        for (int i = 0; i < NS; ++i) {
            mbar_wait(&full[s], (unsigned)ph);
            if (++s == NS) {
                s = 0;
                ph ^= 1;
            }
        }
    }

    // wait for the next tile of the sequence; returns its stage
    __device__ __forceinline__ int stage_wait(int j)
    {
        if (resident) {
            if (evals_done == 0)
                mbar_wait(&full[j], 0);
            return j;
        }
        const int s = cs;
        mbar_wait(&full[s], (unsigned)cph);
        if (++cs == NS) {
            cs = 0;
            cph ^= 1;
        }
        return s;
    }
    // re-fill the stage that tile processing just released (producer lane only)
    __device__ __forceinline__ void refill(int s)
    {
        issue(pt, s);
        if (++pt == ntiles)
            pt = 0;
    }

    __device__ double eval(const double *xcur)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int qd = lane & 3, gid = lane >> 2;
        const int dimp = P.dimp, Cp = P.Cp;
        long long tk0 = 0, tk1 = 0;
        const bool do_prof = prof != nullptr && threadIdx.x == 0;
        if (do_prof)
            tk0 = sqw_now_ns();
#define SQW_EPROF(slot)          \
    if (do_prof) {               \
        tk1 = sqw_now_ns();         \
        prof[slot] += tk1 - tk0; \
        tk0 = tk1;               \
    }
        constexpr bool W_REG = (NCT == 1);  // one class tile: this warp's W fragments live in registers
        constexpr int KSW = KSW_;           // k-steps (4 columns) per DMMA warp: dimp <= 8*KSW*4
        if (!W_REG) {
            for (int i = threadIdx.x; i < Cp * dimp; i += kThreads)
                smW[i] = (i < P.npad) ? __ldcg(xcur + i) : 0.0;
            __syncthreads();
        }
        SQW_EPROF(0)

        const int nks = dimp >> 2;
        const int ngroups = (dimp + 15) >> 4;
        double facc = 0.0;
        double acc[NBG][2][NCT][2];
#pragma unroll
        for (int i = 0; i < NBG; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
                    acc[i][h][ct][0] = acc[i][h][ct][1] = 0.0;

        if (warp < kDmmaWarps) {
            // =========================== DMMA warps ===========================
            const int ks0 = nks * warp / kDmmaWarps, ks1 = nks * (warp + 1) / kDmmaWarps;
            double wreg[W_REG ? KSW : 1];
            if (W_REG) {
                // W[class gid][4 ks + qd] straight from the global trial point (L2)
                const double *wg = xcur + (size_t)gid * dimp + qd;
#pragma unroll
                for (int u = 0; u < KSW; ++u) {
                    wreg[u] = 0.0;
                    if (ks0 + u < ks1 && gid < P.C)
                        wreg[u] = __ldcg(wg + 4 * (ks0 + u));
                }
            }
            auto pass2 = [&](int j, int sj) {
                const double *tile = stages + (size_t)sj * kTileRows * dimp;
                const double *myrs = rs + (size_t)(j & 1) * NCT * 64;
                double a[NCT][2];
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    a[ct][0] = myrs[ct * 64 + qd * 8 + gid];        // R[row qd][class gid]
                    a[ct][1] = myrs[ct * 64 + (4 + qd) * 8 + gid];  // R[row 4+qd][class gid]
                }
                // straight-line code: out-of-range column groups / columns read a valid address and
                // are multiplied by zero, so every LDS can be scheduled ahead of the DMMAs
                double2 x0[NBG], x1[NBG];
#pragma unroll
                for (int i = 0; i < NBG; ++i) {
                    const int col = 16 * (warp + kDmmaWarps * i) + 2 * gid;
                    const int colc = (col < dimp) ? col : 0;
                    x0[i] = *reinterpret_cast<const double2 *>(tile + (size_t)qd * dimp + colc);
                    x1[i] = *reinterpret_cast<const double2 *>(tile + (size_t)(4 + qd) * dimp + colc);
                    if (col >= dimp)
                        x0[i] = x1[i] = make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int i = 0; i < NBG; ++i) {
                    // only a warp's last group can lie beyond the row: one warp-uniform branch
                    if (i == NBG - 1 && warp + kDmmaWarps * i >= ngroups)
                        break;
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        dmma884(acc[i][0][ct], a[ct][0], x0[i].x);
                        dmma884(acc[i][1][ct], a[ct][0], x0[i].y);
                        dmma884(acc[i][0][ct], a[ct][1], x1[i].x);
                        dmma884(acc[i][1][ct], a[ct][1], x1[i].y);
                    }
                }
            };
            int sprev = 0;
            for (int j = 0; j < ntiles; ++j) {
                const int s = stage_wait(j);
                SQW_EPROF(1)
                const double *tile = stages + (size_t)s * kTileRows * dimp;
                {   // pass 1: partial logits over this warp's k range, 4 DMMA chains per class tile
                    double pa[NCT][4][2];
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            pa[ct][c][0] = pa[ct][c][1] = 0.0;
                    const double *xp = tile + (size_t)gid * dimp + qd;
                    if (W_REG) {
                        // two halves (register budget: 168 per thread with 11 warps): the loads of
                        // the second half are kept behind the first half's DMMAs by a compiler
                        // barrier; wreg[u] is zero beyond ks1, so clamped loads need no select
                        constexpr int H1 = (KSW + 1) / 2;
                        {
                            double xa[H1];
#pragma unroll
                            for (int u = 0; u < H1; ++u)
                                xa[u] = xp[4 * min(ks0 + u, ks1 - 1)];
#pragma unroll
                            for (int u = 0; u < H1; ++u)
                                dmma884(pa[0][u & 3], xa[u], wreg[u]);
                        }
                        asm volatile("" ::: "memory");
                        {
                            double xa[KSW - H1];
#pragma unroll
                            for (int u = H1; u < KSW; ++u)
                                xa[u - H1] = xp[4 * min(ks0 + u, ks1 - 1)];
#pragma unroll
                            for (int u = H1; u < KSW; ++u)
                                dmma884(pa[0][u & 3], xa[u - H1], wreg[u]);
                        }
                    } else {
                        const double *wp = smW + (size_t)gid * dimp + qd;
                        int ks = ks0;
                        for (; ks + 3 < ks1; ks += 4) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {  // static chain index: registers
                                const double xa = xp[4 * (ks + u)];
#pragma unroll
                                for (int ct = 0; ct < NCT; ++ct)
                                    dmma884(pa[ct][u], xa, wp[(size_t)ct * 8 * dimp + 4 * (ks + u)]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            if (ks + u < ks1) {
                                const double xa = xp[4 * (ks + u)];
#pragma unroll
                                for (int ct = 0; ct < NCT; ++ct)
                                    dmma884(pa[ct][u], xa, wp[(size_t)ct * 8 * dimp + 4 * (ks + u)]);
                            }
                        }
                    }
                    double2 *dst = plog + ((size_t)(j & 1) * kDmmaWarps + warp) * NCT * 32;
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct)
                        dst[ct * 32 + lane] =
                            make_double2((pa[ct][0][0] + pa[ct][1][0]) + (pa[ct][2][0] + pa[ct][3][0]),
                                         (pa[ct][0][1] + pa[ct][1][1]) + (pa[ct][2][1] + pa[ct][3][1]));
                }
                SQW_EPROF(2)
                nbar_arrive(WsSmemLayout::kBarPl + (j & 1), (kDmmaWarps + 1) * 32);  // partial logits of tile j ready
                if (j >= 1) {
                    nbar_sync(WsSmemLayout::kBarRr + ((j - 1) & 1), (kDmmaWarps + 1) * 32);  // residuals of tile j-1
                    SQW_EPROF(3)
                    pass2(j - 1, sprev);
                    SQW_EPROF(5)
                    if (!resident)  // tell the producer warp that the stage of tile j-1 is free
                        nbar_arrive(WsSmemLayout::kBarFree + sprev, (kDmmaWarps + 1) * 32);
                    SQW_EPROF(6)
                }
                sprev = s;
            }
            if (ntiles > 0) {
                nbar_sync(WsSmemLayout::kBarRr + ((ntiles - 1) & 1), (kDmmaWarps + 1) * 32);
                SQW_EPROF(3)
                pass2(ntiles - 1, sprev);
                SQW_EPROF(5)
                if (!resident)
                    nbar_arrive(WsSmemLayout::kBarFree + sprev, (kDmmaWarps + 1) * 32);
                SQW_EPROF(6)
            }
        } else if (warp == kProducerWarp) {
            // =========================== TMA producer ===========================
            if (!resident) {
                int s = cs;
                for (int j = 0; j < ntiles; ++j) {
                    nbar_sync(WsSmemLayout::kBarFree + s, (kDmmaWarps + 1) * 32);
                    if (lane == 0)
                        refill(s);
                    if (++s == NS)
                        s = 0;
                }
            }
        } else if (warp < kDmmaWarps + kSoftmaxWarps) {
            // =========================== softmax warp ===========================
            const int *yb = P.y + (size_t)blk * P.nb;
            const double *wb = P.w + (size_t)blk * P.nb;
            for (int j = warp - kDmmaWarps; j < ntiles; j += kSoftmaxWarps) {
                const int row = r0 + j * kTileRows + gid;
                const bool rv = row < r1;
                const int rowc = rv ? row : r0;
                const int cls = yb[rowc];
                const double wi = wb[rowc];
                nbar_sync(WsSmemLayout::kBarPl + (j & 1), (kDmmaWarps + 1) * 32);
                const double2 *src = plog + (size_t)(j & 1) * kDmmaWarps * NCT * 32;
                double sv[NCT][2];
                double mx = -INFINITY;
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    double2 v = src[ct * 32 + lane];
#pragma unroll
                    for (int w7 = 1; w7 < kDmmaWarps; ++w7) {
                        const double2 pv = src[(w7 * NCT + ct) * 32 + lane];
                        v.x += pv.x;
                        v.y += pv.y;
                    }
                    sv[ct][0] = v.x;
                    sv[ct][1] = v.y;
                    if (ct * 8 + 2 * qd < P.C)
                        mx = fmax(mx, v.x);
                    if (ct * 8 + 2 * qd + 1 < P.C)
                        mx = fmax(mx, v.y);
                }
                mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                double ex[NCT][2], den = 0.0;
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        ex[ct][jj] = (ct * 8 + 2 * qd + jj < P.C) ? exp(sv[ct][jj] - mx) : 0.0;
                        den += ex[ct][jj];
                    }
                den += __shfl_xor_sync(0xffffffffu, den, 1);
                den += __shfl_xor_sync(0xffffffffu, den, 2);
                double *myrs = rs + (size_t)(j & 1) * NCT * 64;
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct) {
                    double rr[2];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int c = ct * 8 + 2 * qd + jj;
                        double r = wi * (ex[ct][jj] / den);
                        if (c == cls)
                            r -= wi;
                        rr[jj] = rv ? r : 0.0;
                    }
                    *reinterpret_cast<double2 *>(myrs + ct * 64 + gid * 8 + 2 * qd) =
                        make_double2(rr[0], rr[1]);
                }
                __syncwarp();
                nbar_arrive(WsSmemLayout::kBarRr + (j & 1), (kDmmaWarps + 1) * 32);
                // the NLL term is off the critical path: after the residuals are published
                const double logden = log(den);
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
                        if (rv && ct * 8 + 2 * qd + jj == cls)
                            facc -= wi * ((sv[ct][jj] - mx) - logden);
            }
        }
        if (!resident && warp >= kDmmaWarps) {
            // only the DMMA warps wait on stages; keep everybody's copy of the ring position in step
            const int adv = cs + ntiles;
            cph ^= (adv / NS) & 1;
            cs = adv % NS;
        }
        evals_done++;
#if defined(SQW_EVAL_EPILOGUE_SYNC)
        __syncthreads();
#endif

        // ---- partial gradient of this CTA: every DMMA warp stores its accumulators as soon as its
        // last tile is done (no CTA-wide synchronisation here; the caller's barrier follows)
        if (warp < kDmmaWarps) {
            double *gp = P.gpart + ((size_t)blk * P.Gmax + cta) * P.npad;
#pragma unroll
            for (int i = 0; i < NBG; ++i) {
                const int cg = warp + kDmmaWarps * i;
                const int col = 16 * cg + 4 * qd;
                if (cg < ngroups && col < dimp) {
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        const int c = ct * 8 + gid;
                        if (c < P.C) {
                            double2 *dst = reinterpret_cast<double2 *>(gp + (size_t)c * dimp + col);
                            __stcg(dst, make_double2(acc[i][0][ct][0], acc[i][1][ct][0]));
                            __stcg(dst + 1, make_double2(acc[i][0][ct][1], acc[i][1][ct][1]));
                        }
                    }
                }
            }
        }
        SQW_EPROF(7)
#undef SQW_EPROF
        return facc;  // per-thread NLL partial (non-zero in the softmax warps only)
    }
};

}  // namespace sqw

// ----------------------------------------------------------------------------------------------
// Column-chunked variant (MODE 6): rows of any length (k = 4..6, 4..7 feature sets, up to 24
// classes). The row no longer fits shared memory, so the evaluation makes two TMA-staged passes
// over the CTA's rows, KC columns at a time:
//   pass 1  for every column chunk: partial logits of each 8-row tile (DMMA, K split over 8 warps)
//           are summed by three helper warps (one per class tile) into a logits array in L2
//           while the DMMA warps already work on the next tile (one __syncthreads per tile);
//   softmax over the CTA's rows (one row per thread): logits -> residuals w_i (p - y), NLL;
//   pass 2  for every column chunk: G[:, chunk] += R^T X[:, chunk], accumulators in registers for
//           the chunk, one store per chunk.
// X therefore crosses HBM twice per evaluation (the two-pass byte count of SURVEY.md 8d), always
// through the TMA ring, never through per-thread global loads.
namespace sqw {

template <int NCT>
struct ChunkCfg {
    static constexpr int NBG = (NCT == 1) ? 6 : (NCT == 2 ? 4 : 3);  // column groups per DMMA warp
    static constexpr int KC = 8 * NBG * 16 - 12;                      // 756 / 500 / 372 (= 4 mod 8)
    static constexpr int KSW = (KC / 4 + 7) / 8;                      // k-steps per DMMA warp
};

inline TileSmemLayout chunk_smem_layout(int Cp, size_t budget, size_t min_w_bytes)
{
    TileSmemLayout L{};
    const int nct = Cp / 8;
    const int KC = nct == 1 ? ChunkCfg<1>::KC : (nct == 2 ? ChunkCfg<2>::KC : ChunkCfg<3>::KC);
    L.w_off = 0;
    L.w_bytes = std::max((size_t)Cp * KC * 8, min_w_bytes);
    L.w_bytes = (L.w_bytes + 127) & ~(size_t)127;
    L.stage_bytes = (size_t)kTileRows * KC * 8;
    const size_t plog_bytes = (size_t)2 * 8 * nct * 32 * 16;
    const size_t fixed = L.w_bytes + plog_bytes + 128;
    int ns = 0;
    if (budget > fixed)
        ns = (int)std::min<size_t>(4, (budget - fixed) / L.stage_bytes);
    L.NS = ns;
    L.stage_off = L.w_bytes;
    L.plog_off = L.stage_off + (size_t)ns * L.stage_bytes;
    L.rs_off = L.plog_off + plog_bytes;
    L.bar_off = L.rs_off;
    L.total = L.bar_off + 128;
    return L;
}

template <int NCT>
struct TileEvalChunk {
    static constexpr int NBG = ChunkCfg<NCT>::NBG, KC = ChunkCfg<NCT>::KC, KSW = ChunkCfg<NCT>::KSW;
    const AdmmDev &P;
    int blk, cta, G;
    double *smW, *stages;
    double2 *plog;
    unsigned long long *full;
    double *red;
    int NS, r0, r1, ntiles, nch;
    int cs, cph;            // consumer ring position
    int p_pass, p_c, p_j;   // producer lane: next (pass, chunk, tile) to fetch
    const double *Xb;

    __device__ __forceinline__ int width(int c) const { return min(KC, P.dimp - c * KC); }

    __device__ void issue(int c, int tile, int s)
    {
        const int row = r0 + tile * kTileRows;
        const int rows = min(kTileRows, r1 - row);
        const unsigned row_bytes = (unsigned)width(c) * 8u;
        mbar_expect_tx(&full[s], (unsigned)rows * row_bytes);
        double *dst = stages + (size_t)s * kTileRows * KC;
        const double *src = Xb + (size_t)row * P.dimp + (size_t)c * KC;
        for (int r = 0; r < rows; ++r)
            bulk_g2s(dst + (size_t)r * KC, src + (size_t)r * P.dimp, row_bytes, &full[s]);
    }
    __device__ void refill(int s)
    {
        issue(p_c, p_j, s);
        if (++p_j == ntiles) {
            p_j = 0;
            if (++p_c == nch) {
                p_c = 0;
                p_pass ^= 1;
            }
        }
    }

    __device__ void init(unsigned char *smem, const TileSmemLayout &L, double *red_)
    {
        smW = reinterpret_cast<double *>(smem + L.w_off);
        stages = reinterpret_cast<double *>(smem + L.stage_off);
        plog = reinterpret_cast<double2 *>(smem + L.plog_off);
        full = reinterpret_cast<unsigned long long *>(smem + L.bar_off);
        red = red_;
        NS = L.NS;
        cta_row_range(P, cta, G, r0, r1);
        ntiles = max(0, (r1 - r0 + kTileRows - 1) / kTileRows);
        nch = (P.dimp + KC - 1) / KC;
        cs = cph = 0;
        p_pass = p_c = p_j = 0;
        Xb = P.X + (size_t)blk * P.nb * P.dimp;
        for (size_t i = threadIdx.x; i < (size_t)NS * kTileRows * KC; i += kThreads)
            stages[i] = 0.0;
        if (threadIdx.x == 0) {
            for (int s = 0; s < NS; ++s)
                mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == kProducerWarp * 32 && ntiles > 0)
            for (int i = 0; i < NS; ++i)   // 2 * nch * ntiles >= NS always holds for nch >= 2
                refill(i);
    }

    __device__ void set_profile(long long *) {}

    __device__ void drain()
    {
        if (ntiles == 0)
            return;
        int s = cs, ph = cph;
        for (int i = 0; i < NS; ++i) {
            mbar_wait(&full[s], (unsigned)ph);
            if (++s == NS) {
                s = 0;
                ph ^= 1;
            }
        }
    }

    __device__ __forceinline__ int stage_wait()
    {
        const int s = cs;
        mbar_wait(&full[s], (unsigned)cph);
        if (++cs == NS) {
            cs = 0;
            cph ^= 1;
        }
        return s;
    }

    __device__ double eval(const double *xcur)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int qd = lane & 3, gid = lane >> 2;
        const int dimp = P.dimp, Cp = P.Cp;
        double *Lb = P.R + (size_t)blk * P.nb * Cp;   // logits, then residuals, of this block's rows
        const int *yb = P.y + (size_t)blk * P.nb;
        const double *wb = P.w + (size_t)blk * P.nb;
        const bool producer = threadIdx.x == kProducerWarp * 32;

        // ---------------- pass 1: logits, chunk by chunk ----------------
        for (int c = 0; c < nch; ++c) {
            const int wc = width(c);
            __syncthreads();  // previous chunk's W (or the solver's scratch) is no longer in use
            for (int i = threadIdx.x; i < Cp * KC; i += kThreads) {
                const int cc = i / KC, k = i - cc * KC;
                smW[i] = (cc < P.C && k < wc) ? __ldcg(xcur + (size_t)cc * dimp + (size_t)c * KC + k) : 0.0;
            }
            __syncthreads();
            const int nks = (wc + 3) >> 2;
            const int ks0 = nks * min(warp, 8) / 8, ks1 = nks * min(warp + 1, 8) / 8;
            // helper warps: the logits accumulated so far are fetched one tile ahead, so the L2
            // round trip of the read-modify-write is off the per-tile critical path
            double2 old_next = make_double2(0.0, 0.0);
            const bool helper = warp >= 8 && warp - 8 < NCT;
            if (helper && c > 0 && r0 + gid < r1)
                old_next = *reinterpret_cast<const double2 *>(Lb + (size_t)(r0 + gid) * Cp + (warp - 8) * 8 + 2 * qd);
            for (int j = 0; j <= ntiles; ++j) {
                int s = -1;
                if (j < ntiles)
                    s = stage_wait();
                if (warp < 8 && j < ntiles) {
                    const double *tile = stages + (size_t)s * kTileRows * KC;
                    double pa[NCT][4][2];
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            pa[ct][q4][0] = pa[ct][q4][1] = 0.0;
                    const double *xp = tile + (size_t)gid * KC + qd;
                    const double *wp = smW + (size_t)gid * KC + qd;
                    // the chain index must be a compile-time constant (registers, not local memory):
                    // unrolled by 4 from this warp's first k-step, remainder handled below
                    int ks = ks0;
                    for (; ks + 3 < ks1; ks += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double xa = xp[4 * (ks + u)];
#pragma unroll
                            for (int ct = 0; ct < NCT; ++ct)
                                dmma884(pa[ct][u], xa, wp[(size_t)ct * 8 * KC + 4 * (ks + u)]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        if (ks + u < ks1) {
                            const double xa = xp[4 * (ks + u)];
#pragma unroll
                            for (int ct = 0; ct < NCT; ++ct)
                                dmma884(pa[ct][u], xa, wp[(size_t)ct * 8 * KC + 4 * (ks + u)]);
                        }
                    }
                    double2 *dst = plog + ((size_t)(j & 1) * 8 + warp) * NCT * 32;
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct)
                        dst[ct * 32 + lane] =
                            make_double2((pa[ct][0][0] + pa[ct][1][0]) + (pa[ct][2][0] + pa[ct][3][0]),
                                         (pa[ct][0][1] + pa[ct][1][1]) + (pa[ct][2][1] + pa[ct][3][1]));
                } else if (helper && j >= 1) {
                    // helper warp of class tile ct: logits of tile j-1 += sum of the 8 partials
                    const int ct = warp - 8;
                    const int row = r0 + (j - 1) * kTileRows + gid;
                    const double2 old = old_next;
                    if (c > 0 && row + kTileRows < r1)
                        old_next = *reinterpret_cast<const double2 *>(Lb + (size_t)(row + kTileRows) * Cp + ct * 8 + 2 * qd);
                    const double2 *src = plog + (size_t)((j - 1) & 1) * 8 * NCT * 32;
                    double2 v = src[ct * 32 + lane];
#pragma unroll
                    for (int w8 = 1; w8 < 8; ++w8) {
                        const double2 pv = src[(w8 * NCT + ct) * 32 + lane];
                        v.x += pv.x;
                        v.y += pv.y;
                    }
                    if (row < r1) {
                        double2 *lp = reinterpret_cast<double2 *>(Lb + (size_t)row * Cp + ct * 8 + 2 * qd);
                        if (c > 0) {
                            v.x += old.x;
                            v.y += old.y;
                        }
                        *lp = v;
                    }
                }
                __syncthreads();  // plog[j&1] complete; stage s and plog[(j-1)&1] free again
                if (producer && j < ntiles)
                    refill(s);
            }
        }

        // ---------------- softmax over my rows: logits -> residuals, NLL ----------------
        double facc = 0.0;
        for (int row = r0 + threadIdx.x; row < r1; row += kThreads) {
            double *lp = Lb + (size_t)row * Cp;
            double mx = -INFINITY;
            for (int cc = 0; cc < P.C; ++cc)
                mx = fmax(mx, lp[cc]);
            double den = 0.0;
            for (int cc = 0; cc < P.C; ++cc)
                den += exp(lp[cc] - mx);
            const int cls = yb[row];
            const double wi = wb[row];
            facc -= wi * ((lp[cls] - mx) - log(den));
            for (int cc = 0; cc < Cp; ++cc) {
                double r = 0.0;
                if (cc < P.C) {
                    r = wi * (exp(lp[cc] - mx) / den);
                    if (cc == cls)
                        r -= wi;
                }
                lp[cc] = r;
            }
        }
        __syncthreads();

        // ---------------- pass 2: gradient, chunk by chunk ----------------
        double *gp = P.gpart + ((size_t)blk * P.Gmax + cta) * P.npad;
        for (int c = 0; c < nch; ++c) {
            const int wc = width(c);
            const int ngroups = (wc + 15) >> 4;
            double acc[NBG][2][NCT][2];
#pragma unroll
            for (int i = 0; i < NBG; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct)
                        acc[i][h][ct][0] = acc[i][h][ct][1] = 0.0;
            // residual fragments R[row qd (+4)][class gid] are fetched one tile ahead (L2 latency)
            double a_next[NCT][2];
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                a_next[ct][0] = (warp < 8 && r0 + qd < r1) ? Lb[(size_t)(r0 + qd) * Cp + ct * 8 + gid] : 0.0;
                a_next[ct][1] = (warp < 8 && r0 + qd + 4 < r1) ? Lb[(size_t)(r0 + qd + 4) * Cp + ct * 8 + gid] : 0.0;
            }
            for (int j = 0; j < ntiles; ++j) {
                const int s = stage_wait();
                if (warp < 8) {
                    const double *tile = stages + (size_t)s * kTileRows * KC;
                    const int rowa = r0 + (j + 1) * kTileRows + qd, rowb = rowa + 4;
                    double a[NCT][2];
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        a[ct][0] = a_next[ct][0];
                        a[ct][1] = a_next[ct][1];
                        a_next[ct][0] = (rowa < r1) ? Lb[(size_t)rowa * Cp + ct * 8 + gid] : 0.0;
                        a_next[ct][1] = (rowb < r1) ? Lb[(size_t)rowb * Cp + ct * 8 + gid] : 0.0;
                    }
                    double2 x0[NBG], x1[NBG];
#pragma unroll
                    for (int i = 0; i < NBG; ++i) {
                        const int col = 16 * (warp + 8 * i) + 2 * gid;
                        const int colc = (col < wc) ? col : 0;
                        x0[i] = *reinterpret_cast<const double2 *>(tile + (size_t)qd * KC + colc);
                        x1[i] = *reinterpret_cast<const double2 *>(tile + (size_t)(4 + qd) * KC + colc);
                        if (col >= wc)
                            x0[i] = x1[i] = make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int i = 0; i < NBG; ++i)
#pragma unroll
                        for (int ct = 0; ct < NCT; ++ct) {
                            dmma884(acc[i][0][ct], a[ct][0], x0[i].x);
                            dmma884(acc[i][1][ct], a[ct][0], x0[i].y);
                            dmma884(acc[i][0][ct], a[ct][1], x1[i].x);
                            dmma884(acc[i][1][ct], a[ct][1], x1[i].y);
                        }
                }
                __syncthreads();
                if (producer)
                    refill(s);
            }
            if (warp < 8) {
#pragma unroll
                for (int i = 0; i < NBG; ++i) {
                    const int cg = warp + 8 * i;
                    const int col = 16 * cg + 4 * qd;
                    if (cg < ngroups && col < wc) {
#pragma unroll
                        for (int ct = 0; ct < NCT; ++ct) {
                            const int cc = ct * 8 + gid;
                            if (cc < P.C) {
                                double2 *dst = reinterpret_cast<double2 *>(gp + (size_t)cc * dimp +
                                                                           (size_t)c * KC + col);
                                __stcg(dst, make_double2(acc[i][0][ct][0], acc[i][1][ct][0]));
                                __stcg(dst + 1, make_double2(acc[i][0][ct][1], acc[i][1][ct][1]));
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
