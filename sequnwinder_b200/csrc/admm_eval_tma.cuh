// admm_eval_tma.cuh — fused single-pass (objective, gradient) evaluation for row lengths that fit
// shared memory: the fast path of the x-update kernel.
//
// Same mathematics as admm_eval.cuh (LOne.java:1007-1029, 941-979); different data movement:
//   * the CTA's rows travel HBM -> shared memory exactly once per evaluation as 8-row tiles moved
//     by the TMA engine (cp.async.bulk, one bulk copy per tile: rows are contiguous in the
//     block-major layout) into an NS-deep ring guarded by mbarriers; no thread ever waits on a
//     global load, so eight warps per SM are enough to keep ~170 KB in flight;
//   * both passes read the tile from shared memory with 16-byte LDS: pass 1 splits the K (column)
//     range over the 8 warps (partial logits are summed through a 4 KB scratch), pass 2 splits
//     the column groups over the warps with the gradient accumulators pinned in registers for
//     the whole evaluation;
//   * the ring keeps running across evaluations: while the solver's reduce / decide phases run,
//     the first NS tiles of the next evaluation are already in flight; if the CTA's rows fit the
//     ring entirely they are loaded once per kernel and every later evaluation runs from
//     shared memory.
// Row stride dimp is kept = 8 (mod 16) doubles so that the 8-row x 64-byte fragment reads of
// pass 1 and the 8-class x 64-byte reads of W hit all 32 banks (4 wavefronts per LDS.128).
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes,
                                         unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

template <int NCT>
struct TileEvalCfg {
    static constexpr int NBG = (NCT <= 2) ? 6 : 4;  // 16-column groups per warp held in registers
};

// Shared-memory carve-up of the x-update kernel (fast path); offsets in bytes.
struct TileSmemLayout {
    size_t w_off, w_bytes;         // W [Cp][dimp]; doubles as the reduce scratch of the solver
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
    L.w_bytes = std::max((size_t)Cp * dimp * 8, min_w_bytes);
    L.w_bytes = (L.w_bytes + 127) & ~(size_t)127;
    L.stage_bytes = (size_t)kTileRows * dimp * 8;
    const size_t fixed = L.w_bytes + (size_t)kWarps * nct * 32 * 16 + (size_t)kWarps * nct * 64 * 8 + 128;
    int ns = 0;
    if (budget > fixed)
        ns = (int)std::min<size_t>(4, (budget - fixed) / L.stage_bytes);
    L.NS = ns;
    L.stage_off = L.w_bytes;
    L.plog_off = L.stage_off + (size_t)ns * L.stage_bytes;
    L.rs_off = L.plog_off + (size_t)kWarps * nct * 32 * 16;
    L.bar_off = L.rs_off + (size_t)kWarps * nct * 64 * 8;
    L.total = L.bar_off + 128;
    return L;
}

template <int NCT>
struct TileEval {
    static constexpr int NBG = TileEvalCfg<NCT>::NBG;
    const AdmmDev &P;
    int blk, cta;
    double *smW, *stages, *rs;
    double2 *plog;
    unsigned long long *full;
    double *red;
    int NS, r0, r1, ntiles;
    bool resident;
    long long q;       // sequence number of the next tile to consume (streaming mode)
    int evals_done;
    const double *Xb;

    __device__ void issue(long long qq)
    {
        const int tile = (int)(qq % ntiles), s = (int)(qq % NS);
        const int row = r0 + tile * kTileRows;
        const int rows = min(kTileRows, r1 - row);
        const unsigned bytes = (unsigned)rows * (unsigned)P.dimp * 8u;
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(stages + (size_t)s * kTileRows * P.dimp, Xb + (size_t)row * P.dimp, bytes, &full[s]);
    }

    __device__ void init(unsigned char *smem, const TileSmemLayout &L, double *red_)
    {
        smW = reinterpret_cast<double *>(smem + L.w_off);
        stages = reinterpret_cast<double *>(smem + L.stage_off);
        plog = reinterpret_cast<double2 *>(smem + L.plog_off);
        rs = reinterpret_cast<double *>(smem + L.rs_off);
        full = reinterpret_cast<unsigned long long *>(smem + L.bar_off);
        red = red_;
        NS = L.NS;
        cta_row_range(P, cta, r0, r1);
        ntiles = (r1 - r0 + kTileRows - 1) / kTileRows;
        if (ntiles < 0)
            ntiles = 0;
        resident = ntiles <= NS;
        q = 0;
        evals_done = 0;
        Xb = P.X + (size_t)blk * P.nb * P.dimp;
        // stale rows of a partially filled stage must stay finite (0 * NaN would poison pass 2)
        for (size_t i = threadIdx.x; i < (size_t)NS * kTileRows * P.dimp; i += kThreads)
            stages[i] = 0.0;
        if (threadIdx.x == 0) {
            for (int s = 0; s < NS; ++s)
                mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // zero fill before TMA writes
        __syncthreads();
        if (threadIdx.x == 0 && ntiles > 0) {
            const int first = resident ? ntiles : NS;
            for (int i = 0; i < first; ++i)
                issue(i);
        }
    }

    // Outstanding bulk copies must land before the CTA may exit.
    __device__ void drain()
    {
        if (ntiles == 0 || resident) {
            if (resident && evals_done == 0)
                for (int s = 0; s < ntiles; ++s)
                    mbar_wait(&full[s], 0);
            return;
        }
        for (int i = 0; i < NS; ++i) {
            const long long qq = q + i;
            mbar_wait(&full[qq % NS], (unsigned)((qq / NS) & 1));
        }
    }

    __device__ double eval(const double *xcur)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int qd = lane & 3, gid = lane >> 2;
        const int dimp = P.dimp, Cp = P.Cp;
        // W -> shared memory (rows >= C are zero)
        for (int i = threadIdx.x; i < Cp * dimp; i += kThreads)
            smW[i] = (i < P.npad) ? __ldcg(xcur + i) : 0.0;
        __syncthreads();

        const int nkp = dimp >> 3;
        const int kp0 = nkp * warp / kWarps, kp1 = nkp * (warp + 1) / kWarps;
        const int ngroups = (dimp + 15) >> 4;
        const int *yb = P.y + (size_t)blk * P.nb;
        const double *wb = P.w + (size_t)blk * P.nb;

        double acc[NBG][2][NCT][2];
#pragma unroll
        for (int i = 0; i < NBG; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
                    acc[i][h][ct][0] = acc[i][h][ct][1] = 0.0;
        double facc = 0.0;

        for (int t = 0; t < ntiles; ++t) {
            int s;
            if (resident) {
                s = t;
                if (evals_done == 0)
                    mbar_wait(&full[s], 0);
            } else {
                const long long qq = q + t;
                s = (int)(qq % NS);
                mbar_wait(&full[s], (unsigned)((qq / NS) & 1));
            }
            const double *tile = stages + (size_t)s * kTileRows * dimp;

            // ---- pass 1: partial logits of the 8 rows over this warp's k range
            {
                double pa[NCT][2][2];
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
                    pa[ct][0][0] = pa[ct][0][1] = pa[ct][1][0] = pa[ct][1][1] = 0.0;
                const double2 *xp = reinterpret_cast<const double2 *>(tile + (size_t)gid * dimp) + qd;
                const double2 *wp = reinterpret_cast<const double2 *>(smW + (size_t)gid * dimp) + qd;
#pragma unroll 4
                for (int kp = kp0; kp < kp1; ++kp) {
                    const double2 xa = xp[4 * kp];
#pragma unroll
                    for (int ct = 0; ct < NCT; ++ct) {
                        const double2 wv = wp[(size_t)ct * 4 * dimp + 4 * kp];
                        dmma884(pa[ct][0], xa.x, wv.x);
                        dmma884(pa[ct][1], xa.y, wv.y);
                    }
                }
#pragma unroll
                for (int ct = 0; ct < NCT; ++ct)
                    plog[(warp * NCT + ct) * 32 + lane] =
                        make_double2(pa[ct][0][0] + pa[ct][1][0], pa[ct][0][1] + pa[ct][1][1]);
            }
            __syncthreads();

            // ---- logits -> softmax -> residual fragments (every warp, redundantly)
            const int row = r0 + t * kTileRows + gid;
            const bool rv = row < r1;
            const int rowc = rv ? row : r0;
            double sv[NCT][2];
            double mx = -INFINITY;
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                double2 v = make_double2(0.0, 0.0);
#pragma unroll
                for (int w8 = 0; w8 < kWarps; ++w8) {
                    const double2 pv = plog[(w8 * NCT + ct) * 32 + lane];
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
                for (int j = 0; j < 2; ++j) {
                    ex[ct][j] = (ct * 8 + 2 * qd + j < P.C) ? exp(sv[ct][j] - mx) : 0.0;
                    den += ex[ct][j];
                }
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            const int cls = yb[rowc];
            const double wi = wb[rowc];
            const double logden = log(den);
            double *myrs = rs + (size_t)warp * NCT * 64;
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                double rr[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = ct * 8 + 2 * qd + j;
                    double r = wi * (ex[ct][j] / den);
                    if (c == cls) {
                        r -= wi;
                        if (rv && warp == 0)
                            facc -= wi * ((sv[ct][j] - mx) - logden);
                    }
                    rr[j] = rv ? r : 0.0;
                }
                *reinterpret_cast<double2 *>(myrs + ct * 64 + gid * 8 + 2 * qd) =
                    make_double2(rr[0], rr[1]);
            }
            __syncwarp();
            double a[NCT][2];
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
                a[ct][0] = myrs[ct * 64 + qd * 8 + gid];        // R[row qd][class gid]
                a[ct][1] = myrs[ct * 64 + (4 + qd) * 8 + gid];  // R[row 4+qd][class gid]
            }
            __syncwarp();

            // ---- pass 2: G[:, my column groups] += R^T X
#pragma unroll
            for (int i = 0; i < NBG; ++i) {
                const int grp = warp + kWarps * i;
                if (grp < ngroups) {
                    const int col = 16 * grp + 2 * gid;
                    const bool cok = col < dimp;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        double2 xb = make_double2(0.0, 0.0);
                        if (cok)
                            xb = *reinterpret_cast<const double2 *>(tile + (size_t)(4 * ks + qd) * dimp + col);
#pragma unroll
                        for (int ct = 0; ct < NCT; ++ct) {
                            dmma884(acc[i][0][ct], a[ct][ks], xb.x);
                            dmma884(acc[i][1][ct], a[ct][ks], xb.y);
                        }
                    }
                }
            }
            __syncthreads();  // every warp is done with this stage and with plog
            if (!resident && threadIdx.x == 0)
                issue(q + t + NS);
        }
        if (!resident)
            q += ntiles;
        evals_done++;

        // ---- partial gradient of this CTA
        double *gp = P.gpart + ((size_t)blk * P.G + cta) * P.npad;
#pragma unroll
        for (int i = 0; i < NBG; ++i) {
            const int grp = warp + kWarps * i;
            const int col = 16 * grp + 4 * qd;
            if (grp < ngroups && col < dimp) {
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
        return block_sum(facc, red);
    }
};

}  // namespace sqw
