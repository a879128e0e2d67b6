"""Cost model of the phased x-update on recorded evaluation counts (gpurun_out/nfev_cfg2.npy from
tools/gpu_nfev_dump.py): what a schedule of evaluation caps costs per ADMM iteration.
Model (fitted to SQW_ADMM_PROFILE at cfg-2): a trip of a group of G CTAs costs
eval(G) + solver(G); a launch costs a fixed overhead; a phase lasts as long as its slowest block."""
import sys, itertools, numpy as np
nf = np.load(sys.argv[1] if len(sys.argv) > 1 else 'profiles/r02_nfev_cfg2.npy').astype(int)
NB, NCTA, GCAP, ROWS = nf.shape[1], 144, 48, 2500
TRIP = np.zeros(NB + 1)
for act in range(1, NB + 1):
    G = min(NCTA // act, GCAP)
    tiles = -(-(-(-ROWS // G)) // 8)
    TRIP[act] = 3.0 + 1.3 * tiles + 15.5 + 0.07 * (G - 18)

def cost(caps, rows, launch=6.0, empty=3.0):
    tot = np.zeros(len(rows))
    done_at = 0
    for cap in caps:
        act = (rows > done_at).sum(1)
        trips = np.minimum(cap, rows.max(1)) - done_at
        tot += np.where(act > 0, launch + np.maximum(trips, 0) * TRIP[act], empty)
        done_at = cap
    return tot.mean()

base = (8, 12, 16, 20, 26, 10**9)
print('current caps', base, '-> %.1f us per iteration (measured ~884)' % cost(base, nf))
if len(sys.argv) > 2:
    sub = nf[::7]
    best = []
    cands = [4, 5, 6, 7, 8, 10, 12, 14, 16, 18, 20, 23, 26, 30]
    for k in (5, 6, 7, 8):
        for c in itertools.combinations(cands, k):
            best.append((cost(c + (10**9,), sub), c))
    best.sort()
    for b in best[:6]:
        print('%.1f (all rows %.1f)' % (b[0], cost(b[1] + (10**9,), nf)), b[1])
    for k in (5, 6, 7, 8):
        bk = min(b for b in best if len(b[1]) == k)
        print(k + 1, 'phases: %.1f' % cost(bk[1] + (10**9,), nf), bk[1])
