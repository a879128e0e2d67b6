"""Cost model of the ways to run BASELINE config 5 (16 lambda values on the cfg-2 matrix) on one GPU, on
the recorded per-block evaluation counts of a cfg-2 run (profiles/r02_nfev_cfg2.npy, written by tools/gpu_nfev_dump.py):
 * one training after the other on the whole GPU (checks the model: 814 us per iteration against 792 measured),
 * trainings side by side with static CTA budgets (what sweep_concurrent does),
 * ONE launch series over all (lambda, block) pairs with count-based, reproducible re-grouping (not built).
A trip of a group of G CTAs costs eval(rows per CTA) + chain(G); a phase lasts as long as its slowest
block. eval(r): the resident packed kernel 10.4 + 0.0816 r us (fit to SQW_ADMM_PROFILE: 21.7 us at 139
rows, 14.6 at 52); the column-chunked kernel fitted to the measured 2- and 3-slot sweeps. What the model
leaves out is what the budgeted runs then showed (profiles/r02_phase_profile_budget49_mode12.log): with
fewer than 17 CTAs per block a CTA's slice of the solver vectors no longer fits the slice cache and the
reduce phase grows from 4.6 to 25 us, so the measured sweeps (31.4 / 29.3 s with 3 / 4 slots) are slower than
the rows below (20-24 s) - and a single-launch batched solve would meet the same wall. DESIGN.md 2.5."""
import numpy as np, sys
nf = np.load(sys.argv[1] if len(sys.argv) > 1 else 'profiles/r02_nfev_cfg2.npy').astype(int)
ROWS=2500; CAPS=(8,12,16,20,26,10**9)
def chain(G): return 13.1 + 0.07*(G-18) if G>1 else 6.0   # us; G=1: no cross-CTA barriers (guess 6)
def ev8(r): return 10.4 + 0.0816*r
def make_ev10(e0,e1): return lambda r: e0 + e1*r
def iter_cost(counts, ncta, ev, gcap=48, launch=5.0, caps=CAPS, waves=False):
    """counts: eval counts of the blocks sharing the launch series; returns us for the x-update"""
    tot=0.0; done=0
    for cap in caps:
        act=int((counts>done).sum())
        if act==0: tot+=3.0; done=cap; continue
        trips=min(cap,counts.max())-done
        best=None
        for nw in ((1,2,3,4) if waves else (1,)):
            per=-(-act//nw)
            G=min(ncta//per, gcap)
            if G<1: continue
            r=-(-ROWS//G)
            t=nw*(launch+trips*(ev(r)+chain(G)))
            if best is None or t<best: best=t
        tot+=best; done=cap
    return tot
# 1. solo mode 8 check
solo=np.mean([iter_cost(nf[i],148,ev8) for i in range(0,3600,7)])
print('solo mode8 model %.0f us/iter (measured 792)'%solo)
# 2. fit mode 10 to 3-slot (49 CTAs: 1780 us) and 2-slot (74 CTAs: 1230 us)
best=None
for e0 in np.arange(6,30,1.0):
    for e1 in np.arange(0.06,0.25,0.005):
        ev=make_ev10(e0,e1)
        a=np.mean([iter_cost(nf[i],49,ev) for i in range(0,3600,37)])
        b=np.mean([iter_cost(nf[i],74,ev) for i in range(0,3600,37)])
        err=(a/1780-1)**2+(b/1230-1)**2
        if best is None or err<best[0]: best=(err,e0,e1,a,b)
print('mode10 fit e0 %.1f e1 %.3f -> 3-slot %.0f (1780) 2-slot %.0f (1230)'%best[1:])
ev10=make_ev10(best[1],best[2])
# 3. batched: Q lambdas x 8 blocks in one launch series
for Q in (12,16):
  for waves in (False,True):
    for name,ev in (('mode10',ev10),('mode8-like',ev8)):
        c=[]
        for i in range(0,3600-Q,37):
            counts=nf[i:i+Q].ravel()
            c.append(iter_cost(counts,148,ev,waves=waves))
        per=np.mean(c)/Q
        print('Q=%d waves=%s %s: %.0f us per round, %.0f us per lambda-iteration -> 50000 iters %.1f s'%(Q,waves,name,np.mean(c),per,per*50000/1e6))
print('--- budgeted handles with a mode-8-like streamed kernel (0.087 us/row)')
ev12=lambda r: 10.4+0.087*r
for slots in (1,2,3,4,6,8):
    n=148//slots
    a=np.mean([iter_cost(nf[i],n,ev12) for i in range(0,3600,37)])
    print('slots %d (%d CTAs): %.0f us/iter per handle -> %.0f us per lambda-iter -> sweep %.1f s'%(slots,n,a,a/slots,a/slots*50000/1e6))
print('--- budgeted handles with per-panel streamed kernel (12 + 0.11 us/row)')
ev12=lambda r: 12+0.11*r
for slots in (2,3,4,6):
    n=148//slots
    a=np.mean([iter_cost(nf[i],n,ev12) for i in range(0,3600,37)])
    print('slots %d (%d CTAs): %.0f us/iter per handle -> %.0f us per lambda-iter -> sweep %.1f s'%(slots,n,a,a/slots,a/slots*50000/1e6))
