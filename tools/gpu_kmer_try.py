import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import pyoracle as O
from sequnwinder_b200 import synth, KmerCounter
import torch
for name, n in [("cfg1", None), ("cfg2", None), ("cfg3", 20000), ("cfg4", 50000)]:
    cfg = synth.CONFIGS[name]
    d = synth.generate_config(name, n, n_frac=0.005)
    kc = KmerCounter(cfg['kmin'], cfg['kmax'])
    t0 = time.time(); ref, refn = O.kmer_count(d.bases, d.offsets, cfg['kmin'], cfg['kmax']); t_cpu = time.time()-t0
    t0 = time.time(); got, gotn = kc.count(d.bases, d.offsets); t_gpu = time.time()-t0
    print(name, got.shape, 'exact', np.array_equal(ref, got), np.array_equal(refn, gotn), 'cpu %.3fs gpu(host api) %.3fs' % (t_cpu, t_gpu), flush=True)
    b = torch.from_numpy(d.bases).cuda(); o = torch.from_numpy(d.offsets).cuda()
    out, hn, st = kc.count_device(b, o, 150)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): kc.count_device(b, o, 150, out=out, has_n=hn, status=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    nbytes = d.n_sites*(150 + 4*kc.numK)
    print('   dev exact', bool((out.cpu().numpy()==ref).all()), 'kernel %.3f ms  %.1f GB/s  %.2f Gbases/s' % (ms, nbytes/ms/1e6, d.n_sites*150/ms/1e6), flush=True)
# k=8 fallback, odd lengths
seqs = ["ACGTACGTAGCTAGCTAGCTAGGATCGATCGATTTAGCAGCATCGACTAGCATCAGCAGCGCGCGATATATAT", "ACG", "", "ACGTNACGT", "acgtacgtacgt"]
bases, offs = O.pack_sequences([s.upper() for s in seqs])
for kmin,kmax in [(1,3),(8,8),(4,7),(2,9)]:
    kc = KmerCounter(kmin,kmax)
    ref, refn = O.kmer_count(bases, offs, kmin, kmax); got, gotn = kc.count(bases, offs)
    print('k',kmin,kmax,'exact', np.array_equal(ref,got), np.array_equal(refn,gotn))
