"""Prints the main fields of a bench.py JSON line: python tools/show_bench.py file.json"""
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("value", "n_gpus", "ms_per_step", "time_to_converge_s", "evals_per_step", "gpu_launches",
          "kmer_gbases_per_s"):
    print(k, l.get(k))
for k in ("kmer_roofline", "roofline", "e2e", "cpu_baseline", "clocks"):
    print(k, l.get(k))
