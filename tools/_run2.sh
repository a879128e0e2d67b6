set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/gpu_big_perf.py cfg3shape 40000 10921 12 2>&1 | tail -2
python tools/gpu_big_perf.py cfg4shape 100000 2729 17 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:admm_xupdate_kernel -s 1 -c 1 -o gpurun_out/prof_mode6_r1 -f python tools/gpu_big_perf.py cfg3shape 40000 10921 12 > gpurun_out/ncu_mode6.log 2>&1
tail -3 gpurun_out/ncu_mode6.log
