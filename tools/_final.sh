set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; tail -c 600 gpurun_out/bench_r1k.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r1k.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 900 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_r1k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xupdate -s 240 -c 1 -o gpurun_out/prof_xupdate_r1k -f python tools/gpu_admm_full.py cfg2 0 > gpurun_out/ncu_full_r1k.log 2>&1
tail -2 gpurun_out/ncu_full_r1k.log
