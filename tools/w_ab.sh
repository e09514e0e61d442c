python bench.py --steps 1000 --warmup 10 > gpurun_out/w13_bench_full.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/w13_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/w13_ncu_launches.log 2>&1
python tools/run_steps.py --steps 6 > gpurun_out/w13_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wstrip256 -s 3 -c 1 -f -o gpurun_out/prof_wstrip_v5 python tools/run_steps.py --steps 6 > gpurun_out/w13_ncu.log 2>&1
python tools/bench_scoring.py > gpurun_out/w13_scoring.log 2>&1
python tools/bench_filterbank.py > gpurun_out/w13_fb.log 2>&1
