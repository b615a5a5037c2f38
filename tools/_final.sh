# round-end validation + profiles (one GPU call): full GPU test suite, bench, ncu launch list of a tracked frame, ncu --set full of the two big kernels
cd /root/repo
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 400 python bench.py --steps 200 --warmup 10 2>/dev/null | tail -1 > gpurun_out/bench_r01j.json; cat gpurun_out/bench_r01j.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01j_frame.csv python tools/prof_frame.py 3 > gpurun_out/ncu_frame_r01j.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sd_tc_kernel|conv_tc" -s 86 -c 3 -o gpurun_out/prof_frame_r01j -f python tools/prof_frame.py 3 > gpurun_out/ncu_full_r01j.log 2>&1
ncu -i gpurun_out/prof_frame_r01j.ncu-rep --page raw --csv > gpurun_out/prof_frame_r01j_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
