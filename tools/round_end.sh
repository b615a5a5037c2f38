# round-end validation + profiles (one GPU call): full GPU test suite, bench, ncu launch list of tracked frames, ncu --set full of the two big kernels
cd /root/repo
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 400 python bench.py --steps 200 --warmup 10 2>/dev/null | tail -1 > gpurun_out/bench_r01j.json; cut -c1-330 gpurun_out/bench_r01j.json; python -c "import json; d=json.load(open('gpurun_out/bench_r01j.json')); print(d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['frac'])"
B200TRK_NET_FORK=0 timeout 400 python bench.py --steps 200 --warmup 10 2>/dev/null | tail -1 > gpurun_out/bench_r01j_nofork.json; python -c "import json; d=json.load(open('gpurun_out/bench_r01j_nofork.json')); print('nofork', d['value'], d['e2e']['value'], d['roofline']['us_per_launch'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2440 -c 330 --csv --log-file gpurun_out/launches_r01j_frame.csv python tools/prof_frame.py 3 > gpurun_out/ncu_frame_r01j.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sd_tc_kernel -s 2 -c 1 -o gpurun_out/prof_sd_tc_r01j -f python tools/prof_frame.py 3 > gpurun_out/ncu_full_sd_r01j.log 2>&1
ncu -i gpurun_out/prof_sd_tc_r01j.ncu-rep --page raw --csv > gpurun_out/prof_sd_tc_r01j_raw.csv 2>/dev/null
ls -la gpurun_out | tail -4
