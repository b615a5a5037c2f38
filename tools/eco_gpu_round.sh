#!/bin/bash
# First GPU pass owed to the ECO row and the GNSteepestDescent seam -- the round's GPU budget ran out before it could run:
#   gpurun --timeout 900 -- 'bash tools/eco_gpu_round.sh r03a'
# 1. the -m gpu tests that have never run on a B200 (only the 12 golden tests of the online kernel have)
# 2. CUDA-event timings at ECO's default block sizes (gpurun_out/eco_bench.json; also what bench.py appends to rooflines[])
# 3. launch list + one `ncu --set full` capture of each ECO kernel, exported to CSV for profiles/
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r03a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_eco_gpu.py tests/test_zy_gnsd_gpu.py -q 2>&1 | tail -15 > gpurun_out/${TAG}_eco_gpu_tests.txt; cat gpurun_out/${TAG}_eco_gpu_tests.txt
timeout 300 python tools/eco_bench.py --json gpurun_out/${TAG}_eco_bench.json 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_eco_launches.csv \
    python tools/eco_bench.py --json /dev/null > gpurun_out/${TAG}_eco_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:eco_cg_kernel|eco_joint_kernel|eco_sample_fs_kernel|eco_preprocess_kernel|eco_apply_filter_kernel" -s 8 -c 12 -o gpurun_out/${TAG}_eco -f \
    python tools/eco_bench.py --json /dev/null > gpurun_out/${TAG}_eco_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_eco.ncu-rep --page raw --csv > gpurun_out/${TAG}_eco_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_eco.ncu-rep
ls -la gpurun_out/${TAG}_*
