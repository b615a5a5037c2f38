#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): launch list of three tracked frames + `ncu --set full` of one launch of
# every kernel family, exported to CSV on the box (the .ncu-rep files are deleted: gpurun_out/ must stay under 64 MiB).
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02k}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_frames.csv \
    python tools/profile_all.py frames > gpurun_out/${TAG}_ncu_frames.log 2>&1
SEL='sd_tc_kernel|sd_kernel|apply_filter_kernel|feat_transpose_kernel|sample_patch_kernel|localize_kernel|atom_cg_kernel|attention_kernel|prroi_|fourier_interp_kernel|fc_forward_kernel|fc_backward_kernel|tomp_tokens_kernel|groupnorm1_relu_kernel|stem_kernel|conv1x1_kernel|layernorm_kernel|max2d_kernel|feature_normalize_kernel|gn_txt_kernel|maxpool_kernel|import_scaled_kernel'
ncu --set full --clock-control none --profile-from-start off -k "regex:${SEL}" -c 130 -o gpurun_out/${TAG}_sel -f \
    python tools/profile_all.py > gpurun_out/${TAG}_ncu_sel.log 2>&1
ncu -i gpurun_out/${TAG}_sel.ncu-rep --page raw --csv > gpurun_out/${TAG}_sel_raw.csv 2>/dev/null
ncu --set full --clock-control none --profile-from-start off -k "regex:conv_tc_kernel" -c 46 -o gpurun_out/${TAG}_conv -f \
    python tools/profile_all.py convonly > gpurun_out/${TAG}_ncu_conv.log 2>&1
ncu -i gpurun_out/${TAG}_conv.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*; rm -f gpurun_out/${TAG}_sel.ncu-rep gpurun_out/${TAG}_conv.ncu-rep
du -sh gpurun_out
