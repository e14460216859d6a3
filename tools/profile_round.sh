#!/usr/bin/env bash
# Round evidence (GPU box, one GPU), reduced on the box to small text files under gpurun_out/<tag>_*:
#   1. launch list (device time + DRAM bytes per launch) of ONE forward of the benchmarked path at the bench batch
#   2. ncu --set full of the BASELINE config 2 micro-bench kernels (NI-LIF, spike GEMM fc1, SDSA)
#   3. ncu --set full of the first launch of each of the 18 most expensive (kernel, layer shape) groups of that forward
#   tools/profile_round.sh <tag> [bench batch]
set -uo pipefail
tag="$1"; B="${2:-64}"
mkdir -p gpurun_out
if [[ "${PROFILE_STEPS:-123}" == *1* ]]; then
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file "gpurun_out/${tag}_launches.csv" python tools/profile_forward.py ade20k "$B" logits > "gpurun_out/${tag}_launches.log" 2>&1
python tools/ncu_summary.py launches "gpurun_out/${tag}_launches.csv" "gpurun_out/${tag}_launches_b${B}.md" "" "$B" > /dev/null
fi
if [[ "${PROFILE_STEPS:-123}" == *2* ]]; then
ncu --profile-from-start off --set full --clock-control none -f -o "/tmp/${tag}_micro" python tools/profile_forward.py ade20k 64 logits micro > "gpurun_out/${tag}_micro.log" 2>&1
python tools/ncu_summary.py full "/tmp/${tag}_micro.ncu-rep" "gpurun_out/${tag}_micro_cfg2_full.md" > /dev/null
fi
if [[ "${PROFILE_STEPS:-123}" == *3* ]]; then
timeout 600 ncu --profile-from-start off --set full --clock-control none -f -o "/tmp/${tag}_top" python tools/profile_forward.py ade20k "$B" logits top18 > "gpurun_out/${tag}_top.log" 2>&1
python tools/ncu_summary.py full "/tmp/${tag}_top.ncu-rep" "gpurun_out/${tag}_top18_b${B}_full.md" > /dev/null
fi
ls -la gpurun_out/${tag}_*
