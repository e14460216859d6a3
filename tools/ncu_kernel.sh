#!/usr/bin/env bash
# One `ncu --set full` capture of the first launch matching a kernel-name regex, reduced ON THE BOX to two small text files
# (details page + per-instruction source page) under gpurun_out/: the .ncu-rep itself is not copied back.
#   [NCU_SKIP=n] tools/ncu_kernel.sh <tag> <kernel regex> <command ...>      (NCU_SKIP: matching launches to skip first)
set -euo pipefail
tag="$1"; regex="$2"; shift 2
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k "regex:$regex" -s "${NCU_SKIP:-0}" -c 1 -f -o "/tmp/$tag" "$@" > "gpurun_out/${tag}_run.log" 2>&1 || true
ncu -i "/tmp/$tag.ncu-rep" --page details > "gpurun_out/${tag}_details.txt"
ncu -i "/tmp/$tag.ncu-rep" --page source --csv > "gpurun_out/${tag}_source.csv"
ncu -i "/tmp/$tag.ncu-rep" --page raw --csv > "gpurun_out/${tag}_raw.csv"
