#!/bin/bash
# Regenerates the ncu evidence bench.py reads (run on the GPU box, from the repo root):
#   gpurun_out/<tag>_launches.csv   every launch of one bench step with its device time
#   gpurun_out/<tag>_full.ncu-rep   ncu --set full of every kernel of that step (one launch each)
# then, here or there:  python profiles/summarise_profiles.py <tag>   ->  profiles/<tag>_kernels.json + per-kernel .txt
# Usage: bash profiles/make_profiles.sh r02 [batch]
set -e
TAG=${1:-r02}
BATCH=${2:-512}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches.csv python profiles/profile_step.py --batch $BATCH > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:'^k_' \
    -f -o gpurun_out/${TAG}_full python profiles/profile_step.py --batch $BATCH > gpurun_out/${TAG}_full.log 2>&1
# summarise on the box (the full report is too large to travel back), keep the flux kernels' report for the source page
python profiles/summarise_profiles.py ${TAG} ${BATCH} gpurun_out > gpurun_out/${TAG}_summary.txt 2>&1 || true
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'^k_azinv_flux' \
    -f -o gpurun_out/${TAG}_flux python profiles/profile_step.py --batch 64 > gpurun_out/${TAG}_flux.log 2>&1 || true
rm -f gpurun_out/${TAG}_full.ncu-rep
ls -la gpurun_out/${TAG}_*
