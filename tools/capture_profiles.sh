#!/usr/bin/env bash
# One gpurun call worth of evidence for profiles/ (run ON the GPU box, from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r2_v1'
# Writes into gpurun_out/ (merged back by gpurun): bench lines of both arms, the ncu launch list of bench.py (match / RANSAC
# kernels only: node creation launches ~1500 expand kernels first), and one `ncu --set full` report of the three heavy kernels.
# Summarise here afterwards with  python profiles/extract.py gpurun_out/<tag>_kernels.ncu-rep > profiles/<tag>_kernels_ncu_full.txt
set -u
tag="${1:-rX}"
out=gpurun_out
mkdir -p "$out"
python bench.py > "$out/${tag}_bench.json" 2> "$out/${tag}_bench.err"; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > "$out/${tag}_bench_reference_arm.json" 2>/dev/null; echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"select|ransac|tc_match|refine|emm" --launch-skip 60 -c 160 \
    --csv --log-file "$out/${tag}_launches.csv" python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"ransac_hyp|tc_match256" --launch-skip 6 --launch-count 3 \
    -o "$out/${tag}_kernels" python tools/run_batch.py 3 2>&1 | tail -1
ls -la "$out"
