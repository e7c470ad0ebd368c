# Round-end evidence on one B200 (run through gpurun): bench lines, ncu launch lists, `ncu --set full` summaries per kernel family.
# The .ncu-rep files are summarised on the box (profiles/extract.py) and deleted, gpurun_out/ must stay below 64 MiB;
# only the C2 report (match + RANSAC kernels, with source) is kept for source-level analysis.
set -x
mkdir -p gpurun_out
P=gpurun_out/r2_final
python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; tail -c 200 ${P}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference_arm.json 2>/dev/null
timeout 300 python tools/run_families.py > ${P}_families.json 2> ${P}_families.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_families_launches.csv python tools/run_families.py orb sift posegraph emm refine > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file ${P}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c3 --c4-frames 128 > ${P}_bench_under_ncu.json 2>/dev/null
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"tc_hamming_expand|ransac_hyp|select_matches|ransac_select" --launch-skip 10 --launch-count 5 -o ${P}_c2 python tools/run_batch.py 4 2>&1 | tail -1
python profiles/extract.py ${P}_c2.ncu-rep > ${P}_c2_kernels_ncu_full.txt
RB200_ORB_FRAMES=16 RB200_ORB_REPS=2 timeout 300 ncu --set full --clock-control none -k regex:"k_fast_nms|k_describe|k_resize_cells|k_blur|k_cell_extract|k_frame_emit|k_frame_finalize|k_adapt|k_harris|k_cell_select" --launch-skip 40 --launch-count 16 -o ${P}_orb python tools/run_families.py orb 2>&1 | tail -1
python profiles/extract.py ${P}_orb.ncu-rep > ${P}_orb_kernels_ncu_full.txt; rm -f ${P}_orb.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:"tc_match256|k_l2_refine|k_select_sift|k_sift_prepare" --launch-skip 6 --launch-count 8 -o ${P}_sift python tools/run_families.py sift 2>&1 | tail -1
python profiles/extract.py ${P}_sift.ncu-rep > ${P}_sift_kernels_ncu_full.txt; rm -f ${P}_sift.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:"pg_pcg_resident|pg_linearize|pg_assemble|pg_orient|pg_chi2|pg_update|pg_precond" --launch-skip 7 --launch-count 7 -o ${P}_posegraph python tools/run_families.py posegraph 2>&1 | tail -1
python profiles/extract.py ${P}_posegraph.ncu-rep > ${P}_posegraph_kernels_ncu_full.txt; rm -f ${P}_posegraph.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:"ba_" --launch-skip 12 --launch-count 10 -o ${P}_landmark python tools/run_families.py landmark 2>&1 | tail -1
python profiles/extract.py ${P}_landmark.ncu-rep > ${P}_landmark_kernels_ncu_full.txt; rm -f ${P}_landmark.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:"k_emm_pairs|refine_g2o" --launch-skip 1 --launch-count 3 -o ${P}_emm_refine python tools/run_families.py emm refine 2>&1 | tail -1
python profiles/extract.py ${P}_emm_refine.ncu-rep > ${P}_emm_refine_kernels_ncu_full.txt; rm -f ${P}_emm_refine.ncu-rep
ls -la gpurun_out; du -sh gpurun_out
