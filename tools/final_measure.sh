# One-GPU measurement set of a round (run under gpurun from the repo root): driver-like bench line, long run, CPU arm,
# secondary workloads, ncu launch list + full capture of the hot kernels, sanitizer runs of the new kernels.
# Outputs under gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
TAG=${1:-r2_final}
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_n1_200.json 2>> gpurun_out/${TAG}_n1.err
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-secondary --corr 0 > gpurun_out/${TAG}_n1_200_nocorr.json 2>> gpurun_out/${TAG}_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2>> gpurun_out/${TAG}_n1.err
python bench.py --workload stage1 --steps 100 --warmup 5 > gpurun_out/${TAG}_stage1.json 2>> gpurun_out/${TAG}_n1.err
python bench.py --workload dino --steps 20 --warmup 3 > gpurun_out/${TAG}_dino.json 2>> gpurun_out/${TAG}_n1.err
python bench.py --workload preprocess --steps 20 --warmup 3 > gpurun_out/${TAG}_pre.json 2>> gpurun_out/${TAG}_n1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --mesh uv100x200 --camera 1080x1920 --frames-per-gpu 125 > gpurun_out/${TAG}_c3.json 2>> gpurun_out/${TAG}_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(raster|backward|neg_maps|setup_bin|project|pose_prep|pose_update|finalize|corr)" -s 30 -c 100 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_raster|k_backward|k_neg_maps|k_setup_bin|k_corr|k_project|k_pose|k_finalize" -s 30 -c 10 -o gpurun_out/prof_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_prior_features.py -q -x -k "reference_run or oracle" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck.log
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_sharding.py -q -x -k "loopback or emulated_shards_equal_single" > gpurun_out/${TAG}_memcheck2.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck2.log
compute-sanitizer --tool racecheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${TAG}_racecheck.log
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_jointopt.py -q -x -k "list_path_and_bitmap or partly_outside or joint_optimize_matches" > gpurun_out/${TAG}_memcheck3.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck3.log
for f in memcheck memcheck2 memcheck3 racecheck; do echo "== $f"; tail -n 2 gpurun_out/${TAG}_$f.log; done
