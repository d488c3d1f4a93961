# Round-2 final measurements (run on the GPU box from the repo root): bench arms, launch lists, ncu captures.
set -x
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_steps20.json 2>> gpurun_out/r02_final_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference_arm.json 2>> gpurun_out/r02_final_bench.err
# launch lists (per-launch times under ncu are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_final_launches_sync.csv python bench.py --mode sync --steps 30 --warmup 5 --no-cpu --no-cfg3 --no-bg --e2e-steps 3 --no-graph > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_final_launches_pipelined.csv python bench.py --mode pipelined --steps 30 --warmup 5 --no-cpu --no-cfg3 --no-bg --e2e-steps 3 --no-graph > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_final_launches_cfg3.csv python profiles/microbench/warp_loop.py 4096 64 > /dev/null 2>&1
# full captures of the top kernels
ncu --set full --clock-control none --import-source on -k regex:cz_obs_whole_kernel -s 20 -c 1 -o gpurun_out/r02_final_obs python bench.py --mode sync --steps 10 --warmup 5 --no-cpu --no-cfg3 --no-bg --e2e-steps 3 --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cz_env_kernel -s 24 -c 1 -o gpurun_out/r02_final_dyn python bench.py --mode sync --steps 10 --warmup 5 --no-cpu --no-cfg3 --no-bg --e2e-steps 3 --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cz_warp_kernel -s 3 -c 1 -o gpurun_out/r02_final_warp python profiles/microbench/warp_loop.py 4096 64 > /dev/null 2>&1
ls -la gpurun_out/r02_final_*
