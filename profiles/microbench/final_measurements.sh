# Round-1 final measurements (run on the GPU box from the repo root): bench arms, launch list, ncu captures.
set -x
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference > gpurun_out/final_ref.json 2>> gpurun_out/final_bench.err
python bench.py --mode sync --no-cpu > gpurun_out/final_sync.json 2>> gpurun_out/final_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 30 --warmup 5 --no-cpu --no-cfg3 --e2e-steps 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cz_obs_envs_kernel -s 20 -c 1 -o gpurun_out/final_obs python bench.py --steps 10 --warmup 5 --no-cpu --no-cfg3 --e2e-steps 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cz_env_kernel -s 24 -c 1 -o gpurun_out/final_dyn python bench.py --steps 10 --warmup 5 --no-cpu --no-cfg3 --e2e-steps 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cz_env_kernel -s 10 -c 1 -o gpurun_out/final_fused python bench.py --envs 32768 --steps 10 --warmup 5 --no-cpu --no-cfg3 --e2e-steps 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cz_policy_kernel -s 50 -c 1 -o gpurun_out/final_pol python profiles/microbench/policy_loop.py > /dev/null 2>&1
ls -la gpurun_out/final_*
