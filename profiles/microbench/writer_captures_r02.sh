# ncu --set full captures of the round-2 row writers outside the f64 packed class (run on the GPU box from the repo root):
# the any-plan writer (cz_obs_any_kernel, forced on the headline tables with CZ_GENERIC=2) and the two-environments-per-warp
# float32 writer (cz_obs32_pair_kernel), 131072 environments each, through profiles/microbench/obs_time.py.
set -x
CZ_GENERIC=2 ncu --set full --clock-control none --import-source on -k regex:cz_obs_any_kernel -s 10 -c 1 -o gpurun_out/r02g_obs_any python profiles/microbench/obs_time.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cz_obs32_pair_kernel -s 10 -c 1 -o gpurun_out/r02g_obs32_pair python profiles/microbench/obs_time.py > /dev/null 2>&1
CZ_OBS32_PAIR=0 ncu --set full --clock-control none --import-source on -k regex:cz_obs32_fast_kernel -s 10 -c 1 -o gpurun_out/r02g_obs32_one python profiles/microbench/obs_time.py > /dev/null 2>&1
ls -la gpurun_out/r02g_*
