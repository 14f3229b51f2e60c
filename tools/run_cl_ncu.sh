# ncu evidence for the channels-last kernels (run under gpurun, one GPU)
set -x
python tools/engine_pass.py 4 > gpurun_out/engine_pass.log 2>&1 || exit 1
# launch list of the LAST of 3 eager passes (cuDNN autotune happens in the first): ~330 launches per pass
PASSES=3 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_cl_launches_engine_bf16_clips4.csv python tools/engine_pass.py 4 > gpurun_out/ncu_launches.log 2>&1
python tools/cl_ncu_case.py 384 13776 8 0 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:afa_cl_fwd -s 2 -c 2 -o gpurun_out/r01_cl_fwd_bf16_B8_C384_T13776 -f python tools/cl_ncu_case.py 384 13776 8 0 > gpurun_out/ncu_cl_full.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv | tail -5
