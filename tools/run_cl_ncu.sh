# ncu evidence for the channels-last engine and kernels (run under gpurun, one GPU; outputs in gpurun_out/)
set -x
python tools/engine_pass.py 4 > gpurun_out/engine_pass.log 2>&1 || exit 1
# launch list of ONE warm eager pass (4 binaural clips, bf16): every kernel with its device time
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_cl_launches_engine_bf16_clips4.csv python tools/engine_pass.py 4 > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r01_cl_launches_engine_bf16_clips4.csv 1 > gpurun_out/r01_cl_launches_engine_bf16_clips4.txt
# full captures: channels-last activation (plain + residual variant) and the fused activation -> convolution kernel
python tools/cl_ncu_case.py 384 13776 8 0 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:afa_cl_fwd -s 2 -c 2 -o gpurun_out/r01_cl_fwd_bf16_B8_C384_T13776 -f python tools/cl_ncu_case.py 384 13776 8 0 > gpurun_out/ncu_cl_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r01_cl_fwd_bf16_B8_C384_T13776.ncu-rep > gpurun_out/r01_cl_ncu_full_fwd_bf16_B8_C384_T13776.txt
python tools/actconv_ncu_case.py 24 220416 8 7 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:afa_cl_actconv -s 2 -c 2 -o gpurun_out/r01_actconv_bf16_B8_C24_T220416_k7d3 -f python tools/actconv_ncu_case.py 24 220416 8 7 3 > gpurun_out/ncu_actconv_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r01_actconv_bf16_B8_C24_T220416_k7d3.ncu-rep > gpurun_out/r01_actconv_ncu_full_bf16_B8_C24_T220416_k7d3.txt
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_*.txt | tail -8
