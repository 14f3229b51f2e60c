set -x
SWEEP_L=48,72,96,120,144,192 timeout 300 python tools/cl_sweep.py > gpurun_out/cl_sweep3.log 2>&1
for L in 96 192; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:afa_cl_fwd -s 2 -c 2 -o gpurun_out/cl_L$L -f python tools/cl_ncu_case.py 384 13776 8 $L > gpurun_out/ncu_cl_L$L.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
