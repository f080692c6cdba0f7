tag=$1
M2="python tools/conv_micro.py 64 64 3 1 0 96000 64 2 f16 3"
$M2 > gpurun_out/${tag}_micro2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o gpurun_out/${tag}_snake_c64k3 $M2 > gpurun_out/${tag}_ncu2.log 2>&1
cat gpurun_out/${tag}_micro2.log
