tag=${1:-rX}
M1="python tools/conv_micro.py 32 32 3 1 0 192000 64 1 f16 3"
$M1 > gpurun_out/${tag}_micro1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o gpurun_out/${tag}_leaky_c32k3 $M1 > gpurun_out/${tag}_ncu1.log 2>&1
M2="python tools/conv_micro.py 64 64 11 1 0 96000 64 2 f16 3"
$M2 > gpurun_out/${tag}_micro2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o gpurun_out/${tag}_snake_c64k11 $M2 > gpurun_out/${tag}_ncu2.log 2>&1
