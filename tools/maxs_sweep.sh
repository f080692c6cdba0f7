for shape in "128 128 11 1 0 24000 64 2" "128 128 7 1 0 24000 64 2" "128 128 3 1 0 24000 64 2" "256 256 11 1 0 4000 64 2" "256 256 7 1 0 4000 64 2" "256 256 3 1 0 4000 64 2"; do
  for ms in 8 2 1; do
    echo -n "MAX_S=$ms  "; TB200_MAX_S=$ms TB200_PLAN_DEBUG=1 python tools/conv_micro.py $shape f16 3 2>&1 | sort -u | tr '\n' ' ' | sed 's/tb200 plan: Cin=[0-9]* Cout=[0-9]* taps=[0-9]* up=0 act=2 L=[0-9]* ->//' | cut -c1-230; echo
  done
done
