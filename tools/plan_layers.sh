python tools/profile_vocoder.py bigvgan > gpurun_out/pl_new.log 2>/dev/null
TB200_SNAKE_PLAN_OLD=1 python tools/profile_vocoder.py bigvgan > gpurun_out/pl_old.log 2>/dev/null
paste <(cut -c1-45 gpurun_out/pl_old.log) <(cut -c33-45 gpurun_out/pl_new.log) | grep -E "^ 128  128|^  64   64  11|total"
