TB200_TRACE=1 python tools/conv_micro.py 64 32 4 1 2 96000 64 0 f16 3 nores 2>&1 | sed -n '1p;2p;7,10p' | cut -c1-700
TB200_TRACE=1 python tools/conv_micro.py 128 64 8 1 4 24000 64 0 f16 3 nores 2>&1 | sed -n '1p;7,10p' | cut -c1-300
python tools/conv_micro.py 256 128 12 1 6 4000 64 0 f16 3 nores 2>&1 | head -1
python tools/conv_micro.py 512 256 16 1 8 500 64 0 f16 3 nores 2>&1 | head -1
