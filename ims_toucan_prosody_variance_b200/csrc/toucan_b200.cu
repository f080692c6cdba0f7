// toucan_b200.cu -- single translation unit of libtoucan_b200.so (keeps __constant__ data and
// inline helpers shared without relocatable device code).
#include "api.cu"
#include "conv1d_simt.cu"
#include "conv1d_umma.cu"
#include "ragged.cu"
#include "acoustic.cu"
