// toucan_b200.cu -- the translation unit of libtoucan_b200.so that owns the __constant__ data (the per-layer conv
// kernels and the C ABI).  acoustic.cu and respair_umma.cu are compiled separately (build.sh) with
// TB200_NO_AA_CONSTANT and linked into the same library; no relocatable device code is needed.
#include "api.cu"
#include "conv1d_simt.cu"
#include "conv1d_umma.cu"
#include "ragged.cu"
