// Tensor-core (tcgen05, error-compensated 3xTF32) implicit-GEMM convolution -- see DESIGN.md.
// PLACEHOLDER until the tcgen05 kernel lands: reports "unsupported" so that every conv runs on the fp32 CUDA-core kernel.
#include "net.cuh"

namespace b200trk {
struct TcConv { int unused; };
bool tc_conv_supported(const Op&) { return false; }
int tc_conv_prepare(b200trk_net*, Op&, const std::vector<float>&) { return 0; }
int tc_conv_launch(b200trk_net*, const Op&, int, cudaStream_t) { set_error("tc conv not built"); return 9; }
void tc_conv_free(TcConv* tc) { delete tc; }
}  // namespace b200trk
