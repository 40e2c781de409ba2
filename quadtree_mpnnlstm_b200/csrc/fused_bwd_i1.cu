#include "fused_bwd.inl"
namespace qmp {
template int launch_bwd<0, 36>(const FusedBwdArgs&, int, cudaStream_t);
}
