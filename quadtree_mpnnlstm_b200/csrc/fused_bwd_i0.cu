#include "fused_bwd.inl"
namespace qmp {
template int launch_bwd<0, 32>(const FusedBwdArgs&, int, cudaStream_t);
}
