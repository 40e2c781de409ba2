#include "fused_bwd.inl"
namespace qmp {
template int launch_bwd<8, 32>(const FusedBwdArgs&, int, cudaStream_t);
}
