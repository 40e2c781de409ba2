#include "fused_bwd_tc.inl"
namespace qmp {
template int launch_bwd_tc<4, 32, 1>(const FusedBwdArgs&, cudaStream_t);
}
