#include "fused_bwd_tc.inl"
namespace qmp {
template int launch_bwd_tc<0, 32, 1>(const FusedBwdArgs&, cudaStream_t);
}
