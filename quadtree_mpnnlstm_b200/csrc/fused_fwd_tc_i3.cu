#include "fused_fwd_tc.inl"
namespace qmp {
template int launch_fwd_tc<8, 32>(const FusedFwdArgs&, cudaStream_t);
}
