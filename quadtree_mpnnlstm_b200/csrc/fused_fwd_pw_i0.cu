#include "fused_fwd_pw.inl"
namespace qmp {
template int launch_fwd_pw<0>(const FusedFwdArgs&, cudaStream_t);
}
