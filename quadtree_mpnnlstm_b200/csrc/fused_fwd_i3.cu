#include "fused_fwd.inl"
namespace qmp {
template int launch_fwd<8, 32>(const FusedFwdArgs&, cudaStream_t);
}
