#include "fused_fwd.inl"
namespace qmp {
template int launch_fwd<0, 36>(const FusedFwdArgs&, cudaStream_t);
}
