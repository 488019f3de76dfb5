// placeholder until the fused kernels land
#include "msda_host.h"
namespace msda {
int launch_sca_fwd(const FusedProblem&, cudaStream_t) { return set_error(MSDA_ERR_UNSUPPORTED, "sca_fwd: not built"); }
int launch_sca_bwd(const FusedProblem&, cudaStream_t) { return set_error(MSDA_ERR_UNSUPPORTED, "sca_bwd: not built"); }
int launch_tsa_fwd(const FusedProblem&, cudaStream_t) { return set_error(MSDA_ERR_UNSUPPORTED, "tsa_fwd: not built"); }
int launch_tsa_bwd(const FusedProblem&, cudaStream_t) { return set_error(MSDA_ERR_UNSUPPORTED, "tsa_bwd: not built"); }
}
