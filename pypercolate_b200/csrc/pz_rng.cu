// pz_rng.cu -- bond orders generated on the device (placeholder until the
// Philox / MT19937 kernels land).
#include "pz_common.cuh"
#include "pz_internal.h"
namespace pz {
cudaError_t launch_perm_philox(int32_t, int32_t, const uint32_t *, int32_t *, cudaStream_t, int *)
{ return cudaErrorNotSupported; }
cudaError_t launch_perm_mt19937(int32_t, int32_t, const uint32_t *, int32_t *, cudaStream_t, int *)
{ return cudaErrorNotSupported; }
}
