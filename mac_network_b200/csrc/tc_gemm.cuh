// bf16 tcgen05/TMEM read-unit chain (MAC_PREC_BF16).  Placeholder until the tensor-core path lands.
#pragma once
#include "common.cuh"
namespace mac {
inline size_t tc_read_extra_workspace_bytes(int, int, int) { return 0; }
inline int tc_read_chain(const void*, const float*, const float*, const mac_read_weights*, uint32_t, float, uint64_t,
                         int, float*, float*, float*, float*, int*, void*, size_t, int, int, int, bool, cudaStream_t) {
  return MAC_ERR_UNSUPPORTED;
}
}  // namespace mac
