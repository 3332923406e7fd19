// kvquant_b200 -- C-ABI bookkeeping entry points (version, error strings, launch counter).
#include "kvq_common.cuh"

namespace kvq {
unsigned long long g_launch_count = 0;
}

extern "C" {

int kvq_abi_version(void) { return KVQ_ABI_VERSION; }

uint64_t kvq_launch_count(void) { return kvq::g_launch_count; }

const char* kvq_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case KVQ_E_BITS: return "kvquant_b200: bits must be 2, 3 or 4";
    case KVQ_E_SHAPE: return "kvquant_b200: inconsistent sizes";
    case KVQ_E_NULL: return "kvquant_b200: required pointer is NULL";
    case KVQ_E_ALIGN: return "kvquant_b200: alignment requirement violated (cache 16 B, Lmax % 4 == 0, LUT 16 B)";
    case KVQ_E_UNSUPPORTED: return "kvquant_b200: unsupported configuration";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "kvquant_b200: unknown error";
}

}  // extern "C"
