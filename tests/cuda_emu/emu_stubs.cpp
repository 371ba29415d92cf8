// TEST INFRASTRUCTURE — the parts of the library that are not emulated (multi-GPU transports)
#include "nvc_internal.h"
namespace nvc
{
void nccl_destroy(NvcContext*) {}
void gather_destroy(NvcContext*) {}
uint32_t* gather_fused_target(NvcContext*) { return nullptr; }
uint32_t gather_reserved_blocks(NvcContext*) { return 0; }
} // namespace nvc
