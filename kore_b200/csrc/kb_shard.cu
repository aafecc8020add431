// l-sharded (multi-GPU) factor / solve over NCCL.  See DESIGN.md "Multi-GPU".
#include "kb_internal.cuh"

void kbi_nccl_destroy(kb_context* h) { (void)h; }

int kbi_factor_sharded(kb_context* h, zcomplex) {
  return kb_fail(h, KB_EINVAL, "l-sharded factorisation is not available in this build");
}
int kbi_chain_solve_sharded(kb_context* h, const double2*, double2*, int) {
  return kb_fail(h, KB_EINVAL, "l-sharded solve is not available in this build");
}
extern "C" int kb_nccl_unique_id(void* id128) {
  (void)id128;
  return KB_ENCCL;
}
extern "C" int kb_set_sharding(kb_handle h, int rank, int nranks, const void* id) {
  (void)id;
  if (!h) return KB_EINVAL;
  if (nranks == 1 && rank == 0) return KB_OK;
  return kb_fail(h, KB_ENCCL, "l-sharding is not available in this build");
}
