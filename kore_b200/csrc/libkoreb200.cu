// Single translation unit of libkoreb200.so (kernels are launched across the
// files below, so they are compiled together instead of with -rdc).
#include "kb_layout.cu"
#include "kb_setup.cu"
#include "kb_assemble.cu"
#include "kb_chainfac.cu"
#include "kb_factor.cu"
#include "kb_sweep.cu"
#include "kb_sweep1.cu"
#include "kb_sweep2.cu"
#include "kb_solve.cu"
#include "kb_eigs.cu"
#include "kb_shard.cu"
#include "kb_io.cu"
