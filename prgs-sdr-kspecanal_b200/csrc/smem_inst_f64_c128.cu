// fused scan kernels: f64 arithmetic, c128 ingest (see curscan_smem.cuh)
#define KSPEC_INST_T double
#define KSPEC_INST_FMT KSPEC_IN_C128
#define KSPEC_INST_NAME launch_smem_f64_c128
#define KSPEC_INST_MAXLOG2F 13
#include "smem_inst.cuh"
