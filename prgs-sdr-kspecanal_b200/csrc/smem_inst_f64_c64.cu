// fused scan kernels: f64 arithmetic, c64 ingest (see curscan_smem.cuh)
#define KSPEC_INST_T double
#define KSPEC_INST_FMT KSPEC_IN_C64
#define KSPEC_INST_NAME launch_smem_f64_c64
#define KSPEC_INST_MAXLOG2F 13
#ifdef KSPEC_TUNING
#define KSPEC_INST_VARIANTS 1
#endif
#include "smem_inst.cuh"
