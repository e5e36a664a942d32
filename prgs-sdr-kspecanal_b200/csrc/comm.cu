// comm.cu — placeholder for the NCCL stats all-reduce (next milestone).
#include "kspec_internal.h"
extern "C" {
int kspec_comm_unique_id(char*) { kspec::set_error("NCCL layer not built yet"); return KSPEC_ERR_UNSUPPORTED; }
int kspec_comm_init(kspec_comm**, int, int, const char*, int) { kspec::set_error("NCCL layer not built yet"); return KSPEC_ERR_UNSUPPORTED; }
int kspec_comm_allreduce_stats(kspec_comm*, double*, double*, double*, int64_t) { kspec::set_error("NCCL layer not built yet"); return KSPEC_ERR_UNSUPPORTED; }
int kspec_comm_finalize(kspec_comm*) { return KSPEC_OK; }
}
