// nccl_shard.cu -- multi-process j-sharding over NCCL (one process per GPU).  Placeholder entry
// points until the combine kernel lands; they fail loudly rather than silently doing nothing.
#include <cstdio>
#include <cstdlib>
#include "../../include/gpunb_b200.h"
extern "C" {
int gpunb_b200_nccl_unique_id(unsigned char id128[128]) { (void)id128; return -1; }
int gpunb_b200_nccl_init(int rank, int nranks, const unsigned char id128[128]) { (void)rank; (void)nranks; (void)id128; return -1; }
void gpunb_b200_nccl_finalize(void) {}
}
