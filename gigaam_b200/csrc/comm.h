// comm.cu: run-time-bound NCCL for the hypothesis all-gather (gam_comm_* / gam_gather_hyps in the public header)
#pragma once
#include <cuda_runtime.h>

namespace gam {
const char* comm_unavailable_reason();   // nullptr when NCCL is bound
int comm_unique_id(unsigned char* out128);
int comm_init(void** comm, const unsigned char* id128, int rank, int nranks, const char** err);
int comm_all_gather_i32(void* comm, const int* send, int* recv, long long count, cudaStream_t s, const char** err);
void comm_destroy(void* comm);
int comm_nccl_version();
}  // namespace gam
