// comm.h -- thin NCCL binding (dlopen'ed, so single-GPU use has no NCCL dependency).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace s3o {

struct Comm {
    void *nccl = nullptr;   // ncclComm_t
    int rank = 0, world = 1;
};

constexpr int kUniqueIdBytes = 128;

int comm_load();                                            // 0 ok
int comm_unique_id(char out[kUniqueIdBytes]);
int comm_init(Comm &c, int rank, int world, const char id[kUniqueIdBytes]);
void comm_destroy(Comm &c);
int comm_allreduce_sum(Comm &c, double *buf, size_t count, cudaStream_t st);   // in place
int comm_allreduce_max(Comm &c, double *buf, size_t count, cudaStream_t st);
int comm_allgather(Comm &c, const double *send, double *recv, size_t count_per_rank, cudaStream_t st);
int comm_allgatherv(Comm &c, double *buf, const size_t *off, const size_t *count, cudaStream_t st);   // in place
int comm_allgather_bytes(Comm &c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st);
// exchange: send send_count[q] doubles from sendbuf+send_off[q] to q, receive recv_count[q] into recvbuf+recv_off[q]
int comm_halo(Comm &c, const double *sendbuf, const int *send_off, const int *send_count, double *recvbuf,
              const int *recv_off, const int *recv_count, int unit, cudaStream_t st);
const char *comm_last_error();

}  // namespace s3o
