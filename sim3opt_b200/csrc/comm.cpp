// comm.cpp -- NCCL over NVLink/NVSwitch for the partitioned solve.
// One process per GPU; the communicator is created from a unique id that the host program
// (bench.py / tests, via torch.distributed) broadcasts.  Collectives used on the data path:
//   all-reduce of 1-2 fp64 scalars (PCG dot products, chi2, computeScale, max diagonal),
//   grouped send/recv of the halo entries of p before every SpMV (fallback: normally the SpMV loads them from
//   the peers' memory over NVLink, see setup_p2p in problem.cu; the IPC handles travel by an all-gather here),
//   all-gather of the step x before the retraction,
//   in-place all-gather of unequal segments (grouped broadcasts) for the multilevel preconditioner.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include "comm.h"

namespace s3o {

namespace {
struct Api {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g;
char g_comm_err[256] = "";

int fail(const char *what, ncclResult_t r) {
    snprintf(g_comm_err, sizeof g_comm_err, "%s: %s", what, g.GetErrorString ? g.GetErrorString(r) : "nccl error");
    return -1;
}
}  // namespace

const char *comm_last_error() { return g_comm_err; }

int comm_load() {
    if (g.handle) return 0;
    static_assert(sizeof(ncclUniqueId) == kUniqueIdBytes, "ncclUniqueId size");
    // RTLD_NOLOAD first: reuse the copy torch already mapped (same SONAME), else load the system one
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) { snprintf(g_comm_err, sizeof g_comm_err, "dlopen libnccl.so.2: %s", dlerror()); return -1; }
#define S3O_SYM(field, name)                                                         \
    g.field = reinterpret_cast<decltype(g.field)>(dlsym(h, name));                   \
    if (!g.field) { snprintf(g_comm_err, sizeof g_comm_err, "dlsym %s failed", name); return -1; }
    S3O_SYM(GetUniqueId, "ncclGetUniqueId")
    S3O_SYM(CommInitRank, "ncclCommInitRank")
    S3O_SYM(CommDestroy, "ncclCommDestroy")
    S3O_SYM(AllReduce, "ncclAllReduce")
    S3O_SYM(AllGather, "ncclAllGather")
    S3O_SYM(Broadcast, "ncclBroadcast")
    S3O_SYM(Send, "ncclSend")
    S3O_SYM(Recv, "ncclRecv")
    S3O_SYM(GroupStart, "ncclGroupStart")
    S3O_SYM(GroupEnd, "ncclGroupEnd")
    S3O_SYM(GetErrorString, "ncclGetErrorString")
#undef S3O_SYM
    g.handle = h;
    return 0;
}

int comm_unique_id(char out[kUniqueIdBytes]) {
    if (comm_load()) return -1;
    ncclUniqueId id;
    ncclResult_t r = g.GetUniqueId(&id);
    if (r != ncclSuccess) return fail("ncclGetUniqueId", r);
    memcpy(out, &id, kUniqueIdBytes);
    return 0;
}

int comm_init(Comm &c, int rank, int world, const char id_bytes[kUniqueIdBytes]) {
    if (comm_load()) return -1;
    ncclUniqueId id;
    memcpy(&id, id_bytes, kUniqueIdBytes);
    ncclComm_t comm;
    ncclResult_t r = g.CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) return fail("ncclCommInitRank", r);
    c.nccl = comm;
    c.rank = rank;
    c.world = world;
    return 0;
}

void comm_destroy(Comm &c) {
    if (c.nccl && g.CommDestroy) g.CommDestroy((ncclComm_t)c.nccl);
    c.nccl = nullptr;
}

int comm_allreduce_sum(Comm &c, double *buf, size_t count, cudaStream_t st) {
    ncclResult_t r = g.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)c.nccl, st);
    return r == ncclSuccess ? 0 : fail("ncclAllReduce(sum)", r);
}

int comm_allreduce_max(Comm &c, double *buf, size_t count, cudaStream_t st) {
    ncclResult_t r = g.AllReduce(buf, buf, count, ncclDouble, ncclMax, (ncclComm_t)c.nccl, st);
    return r == ncclSuccess ? 0 : fail("ncclAllReduce(max)", r);
}

int comm_allgather(Comm &c, const double *send, double *recv, size_t count_per_rank, cudaStream_t st) {
    ncclResult_t r = g.AllGather(send, recv, count_per_rank, ncclDouble, (ncclComm_t)c.nccl, st);
    return r == ncclSuccess ? 0 : fail("ncclAllGather", r);
}

// in-place all-gather of unequal segments: rank q owns buf[off[q] .. off[q]+count[q])
int comm_allgatherv(Comm &c, double *buf, const size_t *off, const size_t *count, cudaStream_t st) {
    ncclResult_t r = g.GroupStart();
    if (r != ncclSuccess) return fail("ncclGroupStart", r);
    for (int q = 0; q < c.world; ++q) {
        if (count[q] == 0) continue;
        r = g.Broadcast(buf + off[q], buf + off[q], count[q], ncclDouble, q, (ncclComm_t)c.nccl, st);
        if (r != ncclSuccess) { g.GroupEnd(); return fail("ncclBroadcast", r); }
    }
    r = g.GroupEnd();
    return r == ncclSuccess ? 0 : fail("ncclGroupEnd", r);
}

// all-gather of opaque bytes (IPC handles of the peer-to-peer halo), device buffers
int comm_allgather_bytes(Comm &c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st) {
    ncclResult_t r = g.AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t)c.nccl, st);
    return r == ncclSuccess ? 0 : fail("ncclAllGather(bytes)", r);
}

int comm_halo(Comm &c, const double *sendbuf, const int *send_off, const int *send_count, double *recvbuf,
              const int *recv_off, const int *recv_count, int unit, cudaStream_t st) {
    ncclResult_t r = g.GroupStart();
    if (r != ncclSuccess) return fail("ncclGroupStart", r);
    for (int q = 0; q < c.world; ++q) {
        if (q == c.rank) continue;
        if (send_count[q] > 0) {
            r = g.Send(sendbuf + (size_t)send_off[q] * unit, (size_t)send_count[q] * unit, ncclDouble, q, (ncclComm_t)c.nccl, st);
            if (r != ncclSuccess) { g.GroupEnd(); return fail("ncclSend", r); }
        }
        if (recv_count[q] > 0) {
            r = g.Recv(recvbuf + (size_t)recv_off[q] * unit, (size_t)recv_count[q] * unit, ncclDouble, q, (ncclComm_t)c.nccl, st);
            if (r != ncclSuccess) { g.GroupEnd(); return fail("ncclRecv", r); }
        }
    }
    r = g.GroupEnd();
    return r == ncclSuccess ? 0 : fail("ncclGroupEnd", r);
}

}  // namespace s3o
