// comm.h — the exchange layer of "one sample over N GPUs" (shard.cu).
//
// Reads shard naturally; what is global in the reference — the SAM order / QNAME dedup (catt.jl:45-48), the V∩J join
// (catt.jl:237, 288-319), the last-writer-wins k-mer table (catt.jl:426-497, 544-548) and the final lists — needs a handful of
// collectives.  Two transports behind one interface:
//   * NcclComm  — NCCL over NVLink 5 / NVSwitch (ncclCommInitAll for the GPUs of one process, ncclCommInitRank with a passed-in
//                 unique id for one process per GPU).  libnccl.so.2 is dlopen'ed, so a single-GPU deployment needs no NCCL.
//   * LocalComm — ranks are threads of one process that may share a device: peer copies + a reduction kernel that reads the
//                 peers' buffers directly.  It exists so the whole multi-rank path can be tested on ONE GPU (CATT_COMM=local
//                 forces it on several GPUs as well: direct P2P loads over NVLink).
// Buffers are device memory, sizes are bytes unless the name says otherwise; operations are ordered on the caller's stream.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <memory>
#include <vector>

namespace catt {

struct Comm {
  int rank = 0, nranks = 1;
  virtual ~Comm() {}
  virtual const char* kind() const = 0;
  // recv[displs[r] .. displs[r] + counts[r]) = rank r's `send` (counts[rank] bytes); counts / displs are the same on every rank
  virtual int allgatherv(const void* send, void* recv, const size_t* counts, const size_t* displs, cudaStream_t st) = 0;
  // in place on every rank; each element is written by at most one rank (the others hold 0) or is a plain sum
  virtual int allreduce_sum_u32(uint32_t* buf, size_t count, cudaStream_t st) = 0;
  virtual int reduce_sum_u32(uint32_t* buf, size_t count, int root, cudaStream_t st) = 0;
  // send[sdispls[r] .. + scounts[r]) goes to rank r, which finds it at recv[rdispls[me] ..) of ITS arrays
  virtual int alltoallv(const void* send, const size_t* scounts, const size_t* sdispls, void* recv, const size_t* rcounts,
                        const size_t* rdispls, cudaStream_t st) = 0;
  // small host values of every rank (blocking; the stream is drained first): out[r * bytes ..) = rank r's `in`
  virtual int host_allgather(const void* in, void* out, size_t bytes, cudaStream_t st) = 0;
  // bytes this rank put on / took off the wire so far, and the device time spent inside collectives (events on the stream)
  int64_t bytes_moved = 0;
};

// ranks of one process ------------------------------------------------------------------------------------------------------------
struct LocalGroup;                                                   // shared by the LocalComm objects of the process
std::shared_ptr<LocalGroup> local_group_create(int nranks, const int* devices);
Comm* local_comm_create(const std::shared_ptr<LocalGroup>& g, int rank);
void local_group_fail(const std::shared_ptr<LocalGroup>& g);        // a rank gives up: its peers stop waiting for it
void local_group_reset(const std::shared_ptr<LocalGroup>& g);       // before the next run

// NCCL ------------------------------------------------------------------------------------------------------------------------------
bool nccl_available();                                               // libnccl.so.2 could be loaded
int nccl_unique_id(unsigned char out[128]);
// one communicator per device of this process (ncclCommInitAll); out[r] drives devices[r]
int nccl_comms_create_all(int nranks, const int* devices, std::vector<Comm*>* out);
// this process is rank `rank` of `nranks` (one process per GPU; the id comes from rank 0 through the launcher's own channel)
int nccl_comm_create_rank(int nranks, int rank, const unsigned char id[128], Comm** out);

}  // namespace catt
