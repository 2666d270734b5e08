// comm.cu — the two transports of comm.h: NCCL (dlopen'ed) and the in-process peer-copy transport used for single-GPU tests.
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <condition_variable>
#include <mutex>

#include "internal.cuh"

namespace catt {
namespace {

constexpr int MAX_RANKS = 16;

// ---------------------------------------------------------------------------------------------------------------------------------
// LocalComm: ranks are threads of this process
// ---------------------------------------------------------------------------------------------------------------------------------
struct PeerPtrs { const uint32_t* p[MAX_RANKS]; int n; };

__global__ void peer_sum_kernel(PeerPtrs pp, uint32_t* __restrict__ out, size_t count) {
  const size_t n4 = count / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    uint4 s = reinterpret_cast<const uint4*>(pp.p[0])[i];
    for (int r = 1; r < pp.n; r++) {
      const uint4 v = reinterpret_cast<const uint4*>(pp.p[r])[i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<uint4*>(out)[i] = s;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t s = 0;
    for (int r = 0; r < pp.n; r++) s += pp.p[r][i];
    out[i] = s;
  }
}

}  // namespace

struct LocalGroup {
  int n = 0;
  std::vector<int> dev;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t gen = 0;
  bool failed = false;
  std::vector<const void*> ptr;
  std::vector<const size_t*> cnt, dsp;
  std::vector<char> host;
  // returns false when a rank has given up (its error is the one reported): nobody waits for it
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (failed) return false;
    const uint64_t g = gen;
    if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); return true; }
    cv.wait(lk, [&] { return gen != g || failed; });
    return !failed;
  }
  void fail() {
    std::lock_guard<std::mutex> lk(mu);
    failed = true;
    cv.notify_all();
  }
};

namespace {

struct LocalComm : Comm {
  std::shared_ptr<LocalGroup> g;
  void* tmp = nullptr; size_t tmp_cap = 0;
  LocalComm(const std::shared_ptr<LocalGroup>& grp, int r) : g(grp) { rank = r; nranks = grp->n; }
  ~LocalComm() override { if (tmp) cudaFree(tmp); }
  const char* kind() const override { return "local (peer copies, one process)"; }
  int gone() { set_error("a peer rank of the GPU group failed"); return CATT_E_CUDA; }
  int ensure_tmp(size_t bytes) {
    if (bytes <= tmp_cap) return CATT_OK;
    if (tmp) cudaFree(tmp);
    tmp = nullptr; tmp_cap = 0;
    CATT_CUDA(cudaMalloc(&tmp, bytes + bytes / 4 + 256));
    tmp_cap = bytes + bytes / 4 + 256;
    return CATT_OK;
  }

  int allgatherv(const void* send, void* recv, const size_t* counts, const size_t* displs, cudaStream_t st) override {
    CATT_CUDA(cudaStreamSynchronize(st));                         // my block is complete before a peer reads it
    g->ptr[rank] = send;
    if (!g->barrier()) return gone();
    for (int r = 0; r < nranks; r++)
      if (counts[r] && (const char*)recv + displs[r] != (const char*)g->ptr[r])        // (in place: a rank's own block is already there)
        CATT_CUDA(cudaMemcpyAsync((char*)recv + displs[r], g->ptr[r], counts[r], cudaMemcpyDefault, st));
    CATT_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < nranks; r++) if (r != rank) bytes_moved += (int64_t)counts[r];
    if (!g->barrier()) return gone();                             // nobody reuses its send buffer while a peer still copies
    return CATT_OK;
  }
  int alltoallv(const void* send, const size_t* scounts, const size_t* sdispls, void* recv, const size_t* rcounts, const size_t* rdispls,
                cudaStream_t st) override {
    CATT_CUDA(cudaStreamSynchronize(st));
    g->ptr[rank] = send; g->cnt[rank] = scounts; g->dsp[rank] = sdispls;
    if (!g->barrier()) return gone();
    for (int r = 0; r < nranks; r++) {
      const size_t nb = g->cnt[r][rank];
      if (nb != rcounts[r]) { g->fail(); set_error("alltoallv: rank %d sends %zu bytes to rank %d, which expects %zu", r, nb, rank, rcounts[r]); return CATT_E_ARG; }
      if (nb) CATT_CUDA(cudaMemcpyAsync((char*)recv + rdispls[r], (const char*)g->ptr[r] + g->dsp[r][rank], nb, cudaMemcpyDefault, st));
      if (r != rank) bytes_moved += (int64_t)nb;
    }
    CATT_CUDA(cudaStreamSynchronize(st));
    if (!g->barrier()) return gone();
    return CATT_OK;
  }
  int reduce_impl(uint32_t* buf, size_t count, int root, cudaStream_t st) {
    if (count == 0) return CATT_OK;
    const bool mine = root < 0 || root == rank;
    if (mine) { int rc = ensure_tmp(count * 4); if (rc) { g->fail(); return rc; } }
    CATT_CUDA(cudaStreamSynchronize(st));
    g->ptr[rank] = buf;
    if (!g->barrier()) return gone();
    if (mine) {
      PeerPtrs pp; pp.n = nranks;
      for (int r = 0; r < nranks; r++) pp.p[r] = (const uint32_t*)g->ptr[r];
      const int grid = (int)std::min<size_t>((count / 4 + 255) / 256 + 1, 148 * 8);
      peer_sum_kernel<<<grid, 256, 0, st>>>(pp, (uint32_t*)tmp, count);
      CATT_CUDA(cudaGetLastError());
      CATT_CUDA(cudaStreamSynchronize(st));
      bytes_moved += (int64_t)count * 4 * (nranks - 1);
    }
    if (!g->barrier()) return gone();                             // every reader is done: the inputs may be overwritten
    if (mine) CATT_CUDA(cudaMemcpyAsync(buf, tmp, count * 4, cudaMemcpyDeviceToDevice, st));
    return CATT_OK;
  }
  int allreduce_sum_u32(uint32_t* buf, size_t count, cudaStream_t st) override { return reduce_impl(buf, count, -1, st); }
  int reduce_sum_u32(uint32_t* buf, size_t count, int root, cudaStream_t st) override { return reduce_impl(buf, count, root, st); }
  int host_allgather(const void* in, void* out, size_t bytes, cudaStream_t st) override {
    CATT_CUDA(cudaStreamSynchronize(st));
    if (!g->barrier()) return gone();                             // the previous exchange has been read by everyone
    {
      std::lock_guard<std::mutex> lk(g->mu);
      if (g->host.size() < bytes * (size_t)nranks) g->host.resize(bytes * (size_t)nranks);
    }
    if (!g->barrier()) return gone();
    memcpy(g->host.data() + bytes * (size_t)rank, in, bytes);
    if (!g->barrier()) return gone();
    memcpy(out, g->host.data(), bytes * (size_t)nranks);
    return CATT_OK;
  }
};

// ---------------------------------------------------------------------------------------------------------------------------------
// NCCL, loaded at run time
// ---------------------------------------------------------------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // a process that already holds an NCCL (torch's bundled one under bench.py) gets that one: same SONAME
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) return;
    bool all = true;
    auto sym = [&](const char* s) { void* p = dlsym(api.lib, s); if (!p) all = false; return p; };
    *(void**)&api.GetUniqueId = sym("ncclGetUniqueId");
    *(void**)&api.CommInitRank = sym("ncclCommInitRank");
    *(void**)&api.CommInitAll = sym("ncclCommInitAll");
    *(void**)&api.CommDestroy = sym("ncclCommDestroy");
    *(void**)&api.CommAbort = sym("ncclCommAbort");
    *(void**)&api.GetErrorString = sym("ncclGetErrorString");
    *(void**)&api.GroupStart = sym("ncclGroupStart");
    *(void**)&api.GroupEnd = sym("ncclGroupEnd");
    *(void**)&api.AllGather = sym("ncclAllGather");
    *(void**)&api.AllReduce = sym("ncclAllReduce");
    *(void**)&api.Reduce = sym("ncclReduce");
    *(void**)&api.Broadcast = sym("ncclBroadcast");
    *(void**)&api.Send = sym("ncclSend");
    *(void**)&api.Recv = sym("ncclRecv");
    *(void**)&api.GetVersion = sym("ncclGetVersion");
    api.ok = all;
  });
  return api;
}

#define CATT_NCCL(call)                                                                                         \
  do {                                                                                                          \
    ncclResult_t r__ = (call);                                                                                  \
    if (r__ != ncclSuccess) { set_error("NCCL: %s failed: %s", #call, nccl().GetErrorString(r__)); return CATT_E_CUDA; } \
  } while (0)

struct NcclComm : Comm {
  ncclComm_t comm = nullptr;
  void* stage = nullptr;            // device staging of host_allgather
  void* pin = nullptr;
  size_t stage_cap = 0;
  ~NcclComm() override {
    if (comm) nccl().CommDestroy(comm);
    if (stage) cudaFree(stage);
    if (pin) cudaFreeHost(pin);
  }
  const char* kind() const override { return "nccl"; }
  int allgatherv(const void* send, void* recv, const size_t* counts, const size_t* displs, cudaStream_t st) override {
    NcclApi& N = nccl();
    bool equal = true;
    for (int r = 0; r < nranks; r++) equal = equal && counts[r] == counts[0] && displs[r] == (size_t)r * counts[0];
    if (equal) {
      if (counts[0]) CATT_NCCL(N.AllGather(send, recv, counts[0], ncclChar, comm, st));
    } else {
      CATT_NCCL(N.GroupStart());
      for (int r = 0; r < nranks; r++)
        if (counts[r]) CATT_NCCL(N.Broadcast(send, (char*)recv + displs[r], counts[r], ncclChar, r, comm, st));
      CATT_NCCL(N.GroupEnd());
    }
    for (int r = 0; r < nranks; r++) if (r != rank) bytes_moved += (int64_t)counts[r];
    return CATT_OK;
  }
  int alltoallv(const void* send, const size_t* scounts, const size_t* sdispls, void* recv, const size_t* rcounts, const size_t* rdispls,
                cudaStream_t st) override {
    NcclApi& N = nccl();
    if (scounts[rank] != rcounts[rank]) { set_error("alltoallv: own block sizes differ"); return CATT_E_ARG; }
    if (scounts[rank]) CATT_CUDA(cudaMemcpyAsync((char*)recv + rdispls[rank], (const char*)send + sdispls[rank], scounts[rank], cudaMemcpyDeviceToDevice, st));
    CATT_NCCL(N.GroupStart());
    for (int r = 0; r < nranks; r++) {
      if (r == rank) continue;
      if (scounts[r]) CATT_NCCL(N.Send((const char*)send + sdispls[r], scounts[r], ncclChar, r, comm, st));
      if (rcounts[r]) CATT_NCCL(N.Recv((char*)recv + rdispls[r], rcounts[r], ncclChar, r, comm, st));
      bytes_moved += (int64_t)rcounts[r];
    }
    CATT_NCCL(N.GroupEnd());
    return CATT_OK;
  }
  int allreduce_sum_u32(uint32_t* buf, size_t count, cudaStream_t st) override {
    if (count) CATT_NCCL(nccl().AllReduce(buf, buf, count, ncclUint32, ncclSum, comm, st));
    bytes_moved += (int64_t)count * 4;
    return CATT_OK;
  }
  int reduce_sum_u32(uint32_t* buf, size_t count, int root, cudaStream_t st) override {
    if (count) CATT_NCCL(nccl().Reduce(buf, buf, count, ncclUint32, ncclSum, root, comm, st));
    bytes_moved += (int64_t)count * 4;
    return CATT_OK;
  }
  int host_allgather(const void* in, void* out, size_t bytes, cudaStream_t st) override {
    const size_t b = (bytes + 15) & ~(size_t)15, tot = b * (size_t)(nranks + 1);
    if (tot > stage_cap) {
      if (stage) cudaFree(stage);
      if (pin) cudaFreeHost(pin);
      stage = pin = nullptr; stage_cap = 0;
      CATT_CUDA(cudaMalloc(&stage, tot + 1024));
      CATT_CUDA(cudaMallocHost(&pin, tot + 1024));
      stage_cap = tot + 1024;
    }
    char* h = (char*)pin; char* d = (char*)stage;
    memcpy(h, in, bytes);
    CATT_CUDA(cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, st));
    CATT_NCCL(nccl().AllGather(d, d + b, b, ncclChar, comm, st));
    CATT_CUDA(cudaMemcpyAsync(h + b, d + b, b * (size_t)nranks, cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < nranks; r++) memcpy((char*)out + bytes * (size_t)r, h + b * (size_t)(r + 1), bytes);
    return CATT_OK;
  }
};

}  // namespace

std::shared_ptr<LocalGroup> local_group_create(int nranks, const int* devices) {
  if (nranks < 1 || nranks > MAX_RANKS) return nullptr;
  auto g = std::make_shared<LocalGroup>();
  g->n = nranks;
  g->dev.assign(devices, devices + nranks);
  g->ptr.assign(nranks, nullptr); g->cnt.assign(nranks, nullptr); g->dsp.assign(nranks, nullptr);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int a = 0; a < nranks; a++)                                // direct peer loads for the reduction kernel
    for (int b = 0; b < nranks; b++) {
      if (devices[a] == devices[b]) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can) continue;
      cudaSetDevice(devices[a]);
      if (cudaDeviceEnablePeerAccess(devices[b], 0) != cudaSuccess) cudaGetLastError();    // "already enabled" is fine
    }
  cudaSetDevice(cur);
  return g;
}
Comm* local_comm_create(const std::shared_ptr<LocalGroup>& g, int rank) { return new LocalComm(g, rank); }
void local_group_fail(const std::shared_ptr<LocalGroup>& g) { g->fail(); }
void local_group_reset(const std::shared_ptr<LocalGroup>& g) {
  std::lock_guard<std::mutex> lk(g->mu);
  g->failed = false; g->arrived = 0;
}

bool nccl_available() { return nccl().ok; }

int nccl_unique_id(unsigned char out[128]) {
  if (!nccl().ok) { set_error("libnccl.so.2 could not be loaded"); return CATT_E_CUDA; }
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  CATT_NCCL(nccl().GetUniqueId(&id));
  memcpy(out, &id, 128);
  return CATT_OK;
}

int nccl_comms_create_all(int nranks, const int* devices, std::vector<Comm*>* out) {
  if (!nccl().ok) { set_error("libnccl.so.2 could not be loaded"); return CATT_E_CUDA; }
  if (nranks < 1 || nranks > MAX_RANKS) { set_error("GPU group of %d ranks (limit %d)", nranks, MAX_RANKS); return CATT_E_ARG; }
  std::vector<ncclComm_t> cs((size_t)nranks, nullptr);
  CATT_NCCL(nccl().CommInitAll(cs.data(), nranks, devices));
  out->clear();
  for (int r = 0; r < nranks; r++) {
    NcclComm* c = new NcclComm();
    c->comm = cs[r]; c->rank = r; c->nranks = nranks;
    out->push_back(c);
  }
  return CATT_OK;
}

int nccl_comm_create_rank(int nranks, int rank, const unsigned char id[128], Comm** out) {
  if (!nccl().ok) { set_error("libnccl.so.2 could not be loaded"); return CATT_E_CUDA; }
  if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) { set_error("bad rank %d of %d (limit %d)", rank, nranks, MAX_RANKS); return CATT_E_ARG; }
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclComm_t cm = nullptr;
  CATT_NCCL(nccl().CommInitRank(&cm, nranks, uid, rank));
  NcclComm* c = new NcclComm();
  c->comm = cm; c->rank = rank; c->nranks = nranks;
  *out = c;
  return CATT_OK;
}

}  // namespace catt
