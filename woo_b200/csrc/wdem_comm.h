// wdem_comm.h — transports of the slab-decomposed engine: who moves halo / migration buffers between subdomains.
//   NcclTransport  : one process per GPU; ncclSend/ncclRecv + ncclAllReduce/ncclAllGather over NVLink on the engine's stream.
//   LocalTransport : several subdomains inside ONE process (one host thread per handle), device-to-device copies; used to run the
//                    decomposition on a single GPU (tests) and for multi-GPU single-process hosts.
// Host-side only; no compute here.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace wd {

struct CommErr { std::string msg; };

struct Transport {
  int rank = 0, world = 1;
  virtual ~Transport() {}
  virtual void abort() {}   // wake subdomains blocked in a collective after a local failure (in-process transport only)
  virtual bool multiProcess() const { return false; }   // one process per GPU: device buffers of the peers can be mapped with CUDA IPC
  // in-place max over ranks of n ints in device memory, ordered on the stream
  virtual void allreduceMaxInt(int* dptr, int n, cudaStream_t st) = 0;
  // every rank contributes n host int64 and receives world*n (rank-major); blocks the host
  virtual void allgather64(const int64_t* mine, int n, int64_t* all, cudaStream_t st) = 0;
  // every rank contributes `bytes` of DEVICE memory and receives world*bytes (rank-major) in device memory; ordered on the stream, no host wait
  virtual void allgatherDev(const void* dSend, void* dRecv, size_t bytes, cudaStream_t st) = 0;
  // device payload exchange, sizes known on both sides; segment r of send goes to rank r, segment r of recv comes from rank r
  virtual void alltoallv(const char* sendBase, const size_t* sendOff, const size_t* sendBytes, char* recvBase, const size_t* recvOff,
                         const size_t* recvBytes, cudaStream_t st) = 0;
};

#define WD_CUDA(call)                                                                             \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) throw CommErr{std::string(#call) + ": " + cudaGetErrorString(e__)}; \
  } while (0)
#define WD_NCCL(call)                                                                             \
  do {                                                                                            \
    ncclResult_t r__ = (call);                                                                    \
    if (r__ != ncclSuccess) throw CommErr{std::string(#call) + ": " + ncclGetErrorString(r__)}; \
  } while (0)

struct NcclTransport : Transport {
  ncclComm_t comm = nullptr;
  int64_t* dStage = nullptr; int64_t* hStage = nullptr; size_t stageCap = 0;
  NcclTransport(int rank_, int world_, const ncclUniqueId& id) {
    rank = rank_; world = world_;
    WD_NCCL(ncclCommInitRank(&comm, world, id, rank));
  }
  ~NcclTransport() override {
    if (comm) ncclCommDestroy(comm);
    if (dStage) cudaFree(dStage);
    if (hStage) cudaFreeHost(hStage);
  }
  bool multiProcess() const override { return true; }
  void allreduceMaxInt(int* dptr, int n, cudaStream_t st) override { WD_NCCL(ncclAllReduce(dptr, dptr, n, ncclInt32, ncclMax, comm, st)); }
  void allgather64(const int64_t* mine, int n, int64_t* all, cudaStream_t st) override {
    const size_t need = (size_t)n * (world + 1);
    if (need > stageCap) {
      if (dStage) cudaFree(dStage);
      if (hStage) cudaFreeHost(hStage);
      stageCap = need * 2 + 64;
      WD_CUDA(cudaMalloc(&dStage, stageCap * sizeof(int64_t)));
      WD_CUDA(cudaMallocHost(&hStage, stageCap * sizeof(int64_t)));
    }
    memcpy(hStage, mine, n * sizeof(int64_t));
    WD_CUDA(cudaMemcpyAsync(dStage, hStage, n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    WD_NCCL(ncclAllGather(dStage, dStage + n, n, ncclInt64, comm, st));
    WD_CUDA(cudaMemcpyAsync(hStage + n, dStage + n, (size_t)n * world * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    memcpy(all, hStage + n, (size_t)n * world * sizeof(int64_t));
  }
  void allgatherDev(const void* dSend, void* dRecv, size_t bytes, cudaStream_t st) override { WD_NCCL(ncclAllGather(dSend, dRecv, bytes, ncclChar, comm, st)); }
  void alltoallv(const char* sendBase, const size_t* sendOff, const size_t* sendBytes, char* recvBase, const size_t* recvOff, const size_t* recvBytes,
                 cudaStream_t st) override {
    WD_NCCL(ncclGroupStart());
    for (int r = 0; r < world; r++) {
      if (r == rank) continue;
      if (sendBytes[r]) WD_NCCL(ncclSend(sendBase + sendOff[r], sendBytes[r], ncclChar, r, comm, st));
      if (recvBytes[r]) WD_NCCL(ncclRecv(recvBase + recvOff[r], recvBytes[r], ncclChar, r, comm, st));
    }
    WD_NCCL(ncclGroupEnd());
  }
};

// Shared state of the subdomains living in one process.  Every collective is entered by all `world` threads.
struct LocalGroup {
  int world;
  std::mutex m; std::condition_variable cv; int waiting = 0; long gen = 0; bool broken = false;
  std::vector<const char*> sendBase; std::vector<const size_t*> sendOff;
  std::vector<std::vector<int64_t>> slots; std::vector<std::vector<char>> blobs;
  explicit LocalGroup(int w) : world(w), sendBase(w, nullptr), sendOff(w, nullptr), slots(w), blobs(w) {}
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (broken) throw CommErr{"local group: another subdomain failed"};
    long g = gen;
    if (++waiting == world) { waiting = 0; gen++; cv.notify_all(); }
    else {
      cv.wait(lk, [&] { return gen != g || broken; });
      if (gen == g && broken) throw CommErr{"local group: another subdomain failed"};
    }
  }
  void abort() { std::unique_lock<std::mutex> lk(m); broken = true; cv.notify_all(); }
};

struct LocalTransport : Transport {
  LocalGroup* G;
  LocalTransport(LocalGroup* g, int rank_) : G(g) { rank = rank_; world = g->world; }
  void abort() override { G->abort(); }
  void allgather64(const int64_t* mine, int n, int64_t* all, cudaStream_t st) override {
    (void)st;
    G->slots[rank].assign(mine, mine + n);
    G->barrier();
    for (int r = 0; r < world; r++) memcpy(all + (size_t)r * n, G->slots[r].data(), n * sizeof(int64_t));
    G->barrier();
  }
  void allgatherDev(const void* dSend, void* dRecv, size_t bytes, cudaStream_t st) override {
    std::vector<char> mine(bytes);
    WD_CUDA(cudaMemcpyAsync(mine.data(), dSend, bytes, cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    G->blobs[rank] = mine;
    G->barrier();
    std::vector<char> all(bytes * world);
    for (int r = 0; r < world; r++) memcpy(all.data() + (size_t)r * bytes, G->blobs[r].data(), bytes);
    G->barrier();
    WD_CUDA(cudaMemcpyAsync(dRecv, all.data(), all.size(), cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaStreamSynchronize(st));
  }
  void allreduceMaxInt(int* dptr, int n, cudaStream_t st) override {
    std::vector<int> h(n);
    WD_CUDA(cudaMemcpyAsync(h.data(), dptr, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    WD_CUDA(cudaStreamSynchronize(st));
    std::vector<int64_t> mine(h.begin(), h.end()), all((size_t)n * world);
    allgather64(mine.data(), n, all.data(), st);
    for (int k = 0; k < n; k++) { int64_t v = all[k]; for (int r = 1; r < world; r++) v = std::max(v, all[(size_t)r * n + k]); h[k] = (int)v; }
    WD_CUDA(cudaMemcpyAsync(dptr, h.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
    WD_CUDA(cudaStreamSynchronize(st));
  }
  void alltoallv(const char* sendBase, const size_t* sendOff, const size_t* sendBytes, char* recvBase, const size_t* recvOff, const size_t* recvBytes,
                 cudaStream_t st) override {
    (void)sendBytes;
    WD_CUDA(cudaStreamSynchronize(st));   // my send buffer is complete
    G->sendBase[rank] = sendBase; G->sendOff[rank] = sendOff;
    G->barrier();
    for (int r = 0; r < world; r++) {
      if (r == rank || !recvBytes[r]) continue;
      WD_CUDA(cudaMemcpyAsync(recvBase + recvOff[r], G->sendBase[r] + G->sendOff[r][rank], recvBytes[r], cudaMemcpyDefault, st));
    }
    WD_CUDA(cudaStreamSynchronize(st));
    G->barrier();                          // senders may reuse their buffers
  }
};

}  // namespace wd
