// NCCL binding (run-time, see dist.cuh) and the three communication patterns of the slab path.
#include "dist.cuh"

#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>

namespace econpdes {
namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) {
      err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    auto sym = [&](const char* name) {
      void* p = dlsym(h, name);
      if (p == nullptr && err.empty()) err = std::string("libnccl has no symbol ") + name;
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!err.empty()) throw NcclError(err);
  return a;
}

#define EP_NCCL(expr)                                                                                     \
  do {                                                                                                    \
    ncclResult_t _r = (expr);                                                                             \
    if (_r != ncclSuccess)                                                                                \
      throw NcclError(std::string(#expr) + ": " + api().GetErrorString(_r) + " (" + __FILE__ + ":" +      \
                      std::to_string(__LINE__) + ")");                                                    \
  } while (0)

static_assert(sizeof(ncclUniqueId) == kUniqueIdBytes, "unique id size");

}  // namespace

void Dist::unique_id(uint8_t* id128) {
  ncclUniqueId id;
  EP_NCCL(api().GetUniqueId(&id));
  std::memcpy(id128, id.internal, kUniqueIdBytes);
}

Dist::Dist(int rank, int world, const uint8_t* id128) : rank_(rank), world_(world) {
  if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("rank/world out of range");
  ncclUniqueId id;
  std::memcpy(id.internal, id128, kUniqueIdBytes);
  ncclComm_t comm = nullptr;
  EP_NCCL(api().CommInitRank(&comm, world, id, rank));
  comm_ = comm;
}

Dist::~Dist() {
  if (comm_) api().CommDestroy(static_cast<ncclComm_t>(comm_));
}

void Dist::exchange_halo(double* v, long long plane, long long nplanes, cudaStream_t s) {
  if (world_ == 1) return;
  const NcclApi& n = api();
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  const size_t cnt = static_cast<size_t>(plane);
  EP_NCCL(n.GroupStart());
  if (has_lo()) {
    EP_NCCL(n.Send(v, cnt, ncclDouble, rank_ - 1, comm, s));
    EP_NCCL(n.Recv(v - plane, cnt, ncclDouble, rank_ - 1, comm, s));
  }
  if (has_hi()) {
    EP_NCCL(n.Send(v + (nplanes - 1) * plane, cnt, ncclDouble, rank_ + 1, comm, s));
    EP_NCCL(n.Recv(v + nplanes * plane, cnt, ncclDouble, rank_ + 1, comm, s));
  }
  EP_NCCL(n.GroupEnd());
  ++ncoll_;
}

void Dist::send_lo_recv_hi(const double* send_first, double* recv_hi_ghost, long long count, cudaStream_t s) {
  if (world_ == 1) return;
  const NcclApi& n = api();
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  EP_NCCL(n.GroupStart());
  if (has_lo()) EP_NCCL(n.Send(send_first, static_cast<size_t>(count), ncclDouble, rank_ - 1, comm, s));
  if (has_hi()) EP_NCCL(n.Recv(recv_hi_ghost, static_cast<size_t>(count), ncclDouble, rank_ + 1, comm, s));
  EP_NCCL(n.GroupEnd());
  ++ncoll_;
}

void Dist::send_hi_recv_lo(const double* send_last, double* recv_lo_ghost, long long count, cudaStream_t s) {
  if (world_ == 1) return;
  const NcclApi& n = api();
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  EP_NCCL(n.GroupStart());
  if (has_hi()) EP_NCCL(n.Send(send_last, static_cast<size_t>(count), ncclDouble, rank_ + 1, comm, s));
  if (has_lo()) EP_NCCL(n.Recv(recv_lo_ghost, static_cast<size_t>(count), ncclDouble, rank_ - 1, comm, s));
  EP_NCCL(n.GroupEnd());
  ++ncoll_;
}

void Dist::all_gather(const double* send, double* recv, long long count, cudaStream_t s) {
  if (world_ == 1) {
    if (send != recv)
      cudaMemcpyAsync(recv, send, static_cast<size_t>(count) * sizeof(double), cudaMemcpyDeviceToDevice, s);
    return;
  }
  EP_NCCL(api().AllGather(send, recv, static_cast<size_t>(count), ncclDouble, static_cast<ncclComm_t>(comm_), s));
  ++ncoll_;
}

}  // namespace econpdes

// ---- phase profile (common.cuh) ----
namespace econpdes {
namespace {
std::map<std::string, std::pair<double, long>>& prof_table() {
  static std::map<std::string, std::pair<double, long>> t;
  return t;
}
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace
bool Prof::enabled() {
  static const bool on = [] { const char* e = std::getenv("ECONPDES_PROFILE"); return e && e[0] == '1'; }();
  return on;
}
void Prof::add(const char* name, double ms) {
  auto& e = prof_table()[name];
  e.first += ms;
  e.second += 1;
}
void Prof::report(int rank) {
  if (!enabled()) return;
  for (const auto& kv : prof_table())
    std::fprintf(stderr, "[econpdes profile rank %d] %-22s %10.3f ms %8ld calls\n", rank, kv.first.c_str(),
                 kv.second.first, kv.second.second);
  prof_table().clear();
}
ProfScope::ProfScope(const char* n, cudaStream_t st) : name(n), s(st) {
  if (!Prof::enabled() || name[0] == 0) return;
  cudaStreamSynchronize(s);
  t0 = now_ms();
}
ProfScope::~ProfScope() {
  if (!Prof::enabled() || name[0] == 0) return;
  cudaStreamSynchronize(s);
  Prof::add(name, now_ms() - t0);
}
}  // namespace econpdes
