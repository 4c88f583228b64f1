// shard.cu -- multi-GPU sharding of the path behind the C ABI: output-row bands, one rank per GPU, and
// the halo rows a band needs from its lower neighbour moved with ncclSend / ncclRecv inside one NCCL group over
// NVLink / NVSwitch.  The unit of the path is the output tile; bands are independent, so this exchange of INPUT rows is the
// only communication (no reduction anywhere on the path).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- already resident in a process that imported torch, or found on the
// loader path of a C++ host), so libvwb200.so keeps loading on machines without NCCL; only vwb200_shard_create needs it.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_FLOAT32 = 7 };          // ncclDataType_t (nccl.h)
struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};
NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (const char* p = getenv("VWB200_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);       // torch's copy, if it is in the process
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("libnccl.so.2 not found (") + (dlerror() ? dlerror() : "?") + "); set VWB200_NCCL_LIB"; return; }
#define VWB_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); if (!api.field) { api.why = std::string("missing ") + name; return; }
    VWB_SYM(GetUniqueId, "ncclGetUniqueId") VWB_SYM(CommInitRank, "ncclCommInitRank") VWB_SYM(CommDestroy, "ncclCommDestroy")
    VWB_SYM(Send, "ncclSend") VWB_SYM(Recv, "ncclRecv") VWB_SYM(GroupStart, "ncclGroupStart") VWB_SYM(GroupEnd, "ncclGroupEnd")
    VWB_SYM(GetErrorString, "ncclGetErrorString")
#undef VWB_SYM
    api.ok = true;
  });
  return api;
}
int need_nccl() {
  if (nccl().ok) return VWB200_OK;
  vwb200::set_error("NCCL is not available: %s", nccl().why.c_str());
  return VWB200_ENOIMPL;
}
#define VWB_NCCL(expr)                                                                                             \
  do {                                                                                                             \
    int _r = (expr);                                                                                               \
    if (_r != 0) { vwb200::set_error("%s failed: %s", #expr, nccl().GetErrorString(_r)); return VWB200_ECUDA; }    \
  } while (0)
}  // namespace

struct vwb200_shard { ncclComm_t comm = nullptr; int rank = 0, world = 1; };

extern "C" {

int vwb200_shard_plan(int rank, int world, int out_rows, int ky, int sy, int left_total_rows, int right_total_rows, vwb200_band_plan* p) {
  if (!p || world < 1 || rank < 0 || rank >= world || out_rows <= 0 || ky < 1 || sy < 1) { vwb200::set_error("shard_plan: bad arguments"); return VWB200_EARG; }
  if (left_total_rows <= 0) left_total_rows = out_rows + ky - 1;
  if (right_total_rows <= 0) right_total_rows = out_rows + ky - 1 + sy - 1;
  const int band = (out_rows + world - 1) / world;
  if (band < ky - 1 + sy - 1 && world > 1) {
    vwb200::set_error("shard_plan: band height %d smaller than the halo %d: a rank would need rows from beyond its neighbour", band, ky - 1 + sy - 1);
    return VWB200_EARG;
  }
  const int y0 = std::min(out_rows, rank * band), y1 = std::min(out_rows, (rank + 1) * band), h = y1 - y0;
  p->rank = rank; p->world = world; p->y0 = y0; p->y1 = y1;
  p->left_rows = h + ky - 1; p->right_rows = h + ky - 1 + sy - 1;
  const bool last = rank == world - 1 || y1 >= out_rows;
  p->own_left = std::min(p->left_rows, last ? left_total_rows - y0 : h);
  p->own_right = std::min(p->right_rows, last ? right_total_rows - y0 : h);
  p->recv_left = p->left_rows - p->own_left; p->recv_right = p->right_rows - p->own_right;
  p->send_left = (rank > 0 && h > 0) ? ky - 1 : 0;
  p->send_right = (rank > 0 && h > 0) ? ky - 1 + sy - 1 : 0;
  return VWB200_OK;
}

int vwb200_shard_unique_id(void* id128) {
  if (!id128) { vwb200::set_error("shard_unique_id: null pointer"); return VWB200_EARG; }
  VWB_TRY(need_nccl());
  ncclUniqueId id;
  VWB_NCCL(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return VWB200_OK;
}

int vwb200_shard_create(const void* id128, int rank, int world, vwb200_shard** out) {
  if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) { vwb200::set_error("shard_create: bad arguments"); return VWB200_EARG; }
  vwb200_shard* s = new vwb200_shard();
  s->rank = rank; s->world = world;
  if (world > 1) {
    VWB_TRY(vwb200::ensure_device());
    int rc = need_nccl();
    if (rc) { delete s; return rc; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    const int r = nccl().CommInitRank(&s->comm, world, id, rank);
    if (r != 0) { vwb200::set_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r)); delete s; return VWB200_ECUDA; }
  }
  *out = s;
  return VWB200_OK;
}

// left_band / right_band: device rasters of plan->left_rows / right_rows rows (pitch in elements) whose first own_* rows are
// valid; fills the halo rows from rank + 1 and serves rank - 1, all inside one NCCL group, on `stream` (asynchronous).
int vwb200_shard_exchange_halos(vwb200_shard* s, const vwb200_band_plan* p, float* left_band, int lcols, ptrdiff_t lpitch, float* right_band,
                                int rcols, ptrdiff_t rpitch, void* stream) {
  if (!s || !p || !left_band || !right_band) { vwb200::set_error("shard_exchange_halos: null pointer"); return VWB200_EARG; }
  if (s->world == 1) return VWB200_OK;
  if (p->rank != s->rank || p->world != s->world) { vwb200::set_error("shard_exchange_halos: the plan belongs to another rank"); return VWB200_EARG; }
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : cudaStreamLegacy;
  auto rows = [&](bool send, float* base, int cols, ptrdiff_t pitch, int row0, int n, int peer) -> int {
    if (n <= 0) return VWB200_OK;
    if (pitch == cols) {
      if (send) VWB_NCCL(nccl().Send(base + (ptrdiff_t)row0 * pitch, (size_t)n * cols, NCCL_FLOAT32, peer, s->comm, st));
      else VWB_NCCL(nccl().Recv(base + (ptrdiff_t)row0 * pitch, (size_t)n * cols, NCCL_FLOAT32, peer, s->comm, st));
    } else {
      for (int r = 0; r < n; ++r) {
        if (send) VWB_NCCL(nccl().Send(base + (ptrdiff_t)(row0 + r) * pitch, (size_t)cols, NCCL_FLOAT32, peer, s->comm, st));
        else VWB_NCCL(nccl().Recv(base + (ptrdiff_t)(row0 + r) * pitch, (size_t)cols, NCCL_FLOAT32, peer, s->comm, st));
      }
    }
    return VWB200_OK;
  };
  VWB_NCCL(nccl().GroupStart());
  int rc = rows(false, left_band, lcols, lpitch, p->own_left, p->recv_left, s->rank + 1);
  if (!rc) rc = rows(false, right_band, rcols, rpitch, p->own_right, p->recv_right, s->rank + 1);
  if (!rc) rc = rows(true, left_band, lcols, lpitch, 0, p->send_left, s->rank - 1);
  if (!rc) rc = rows(true, right_band, rcols, rpitch, 0, p->send_right, s->rank - 1);
  VWB_NCCL(nccl().GroupEnd());
  return rc;
}

void vwb200_shard_destroy(vwb200_shard* s) {
  if (!s) return;
  if (s->comm && nccl().ok) nccl().CommDestroy(s->comm);
  delete s;
}

}  // extern "C"
