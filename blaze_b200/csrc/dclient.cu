// DriverClient half of the C ABI (include/blaze_b200.h): device selection, the card address space, the
// FPGA-management stubs, multi-device members and the NCCL communicator.
//
// Mirrors (behaviour, not code) /root/reference/src/driver_client/dclient.rs: `DriverClient::new(id, cfg)` (:79-86)
// opens a card, `dma_write` / `dma_read` (:456-517) move bytes to / from its flat address space, `reset` (:88-93)
// restarts the user logic and keeps HBM contents.  XDMA pwrite/pread become cudaMemcpyAsync into virtual-memory-managed
// HBM.  No CPU fallback: without a device the constructor fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "api_common.h"
#include "client_internal.h"
#include "msm_internal.h"

using namespace bz;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

int32_t bz_fail(int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
void bz_set_last_error(const std::string& s) { g_last_error = s; }

std::atomic<uint64_t> bz::g_kernel_launches{0};
extern "C" uint64_t bz_kernel_launch_count(void) { return bz::g_kernel_launches.load(); }
extern "C" const char* bz_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* bz_version(void) { return "blaze_b200 0.2.0 sm_100a"; }

// ------------------------------------------------------------------------------------ driver entry points (VMM)
// Resolved through the runtime (cudaGetDriverEntryPoint): the library does not link libcuda, so it still loads on a
// machine without a driver (where every constructor then reports BZ_ERR_NO_DEVICE).
namespace {
struct Drv {
  CUresult (*addressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};
Drv g_drv;
std::once_flag g_drv_once;

template <class T>
bool drv_sym(const char* name, T& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<T>(p);
  return true;
}
const Drv& drv() {
  std::call_once(g_drv_once, [] {
    bool ok = drv_sym("cuMemAddressReserve", g_drv.addressReserve);
    ok = drv_sym("cuMemAddressFree", g_drv.addressFree) && ok;
    ok = drv_sym("cuMemCreate", g_drv.create) && ok;
    ok = drv_sym("cuMemRelease", g_drv.release) && ok;
    ok = drv_sym("cuMemMap", g_drv.map) && ok;
    ok = drv_sym("cuMemUnmap", g_drv.unmap) && ok;
    ok = drv_sym("cuMemSetAccess", g_drv.setAccess) && ok;
    ok = drv_sym("cuMemGetAllocationGranularity", g_drv.granularity) && ok;
    g_drv.ok = ok;
  });
  return g_drv;
}
}  // namespace

// ------------------------------------------------------------------------------------ card address space
// [0, 2^40) is HBM.  The MSM stream ports of the reference (msm_cfg.rs:44-92: 0x0000_0100_0000_0000 /
// 0x0000_0200_0000_0000) lie at and above 2^40 and are served by bz_msm_set_data, never by dma_write.
static const uint64_t ARENA_LIMIT = 1ull << 40;
static const size_t ARENA_CHUNK = 256ull << 20;
static const size_t WLOG_MAX = 256;

static CUmemAllocationProp arena_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  return prop;
}

static int32_t arena_open(bz_dclient* dc) {
  const Drv& d = drv();
  if (!d.ok) return bz_fail(BZ_ERR_NO_DEVICE, "CUDA virtual memory management entry points are not available in this driver");
  CUmemAllocationProp prop = arena_prop(dc->device);
  size_t gran = 0;
  if (d.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) gran = 2ull << 20;
  dc->chunk = (ARENA_CHUNK + gran - 1) / gran * gran;
  // the whole window if the driver grants it, else progressively less (a B200 has 180 GB to back it anyway)
  for (uint64_t va = ARENA_LIMIT; va >= (64ull << 30); va >>= 1) {
    CUdeviceptr p = 0;
    if (d.addressReserve(&p, (size_t)va, dc->chunk, 0, 0) == CUDA_SUCCESS) {
      dc->arena = reinterpret_cast<uint8_t*>(p);
      dc->arena_reserved = va;
      dc->arena_va = va / dc->chunk * dc->chunk;
      dc->mapped.assign((size_t)(dc->arena_va / dc->chunk), 0);
      dc->handles.assign(dc->mapped.size(), 0);
      return BZ_OK;
    }
  }
  return bz_fail(BZ_ERR_NO_DEVICE, "cannot reserve virtual address space for the card's HBM window");
}

static void arena_close(bz_dclient* dc) {
  const Drv& d = drv();
  if (!dc->arena || !d.ok) return;
  for (size_t c = 0; c < dc->mapped.size(); c++)
    if (dc->mapped[c]) {
      d.unmap((CUdeviceptr)(uintptr_t)(dc->arena + c * dc->chunk), dc->chunk);
      d.release((CUmemGenericAllocationHandle)dc->handles[c]);
    }
  d.addressFree((CUdeviceptr)(uintptr_t)dc->arena, (size_t)dc->arena_reserved);
  dc->arena = nullptr;
}

int32_t bz::arena_map(bz_dclient* dc, uint64_t lo, uint64_t hi) {
  if (hi > ARENA_LIMIT || hi < lo) return bz_fail(BZ_ERR_WRITE, "address 0x%llx beyond the HBM window", (unsigned long long)hi);
  if (hi > dc->arena_va) return bz_fail(BZ_ERR_WRITE, "address 0x%llx beyond the %llu GiB this driver lets the card map", (unsigned long long)hi, (unsigned long long)(dc->arena_va >> 30));
  if (hi == lo) return BZ_OK;
  const Drv& d = drv();
  CUmemAllocationProp prop = arena_prop(dc->device);
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (uint64_t c = lo / dc->chunk; c <= (hi - 1) / dc->chunk; c++) {
    if (dc->mapped[c]) continue;
    CUmemGenericAllocationHandle h;
    CUresult r = d.create(&h, dc->chunk, &prop, 0);
    if (r != CUDA_SUCCESS) return bz_fail(BZ_ERR_WRITE, "out of HBM: cannot back card address 0x%llx (cuMemCreate error %d)", (unsigned long long)(c * dc->chunk), (int)r);
    CUdeviceptr at = (CUdeviceptr)(uintptr_t)(dc->arena + c * dc->chunk);
    if ((r = d.map(at, dc->chunk, 0, h, 0)) != CUDA_SUCCESS || (r = d.setAccess(at, dc->chunk, &acc, 1)) != CUDA_SUCCESS) {
      d.release(h);
      return bz_fail(BZ_ERR_WRITE, "cannot map card address 0x%llx (error %d)", (unsigned long long)(c * dc->chunk), (int)r);
    }
    dc->handles[c] = (unsigned long long)h;
    dc->mapped[c] = 1;
    CUDA_TRY(BZ_ERR_WRITE, cudaMemsetAsync(dc->arena + c * dc->chunk, 0, dc->chunk, dc->stream));   // never-written HBM reads as zeros
  }
  return BZ_OK;
}

void bz::arena_note_write(bz_dclient* dc, uint64_t lo, uint64_t hi) {
  dc->epoch++;
  dc->wlog.push_back({dc->epoch, lo, hi});
  while (dc->wlog.size() > WLOG_MAX) {
    dc->wlog_floor = dc->wlog.front().epoch;
    dc->wlog.pop_front();
  }
}

bool bz::arena_dirty_since(bz_dclient* dc, uint64_t epoch, uint64_t lo, uint64_t hi) {
  if (epoch < dc->wlog_floor) return true;   // older than the log remembers: assume the worst
  for (auto it = dc->wlog.rbegin(); it != dc->wlog.rend() && it->epoch > epoch; ++it)
    if (it->lo < hi && lo < it->hi) return true;
  return false;
}

// ------------------------------------------------------------------------------------ handle helpers
int32_t dc_select(bz_dclient* dc) {
  if (!dc) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient");
  CUDA_TRY(BZ_ERR_NO_DEVICE, cudaSetDevice(dc->device));
  return BZ_OK;
}
cudaStream_t dc_stream(bz_dclient* dc) { return dc->stream; }
int dc_device(bz_dclient* dc) { return dc->device; }

static int32_t dclient_open_one(int dev, int card_type, bz_dclient** out) {
  bz_dclient* dc = new bz_dclient();
  dc->device = dev;
  dc->card_type = card_type;
  if (cudaSetDevice(dev) != cudaSuccess || cudaFree(0) != cudaSuccess ||
      cudaStreamCreateWithFlags(&dc->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete dc;
    return bz_fail(BZ_ERR_NO_DEVICE, "cannot initialise device %d: %s", dev, cudaGetErrorString(cudaGetLastError()));
  }
  int32_t rc = arena_open(dc);
  if (rc) { cudaStreamDestroy(dc->stream); delete dc; return rc; }
  *out = dc;
  return BZ_OK;
}

static void dclient_close_one(bz_dclient* dc) {
  cudaSetDevice(dc->device);
  if (dc->stream) cudaStreamSynchronize(dc->stream);
  if (dc->comm_scratch) cudaFree(dc->comm_scratch);
  arena_close(dc);
  if (dc->stream) cudaStreamDestroy(dc->stream);
  delete dc;
}

// `id`: the reference's FPGA slot string (dclient.rs:79-86, env ID) = CUDA device ordinal; a comma-separated list
// ("0,1,2,3") opens one multi-device client (device list for the multi-GPU configs).
extern "C" int32_t bz_dclient_new(const char* id, int32_t card_type, bz_dclient** out) {
  if (!out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return bz_fail(BZ_ERR_NO_DEVICE, "no CUDA device (%s); blaze_b200 has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  std::vector<int> devs;
  {
    const char* p = id && *id ? id : "0";
    while (*p) {
      char* end = nullptr;
      long v = strtol(p, &end, 10);
      if (end == p || v < 0 || v >= ndev) return bz_fail(BZ_ERR_NO_DEVICE, "device id '%s' is not a list of ordinals below %d", id, ndev);
      devs.push_back((int)v);
      p = end;
      if (*p == ',') p++;
      else if (*p) return bz_fail(BZ_ERR_NO_DEVICE, "device id '%s' is not a comma-separated list of ordinals", id);
    }
  }
  if (devs.empty() || devs.size() > 16) return bz_fail(BZ_ERR_NO_DEVICE, "device id '%s': 1 to 16 devices", id ? id : "");
  bz_dclient* dc = nullptr;
  int32_t rc = dclient_open_one(devs[0], card_type, &dc);
  if (rc) return rc;
  for (size_t g = 1; g < devs.size(); g++) {
    bz_dclient* k = nullptr;
    rc = dclient_open_one(devs[g], card_type, &k);
    if (rc) { bz_dclient_free(dc); return rc; }
    k->is_member = true;
    dc->peers.push_back(k);
  }
  cudaSetDevice(dc->device);
  *out = dc;
  return BZ_OK;
}

extern "C" int32_t bz_dclient_device_count(bz_dclient* dc, uint32_t* n) {
  if (!dc || !n) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null argument");
  *n = (uint32_t)dc_members(dc);
  return BZ_OK;
}

// ------------------------------------------------------------------------------------ NCCL (dlopen)
namespace {
struct Nccl {
  ncclResult_t (*getUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*commInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*commDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*allGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*allReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*getErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};
Nccl g_nccl;
std::once_flag g_nccl_once;
const Nccl& nccl() {
  std::call_once(g_nccl_once, [] {
    // a process that already carries NCCL (e.g. under torch.distributed) gets that copy; otherwise the system one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      const char* why = dlerror();   // one call: dlerror() clears the message it returns
      g_nccl.why = why ? why : "libnccl.so.2 not found";
      return;
    }
    auto sym = [&](const char* n) { return dlsym(h, n); };
    g_nccl.getUniqueId = reinterpret_cast<decltype(g_nccl.getUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.commInitRank = reinterpret_cast<decltype(g_nccl.commInitRank)>(sym("ncclCommInitRank"));
    g_nccl.commDestroy = reinterpret_cast<decltype(g_nccl.commDestroy)>(sym("ncclCommDestroy"));
    g_nccl.allGather = reinterpret_cast<decltype(g_nccl.allGather)>(sym("ncclAllGather"));
    g_nccl.allReduce = reinterpret_cast<decltype(g_nccl.allReduce)>(sym("ncclAllReduce"));
    g_nccl.getErrorString = reinterpret_cast<decltype(g_nccl.getErrorString)>(sym("ncclGetErrorString"));
    g_nccl.ok = g_nccl.getUniqueId && g_nccl.commInitRank && g_nccl.commDestroy && g_nccl.allGather && g_nccl.allReduce &&
                g_nccl.getErrorString;
    if (!g_nccl.ok) g_nccl.why = "libnccl.so.2 lacks a required symbol";
  });
  return g_nccl;
}
}  // namespace

#define NCCL_TRY(expr)                                                                                        \
  do {                                                                                                        \
    ncclResult_t _r = (expr);                                                                                 \
    if (_r != ncclSuccess) return bz_fail(BZ_ERR_UNKNOWN, "%s failed: %s", #expr, nccl().getErrorString(_r)); \
  } while (0)

extern "C" int32_t bz_comm_unique_id(uint8_t out[128]) {
  if (!out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "out is null");
  const Nccl& n = nccl();
  if (!n.ok) return bz_fail(BZ_ERR_NO_DEVICE, "NCCL is not available: %s", n.why.c_str());
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id size");
  ncclUniqueId id;
  NCCL_TRY(n.getUniqueId(&id));
  memcpy(out, &id, 128);
  return BZ_OK;
}

extern "C" int32_t bz_dclient_comm_init(bz_dclient* dc, int32_t rank, int32_t world, const uint8_t unique_id[128]) {
  int32_t rc = dc_select(dc);
  if (rc) return rc;
  if (!unique_id || world < 1 || rank < 0 || rank >= world) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "bad rank/world %d/%d", rank, world);
  if (!dc->peers.empty()) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "a multi-device client cannot also be a rank of a process world");
  const Nccl& n = nccl();
  if (!n.ok) return bz_fail(BZ_ERR_NO_DEVICE, "NCCL is not available: %s", n.why.c_str());
  std::lock_guard<std::mutex> lk(dc->mu);
  if (dc->comm) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, unique_id, 128);
  ncclComm_t comm = nullptr;
  NCCL_TRY(n.commInitRank(&comm, world, id, rank));
  dc->comm = comm;
  dc->rank = rank;
  dc->world = world;
  CUDA_TRY(BZ_ERR_WRITE, cudaMalloc((void**)&dc->comm_scratch, 256));
  CUDA_TRY(BZ_ERR_WRITE, cudaMemset(dc->comm_scratch, 0, 256));
  return BZ_OK;
}

extern "C" int32_t bz_dclient_comm_info(bz_dclient* dc, int32_t* rank, int32_t* world) {
  if (!dc) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient");
  if (rank) *rank = dc->rank;
  if (world) *world = dc->world;
  return BZ_OK;
}

int32_t bz::comm_allgather(bz_dclient* dc, const void* send, void* recv, size_t bytes, cudaStream_t st) {
  if (!dc->comm) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "client has no communicator (bz_dclient_comm_init)");
  NCCL_TRY(nccl().allGather(send, recv, bytes, ncclUint8, (ncclComm_t)dc->comm, st));
  return BZ_OK;
}
int32_t bz::comm_barrier(bz_dclient* dc, cudaStream_t st) {
  if (!dc->comm) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "client has no communicator (bz_dclient_comm_init)");
  NCCL_TRY(nccl().allReduce(dc->comm_scratch, dc->comm_scratch + 32, 1, ncclInt32, ncclSum, (ncclComm_t)dc->comm, st));
  return BZ_OK;
}

extern "C" int32_t bz_dclient_free(bz_dclient* dc) {
  if (!dc) return BZ_OK;
  for (bz_dclient* k : dc->peers) dclient_close_one(k);
  dc->peers.clear();
  if (dc->comm) {
    cudaSetDevice(dc->device);
    cudaStreamSynchronize(dc->stream);
    nccl().commDestroy((ncclComm_t)dc->comm);
    dc->comm = nullptr;
  }
  dclient_close_one(dc);
  return BZ_OK;
}

extern "C" int32_t bz_dclient_reset(bz_dclient* dc) {
  // The reference toggles the DFX decoupler (dclient.rs:88-93): user logic is reset, HBM contents survive.
  // Here: drain the device(s) -- work streams and the clients' copy / tail streams; the address space keeps its bytes.
  if (!dc) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient");
  for (int g = dc_members(dc) - 1; g >= 0; g--) {
    bz_dclient* k = dc_member(dc, g);
    int32_t rc = dc_select(k);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(k->mu);
    CUDA_TRY(BZ_ERR_UNKNOWN, cudaDeviceSynchronize());
  }
  return BZ_OK;
}

// dma_write / dma_read on a multi-device client address the FIRST device's HBM (an MSMClient on such a client shards
// its own point set over the members through load_data_to_hbm, see msm_api.cu).
extern "C" int32_t bz_dclient_dma_write(bz_dclient* dc, uint64_t base, uint64_t offset, const uint8_t* data, size_t len) {
  int32_t rc = dc_select(dc);
  if (rc) return rc;
  if (!data && len) return bz_fail(BZ_ERR_WRITE, "null data");
  std::lock_guard<std::mutex> lk(dc->mu);
  uint64_t a = base + offset;
  if (a < base || a + len < a) return bz_fail(BZ_ERR_WRITE, "address overflow");
  rc = arena_map(dc, a, a + len);
  if (rc) return rc;
  if (len) {
    CUDA_TRY(BZ_ERR_WRITE, cudaMemcpyAsync(dc->arena + a, data, len, cudaMemcpyHostToDevice, dc->stream));
    CUDA_TRY(BZ_ERR_WRITE, cudaStreamSynchronize(dc->stream));   // caller may free `data` on return
    arena_note_write(dc, a, a + len);
  }
  return BZ_OK;
}

extern "C" int32_t bz_dclient_dma_read(bz_dclient* dc, uint64_t base, uint64_t offset, uint8_t* out, size_t len) {
  int32_t rc = dc_select(dc);
  if (rc) return rc;
  if (!out && len) return bz_fail(BZ_ERR_READ, "null out");
  std::lock_guard<std::mutex> lk(dc->mu);
  uint64_t a = base + offset;
  if (a < base || a + len < a) return bz_fail(BZ_ERR_READ, "address overflow");
  if (a + len > ARENA_LIMIT) return bz_fail(BZ_ERR_READ, "address 0x%llx beyond the HBM window", (unsigned long long)(a + len));
  // never-written (unmapped) HBM reads back as zeros
  uint64_t pos = a;
  const uint64_t end = a + len;
  bool any = false;
  while (pos < end) {
    const uint64_t c = pos / dc->chunk;
    const uint64_t stop = std::min<uint64_t>(end, (c + 1) * dc->chunk);
    if (pos < dc->arena_va && dc->mapped[c]) {
      CUDA_TRY(BZ_ERR_READ, cudaMemcpyAsync(out + (pos - a), dc->arena + pos, stop - pos, cudaMemcpyDeviceToHost, dc->stream));
      any = true;
    } else {
      memset(out + (pos - a), 0, stop - pos);
    }
    pos = stop;
  }
  if (any) CUDA_TRY(BZ_ERR_READ, cudaStreamSynchronize(dc->stream));
  return BZ_OK;
}

extern "C" int32_t bz_dclient_firewalls_status(bz_dclient* dc, uint32_t* blocked_mask) {
  if (!dc) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient");
  if (blocked_mask) *blocked_mask = 0;
  return BZ_OK;
}
extern "C" int32_t bz_dclient_unblock_firewalls(bz_dclient* dc) { return dc ? BZ_OK : bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient"); }
extern "C" int32_t bz_dclient_initialize_cms(bz_dclient* dc) { return dc ? BZ_OK : bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient"); }
extern "C" int32_t bz_dclient_reset_sensor_data(bz_dclient* dc) { return dc ? BZ_OK : bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient"); }
extern "C" int32_t bz_dclient_setup_before_load_binary(bz_dclient* dc) { return dc ? BZ_OK : bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient"); }
extern "C" int32_t bz_dclient_load_binary(bz_dclient* dc, const uint8_t*, size_t) {
  // The kernels are part of this library (fatbin, sm_100a); there is no image to load.
  return dc ? BZ_OK : bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "null DriverClient");
}

extern "C" int32_t bz_dclient_device_info(bz_dclient* dc, char* name, size_t name_len, uint64_t* hbm_total, uint64_t* hbm_free) {
  int32_t rc = dc_select(dc);
  if (rc) return rc;
  cudaDeviceProp prop;
  CUDA_TRY(BZ_ERR_READ, cudaGetDeviceProperties(&prop, dc->device));
  if (name && name_len) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
  size_t f = 0, t = 0;
  CUDA_TRY(BZ_ERR_READ, cudaMemGetInfo(&f, &t));
  if (hbm_total) *hbm_total = t;
  if (hbm_free) *hbm_free = f;
  return BZ_OK;
}

extern "C" int32_t bz_host_alloc(size_t bytes, void** out) {
  if (!out) return bz_fail(BZ_ERR_INVALID_PRIMITIVE_PARAM, "out is null");
  CUDA_TRY(BZ_ERR_NO_DEVICE, cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return BZ_OK;
}
extern "C" int32_t bz_host_free(void* p) {
  if (p) CUDA_TRY(BZ_ERR_UNKNOWN, cudaFreeHost(p));
  return BZ_OK;
}
