// host.cu — the C ABI of libb200ddp.so (include/b200ddp.h): context, shared-arena bootstrap across the
// per-GPU processes (cudaIpc or VMM + POSIX fd over a unix socket, optional NVLS multicast), argument
// validation, and the launches of the kernels in kernels.cu.
//
// Replaces, for the one hot path of the reference (ref:multigpu.py:36,42), what torch's
// ProcessGroupNCCL communicator setup and c10d::Reducer's bucket copy/allreduce calls do; each entry
// point cites its counterpart in include/b200ddp.h.  No torch types, no exceptions across the ABI.
#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "internal.h"

using namespace b200ddp;

namespace {

// ---------------------------------------------------------------------------------------------------
// driver API, resolved at run time (no link-time libcuda dependency: the library must load on a box
// without a driver so the CPU-only checks can inspect its symbols)
// ---------------------------------------------------------------------------------------------------
struct Driver {
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) =
      nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) =
      nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  bool ok = false;
  bool has_mc = false;
};

template <typename F>
bool resolve(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess ||
      p == nullptr) {
    (void)cudaGetLastError();
    fn = nullptr;
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    bool ok = true;
    ok &= resolve("cuGetErrorString", d.GetErrorString);
    ok &= resolve("cuDeviceGet", d.DeviceGet);
    ok &= resolve("cuDeviceGetAttribute", d.DeviceGetAttribute);
    ok &= resolve("cuMemGetAllocationGranularity", d.MemGetAllocationGranularity);
    ok &= resolve("cuMemCreate", d.MemCreate);
    ok &= resolve("cuMemRelease", d.MemRelease);
    ok &= resolve("cuMemAddressReserve", d.MemAddressReserve);
    ok &= resolve("cuMemAddressFree", d.MemAddressFree);
    ok &= resolve("cuMemMap", d.MemMap);
    ok &= resolve("cuMemUnmap", d.MemUnmap);
    ok &= resolve("cuMemSetAccess", d.MemSetAccess);
    ok &= resolve("cuMemExportToShareableHandle", d.MemExportToShareableHandle);
    ok &= resolve("cuMemImportFromShareableHandle", d.MemImportFromShareableHandle);
    d.ok = ok;
    bool mc = true;
    mc &= resolve("cuMulticastCreate", d.MulticastCreate);
    mc &= resolve("cuMulticastAddDevice", d.MulticastAddDevice);
    mc &= resolve("cuMulticastBindMem", d.MulticastBindMem);
    mc &= resolve("cuMulticastUnbind", d.MulticastUnbind);
    mc &= resolve("cuMulticastGetGranularity", d.MulticastGetGranularity);
    d.has_mc = ok && mc;
  });
  return d;
}

const char* cu_err(CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString && driver().GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown CUresult";
}

thread_local std::string g_init_error;

// ---------------------------------------------------------------------------------------------------
// bootstrap handle (opaque to callers; B200DDP_HANDLE_BYTES)
// ---------------------------------------------------------------------------------------------------
struct Handle {
  uint32_t magic;
  int32_t rank;
  int32_t world;
  int32_t pid;
  int32_t device;
  int32_t mem_mode;
  uint64_t base;         // this rank's mapping of its own allocation (valid inside its process)
  uint64_t alloc_bytes;  // signal region + arena
  uint64_t nonce;        // names the rank's bootstrap socket
  cudaIpcMemHandle_t ipc;
  uint8_t pad[B200DDP_HANDLE_BYTES - 48 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(Handle) == B200DDP_HANDLE_BYTES, "handle layout");
constexpr uint32_t kMagic = 0xB200DD90u;

// messages on the bootstrap socket
struct Msg {
  uint32_t magic;
  int32_t type;  // 1 = arena fd, 2 = multicast fd, 3 = stage barrier
  int32_t src;
  int32_t stage;
  int32_t ok;
};
enum { kMsgMemFd = 1, kMsgMcFd = 2, kMsgStage = 3 };
struct Rx {
  Msg m;
  int fd;
};

}  // namespace

struct b200ddp_ctx {
  int rank = 0, world = 1, device = 0, mem_mode = B200DDP_MEM_IPC;
  uint64_t arena_bytes = 0, alloc_bytes = 0;
  bool opened = false;
  void* base = nullptr;  // own allocation
  PeerTable pt{};
  bool peer_is_ipc[kMaxWorld] = {};
  // VMM
  CUmemGenericAllocationHandle mem_handle = 0;
  CUmemGenericAllocationHandle peer_handle[kMaxWorld] = {};
  bool peer_is_vmm[kMaxWorld] = {};
  int share_fd = -1;
  // multicast
  CUmemGenericAllocationHandle mc_handle = 0;
  CUdeviceptr mc_va = 0;
  bool mc_bound = false, has_mc = false;
  // bootstrap socket
  int sock = -1;
  uint64_t nonce = 0;
  std::deque<Rx> pending;
  // runtime
  uint32_t* err_host = nullptr;
  uint32_t* err_dev = nullptr;
  cudaEvent_t ev = nullptr;
  // like a collective watchdog, not a latency bound: a peer may legitimately be late by minutes (rank 0 writing
  // a checkpoint while the others already wait in the next step's first bucket)
  uint32_t timeout_ms = 600000;
  std::atomic<uint64_t> launches{0};
  std::mutex mu;
  std::string err;
  int oneshot_max_bytes = 0;  // AUTO never picks the one-shot kernel unless B200DDP_ONESHOT_MAX says so
};

namespace {

int fail(b200ddp_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c)
    c->err = buf;
  else
    g_init_error = buf;
  return code;
}

#define RT_CHECK(c, call)                                                                              \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      (void)cudaGetLastError();                                                                        \
      return fail(c, B200DDP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                                           \
    }                                                                                                  \
  } while (0)

#define CU_CHECK(c, call)                                                                                \
  do {                                                                                                   \
    CUresult r__ = (call);                                                                               \
    if (r__ != CUDA_SUCCESS)                                                                             \
      return fail(c, B200DDP_ECUDA, "%s failed: %s (%s:%d)", #call, cu_err(r__), __FILE__, __LINE__);    \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
      changed = true;
    }
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};

uint64_t round_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

Ctl make_ctl(b200ddp_ctx* c) {
  Ctl ctl;
  ctl.err = c->err_dev;
  ctl.timeout_ns = uint64_t(c->timeout_ms) * 1000000ull;
  ctl.rank = c->rank;
  ctl.world = c->world;
  return ctl;
}

int dtype_size(int dtype) {
  switch (dtype) {
    case B200DDP_F32:
      return 4;
    case B200DDP_BF16:
    case B200DDP_F16:
      return 2;
    default:
      return 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// bootstrap socket: abstract-namespace unix datagram socket per context, named by (pid, nonce)
// ---------------------------------------------------------------------------------------------------
socklen_t sock_addr(sockaddr_un* a, int pid, uint64_t nonce) {
  memset(a, 0, sizeof *a);
  a->sun_family = AF_UNIX;
  int n = snprintf(a->sun_path + 1, sizeof(a->sun_path) - 1, "b200ddp-%d-%llx", pid, (unsigned long long)nonce);
  return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + n);
}

int sock_open(b200ddp_ctx* c) {
  c->sock = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (c->sock < 0) return fail(c, B200DDP_ESYS, "socket(): %s", strerror(errno));
  sockaddr_un a;
  socklen_t len = sock_addr(&a, getpid(), c->nonce);
  if (bind(c->sock, reinterpret_cast<sockaddr*>(&a), len) != 0)
    return fail(c, B200DDP_ESYS, "bind(bootstrap socket): %s", strerror(errno));
  timeval tv{120, 0};
  setsockopt(c->sock, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  return 0;
}

int sock_send(b200ddp_ctx* c, const Handle& to, const Msg& m, int fd) {
  sockaddr_un a;
  socklen_t len = sock_addr(&a, to.pid, to.nonce);
  msghdr mh{};
  iovec iov{const_cast<Msg*>(&m), sizeof m};
  mh.msg_name = &a;
  mh.msg_namelen = len;
  mh.msg_iov = &iov;
  mh.msg_iovlen = 1;
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  if (fd >= 0) {
    memset(ctrl, 0, sizeof ctrl);
    mh.msg_control = ctrl;
    mh.msg_controllen = sizeof ctrl;
    cmsghdr* cm = CMSG_FIRSTHDR(&mh);
    cm->cmsg_level = SOL_SOCKET;
    cm->cmsg_type = SCM_RIGHTS;
    cm->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  }
  for (int attempt = 0; attempt < 200; ++attempt) {
    if (sendmsg(c->sock, &mh, 0) >= 0) return 0;
    if (errno == EINTR || errno == EAGAIN || errno == ENOBUFS || errno == ECONNREFUSED) {
      usleep(10000);
      continue;
    }
    break;
  }
  return fail(c, B200DDP_ESYS, "sendmsg(to rank %d): %s", to.rank, strerror(errno));
}

// receive the next message of (type, stage); other messages are queued
int sock_recv(b200ddp_ctx* c, int type, int stage, Rx* out) {
  for (auto it = c->pending.begin(); it != c->pending.end(); ++it) {
    if (it->m.type == type && it->m.stage == stage) {
      *out = *it;
      c->pending.erase(it);
      return 0;
    }
  }
  for (;;) {
    Rx rx{};
    rx.fd = -1;
    msghdr mh{};
    iovec iov{&rx.m, sizeof rx.m};
    mh.msg_iov = &iov;
    mh.msg_iovlen = 1;
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    mh.msg_control = ctrl;
    mh.msg_controllen = sizeof ctrl;
    ssize_t n = recvmsg(c->sock, &mh, MSG_CMSG_CLOEXEC);
    if (n < 0) {
      if (errno == EINTR) continue;
      return fail(c, B200DDP_ESYS, "recvmsg(bootstrap, waiting type %d stage %d): %s", type, stage, strerror(errno));
    }
    if (n != static_cast<ssize_t>(sizeof(Msg)) || rx.m.magic != kMagic) continue;
    for (cmsghdr* cm = CMSG_FIRSTHDR(&mh); cm; cm = CMSG_NXTHDR(&mh, cm))
      if (cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS) memcpy(&rx.fd, CMSG_DATA(cm), sizeof(int));
    if (rx.m.type == type && rx.m.stage == stage) {
      *out = rx;
      return 0;
    }
    c->pending.push_back(rx);
  }
}

// all-to-all stage barrier over the sockets carrying an AND-reduced ok bit
int sock_stage(b200ddp_ctx* c, const Handle* hs, int stage, bool my_ok, bool* all_ok) {
  Msg m{kMagic, kMsgStage, c->rank, stage, my_ok ? 1 : 0};
  for (int q = 0; q < c->world; ++q)
    if (q != c->rank) {
      int rc = sock_send(c, hs[q], m, -1);
      if (rc) return rc;
    }
  bool ok = my_ok;
  for (int i = 0; i < c->world - 1; ++i) {
    Rx rx;
    int rc = sock_recv(c, kMsgStage, stage, &rx);
    if (rc) return rc;
    ok = ok && rx.m.ok;
  }
  *all_ok = ok;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// allocation
// ---------------------------------------------------------------------------------------------------
CUmemAllocationProp vmm_prop(int device) {
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int vmm_map(b200ddp_ctx* c, CUmemGenericAllocationHandle h, uint64_t bytes, uint64_t align, CUdeviceptr* va) {
  Driver& d = driver();
  CU_CHECK(c, d.MemAddressReserve(va, bytes, align, 0, 0));
  CUresult r = d.MemMap(*va, bytes, 0, h, 0);
  if (r != CUDA_SUCCESS) {
    d.MemAddressFree(*va, bytes);
    *va = 0;
    return fail(c, B200DDP_ECUDA, "cuMemMap failed: %s", cu_err(r));
  }
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = c->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.MemSetAccess(*va, bytes, &acc, 1);
  if (r != CUDA_SUCCESS) {
    d.MemUnmap(*va, bytes);
    d.MemAddressFree(*va, bytes);
    *va = 0;
    return fail(c, B200DDP_ECUDA, "cuMemSetAccess failed: %s", cu_err(r));
  }
  return 0;
}

uint64_t vmm_granularity(b200ddp_ctx* c) {
  Driver& d = driver();
  CUmemAllocationProp prop = vmm_prop(c->device);
  size_t gran = 2u << 20;
  if (d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS)
    gran = 2u << 20;
  if (d.has_mc && c->world > 1) {
    CUmulticastObjectProp mp{};
    mp.numDevices = static_cast<unsigned>(c->world);
    mp.size = gran;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (d.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) == CUDA_SUCCESS && mg > gran) gran = mg;
  }
  return gran;
}

int alloc_vmm(b200ddp_ctx* c) {
  Driver& d = driver();
  if (!d.ok) return fail(c, B200DDP_EUNSUP, "CUDA driver VMM entry points unavailable");
  const uint64_t gran = vmm_granularity(c);
  c->alloc_bytes = round_up(kSignalBytes + c->arena_bytes, gran);
  CUmemAllocationProp prop = vmm_prop(c->device);
  CUresult r = d.MemCreate(&c->mem_handle, c->alloc_bytes, &prop, 0);
  if (r != CUDA_SUCCESS) {
    c->mem_handle = 0;
    return fail(c, r == CUDA_ERROR_NOT_SUPPORTED || r == CUDA_ERROR_INVALID_VALUE ? B200DDP_EUNSUP : B200DDP_ECUDA,
                "cuMemCreate(%llu bytes, POSIX fd) failed: %s", (unsigned long long)c->alloc_bytes, cu_err(r));
  }
  CUdeviceptr va = 0;
  int rc = vmm_map(c, c->mem_handle, c->alloc_bytes, gran, &va);
  if (rc) {
    d.MemRelease(c->mem_handle);
    c->mem_handle = 0;
    return rc;
  }
  c->base = reinterpret_cast<void*>(va);
  int fd = -1;
  r = d.MemExportToShareableHandle(&fd, c->mem_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    d.MemUnmap(va, c->alloc_bytes);
    d.MemAddressFree(va, c->alloc_bytes);
    d.MemRelease(c->mem_handle);
    c->mem_handle = 0;
    c->base = nullptr;
    return fail(c, B200DDP_EUNSUP, "cuMemExportToShareableHandle failed: %s", cu_err(r));
  }
  c->share_fd = fd;
  c->mem_mode = B200DDP_MEM_VMM;
  return 0;
}

int alloc_ipc(b200ddp_ctx* c) {
  c->alloc_bytes = round_up(kSignalBytes + c->arena_bytes, 2u << 20);
  RT_CHECK(c, cudaMalloc(&c->base, c->alloc_bytes));
  c->mem_mode = B200DDP_MEM_IPC;
  return 0;
}

void free_local(b200ddp_ctx* c) {
  if (!c->base) return;
  if (c->mem_mode == B200DDP_MEM_VMM) {
    Driver& d = driver();
    d.MemUnmap(reinterpret_cast<CUdeviceptr>(c->base), c->alloc_bytes);
    d.MemAddressFree(reinterpret_cast<CUdeviceptr>(c->base), c->alloc_bytes);
    if (c->mem_handle) d.MemRelease(c->mem_handle);
    c->mem_handle = 0;
  } else {
    cudaFree(c->base);
  }
  c->base = nullptr;
}

// ---------------------------------------------------------------------------------------------------
// multicast (NVLS) — best effort: every rank ends with the same has_mc
// ---------------------------------------------------------------------------------------------------
int setup_multicast(b200ddp_ctx* c, const Handle* hs) {
  Driver& d = driver();
  bool ok = d.has_mc;
  int stage = 100;
  if (ok) {
    CUdevice dev;
    int sup = 0;
    ok = d.DeviceGet(&dev, c->device) == CUDA_SUCCESS &&
         d.DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS && sup != 0;
  }
  if (getenv("B200DDP_NO_MULTICAST")) ok = false;
  bool all = false;
  int rc = sock_stage(c, hs, stage++, ok, &all);
  if (rc) return rc;
  if (!all) return 0;  // someone cannot: nobody tries

  // rank 0 creates the multicast object and ships its fd
  CUmulticastObjectProp mp{};
  mp.numDevices = static_cast<unsigned>(c->world);
  mp.size = c->alloc_bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  std::string why;
  if (c->rank == 0) {
    int fd = -1;
    CUresult r = d.MulticastCreate(&c->mc_handle, &mp);
    if (r != CUDA_SUCCESS) {
      ok = false;
      why = std::string("cuMulticastCreate: ") + cu_err(r);
      c->mc_handle = 0;
    } else {
      r = d.MemExportToShareableHandle(&fd, c->mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
      if (r != CUDA_SUCCESS) {
        ok = false;
        why = std::string("export multicast fd: ") + cu_err(r);
      }
    }
    for (int q = 1; q < c->world; ++q) {
      Msg m{kMagic, kMsgMcFd, 0, 0, ok ? 1 : 0};
      rc = sock_send(c, hs[q], m, ok ? fd : -1);
      if (rc) return rc;
    }
    if (fd >= 0) close(fd);
  } else {
    Rx rx;
    rc = sock_recv(c, kMsgMcFd, 0, &rx);
    if (rc) return rc;
    ok = rx.m.ok && rx.fd >= 0;
    if (ok) {
      CUresult r = d.MemImportFromShareableHandle(&c->mc_handle, reinterpret_cast<void*>(static_cast<intptr_t>(rx.fd)),
                                                  CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      if (r != CUDA_SUCCESS) {
        ok = false;
        why = std::string("import multicast fd: ") + cu_err(r);
        c->mc_handle = 0;
      }
    }
    if (rx.fd >= 0) close(rx.fd);
  }
  if (ok) {
    CUdevice dev;
    d.DeviceGet(&dev, c->device);
    CUresult r = d.MulticastAddDevice(c->mc_handle, dev);
    if (r != CUDA_SUCCESS) {
      ok = false;
      why = std::string("cuMulticastAddDevice: ") + cu_err(r);
    }
  }
  rc = sock_stage(c, hs, stage++, ok, &all);  // every device added before anyone binds
  if (rc) return rc;
  if (all) {
    CUresult r = d.MulticastBindMem(c->mc_handle, 0, c->mem_handle, 0, c->alloc_bytes, 0);
    if (r != CUDA_SUCCESS) {
      ok = false;
      why = std::string("cuMulticastBindMem: ") + cu_err(r);
    } else {
      c->mc_bound = true;
      CUdeviceptr va = 0;
      if (vmm_map(c, c->mc_handle, c->alloc_bytes, vmm_granularity(c), &va) != 0) {
        ok = false;
        why = "map multicast: " + c->err;
      } else {
        c->mc_va = va;
      }
    }
  } else {
    ok = false;
  }
  rc = sock_stage(c, hs, stage++, ok, &all);  // every bind done before anyone uses the mapping
  if (rc) return rc;
  c->has_mc = all;
  if (!all && !why.empty()) c->err = "multicast disabled: " + why;
  return 0;
}

void teardown_multicast(b200ddp_ctx* c) {
  Driver& d = driver();
  if (c->mc_va) {
    d.MemUnmap(c->mc_va, c->alloc_bytes);
    d.MemAddressFree(c->mc_va, c->alloc_bytes);
    c->mc_va = 0;
  }
  if (c->mc_bound) {
    CUdevice dev;
    if (d.DeviceGet(&dev, c->device) == CUDA_SUCCESS) d.MulticastUnbind(c->mc_handle, dev, 0, c->alloc_bytes);
    c->mc_bound = false;
  }
  if (c->mc_handle) {
    d.MemRelease(c->mc_handle);
    c->mc_handle = 0;
  }
  c->has_mc = false;
}

int check_ready(b200ddp_ctx* c) {
  if (!c) return B200DDP_EINVAL;
  if (!c->opened) return fail(c, B200DDP_ESTATE, "b200ddp_open_peers has not completed on this context");
  return 0;
}

int check_range(b200ddp_ctx* c, const char* what, uint64_t off, uint64_t nbytes) {
  if ((off & 15) || (nbytes & 15)) return fail(c, B200DDP_EINVAL, "%s: offset/size must be multiples of 16 bytes", what);
  if (off > c->arena_bytes || nbytes > c->arena_bytes - off)
    return fail(c, B200DDP_EINVAL, "%s: [%llu, +%llu) exceeds the %llu-byte arena", what, (unsigned long long)off,
                (unsigned long long)nbytes, (unsigned long long)c->arena_bytes);
  return 0;
}

// Build the kernel-parameter tables for tensors [i0, i1) of a bucket; validates ordering/alignment.
int build_table(b200ddp_ctx* c, const void* const* ptrs, const uint64_t* off_el, const uint64_t* len_el, int i0, int i1,
                int esize, TensorTable* tt, uint32_t* vec_begin, uint32_t* vec_end) {
  const int epv = 16 / esize;
  tt->n = i1 - i0;
  for (int i = i0; i < i1; ++i) {
    const uint64_t ob = off_el[i] * esize;
    if (ob & 15)
      return fail(c, B200DDP_EINVAL, "tensor %d: bucket offset %llu elements is not 16-byte aligned", i,
                  (unsigned long long)off_el[i]);
    const uint64_t ov = ob / 16;
    const uint64_t nv = (len_el[i] + epv - 1) / epv;
    if (ov + nv > 0xffffffffull) return fail(c, B200DDP_EINVAL, "bucket larger than 64 GiB");
    if (i > i0 && ov < tt->off_vec[i - i0 - 1] + (len_el[i - 1] + epv - 1) / epv)
      return fail(c, B200DDP_EINVAL, "tensor %d overlaps its predecessor (offsets must ascend)", i);
    if (ptrs[i] == nullptr && len_el[i] != 0) return fail(c, B200DDP_EINVAL, "tensor %d: null pointer", i);
    if (reinterpret_cast<uint64_t>(ptrs[i]) % esize)
      return fail(c, B200DDP_EINVAL, "tensor %d: pointer not aligned to its element size", i);
    tt->ptr[i - i0] = reinterpret_cast<uint64_t>(ptrs[i]);
    tt->len_el[i - i0] = len_el[i];
    tt->off_vec[i - i0] = static_cast<uint32_t>(ov);
    if (i == i1 - 1) tt->off_vec[i - i0 + 1] = static_cast<uint32_t>(ov + nv);
  }
  *vec_begin = tt->off_vec[0];
  *vec_end = tt->off_vec[tt->n];
  return 0;
}

// parameter pointers of tensors [i0, i1) + the learning rate -> the kernels' optimizer-epilogue table
int build_sgd(b200ddp_ctx* c, void* const* param_ptrs, int i0, int i1, int esize, float lr, SgdTable* st) {
  if (!(lr == lr) || lr == 0.0f) return fail(c, B200DDP_EINVAL, "sgd: the learning rate must be a non-zero number");
  for (int i = i0; i < i1; ++i) {
    if (reinterpret_cast<uint64_t>(param_ptrs[i]) % esize)
      return fail(c, B200DDP_EINVAL, "parameter %d: pointer not aligned to its element size", i);
    st->ptr[i - i0] = reinterpret_cast<uint64_t>(param_ptrs[i]);  // NULL = this tensor has no parameter to update
  }
  st->neg_lr = -lr;
  return 0;
}

int copy_bucket(b200ddp_ctx* c, const void* const* ptrs, const uint64_t* off_el, const uint64_t* len_el, int n,
                int dtype, uint64_t bucket_off, float scale, bool unflatten, void* stream,
                void* const* param_ptrs = nullptr, float lr = 0.0f) {
  int rc = check_ready(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  const int esize = dtype_size(dtype);
  if (!esize) return fail(c, B200DDP_EINVAL, "unknown dtype %d", dtype);
  if (n < 0 || (n > 0 && (!ptrs || !off_el || !len_el))) return fail(c, B200DDP_EINVAL, "bad tensor table");
  if (n == 0) return 0;
  DeviceGuard g(c->device);
  unsigned char* bucket = static_cast<unsigned char*>(c->base) + kSignalBytes + bucket_off;
  for (int i0 = 0; i0 < n; i0 += kMaxTableTensors) {
    const int i1 = i0 + kMaxTableTensors < n ? i0 + kMaxTableTensors : n;
    TensorTable tt;
    uint32_t vb, ve;
    rc = build_table(c, ptrs, off_el, len_el, i0, i1, esize, &tt, &vb, &ve);
    if (rc) return rc;
    rc = check_range(c, unflatten ? "unflatten" : "flatten_scale", bucket_off, uint64_t(ve) * 16);
    if (rc) return rc;
    // an unflatten after a peer timeout must not scatter a partial sum: it checks the rank's sticky poison word
    const uint32_t* poison = (unflatten && c->world > 1)
                                 ? reinterpret_cast<const uint32_t*>(static_cast<unsigned char*>(c->base) + kFlagBytes + kEpochBytes)
                                 : nullptr;
    SgdTable st;
    if (param_ptrs && (rc = build_sgd(c, param_ptrs, i0, i1, esize, lr, &st))) return rc;
    RT_CHECK(c, launch_flatten(tt, bucket, vb, ve, dtype, scale, unflatten, poison, static_cast<cudaStream_t>(stream),
                               param_ptrs ? &st : nullptr));
    c->launches++;
  }
  return 0;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  if (!s || !*s) return dflt;
  return atoi(s);
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int b200ddp_abi_version(void) { return B200DDP_ABI_VERSION; }
int b200ddp_max_table_tensors(void) { return kMaxTableTensors; }
uint64_t b200ddp_fused_region_bytes(b200ddp_ctx* c, uint64_t bucket_bytes) {
  return c ? fused_region_bytes((bucket_bytes + 15) / 16 * 16, c->world) : 0;
}
uint64_t b200ddp_fused_max_bytes(void) { return kFusedMaxBytes; }
uint64_t b200ddp_fused_region_bytes_for(int world, uint64_t bucket_bytes) {
  return fused_region_bytes((bucket_bytes + 15) / 16 * 16, world);
}

const char* b200ddp_last_error(b200ddp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int b200ddp_init(int rank, int world, int device, uint64_t arena_bytes, int mem_mode, b200ddp_ctx** out) {
  if (!out) return fail(nullptr, B200DDP_EINVAL, "out is null");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
    return fail(nullptr, B200DDP_EINVAL, "rank %d / world %d out of range (max world %d)", rank, world, kMaxWorld);
  if (mem_mode < B200DDP_MEM_AUTO || mem_mode > B200DDP_MEM_VMM)
    return fail(nullptr, B200DDP_EINVAL, "unknown mem_mode %d", mem_mode);
  if (arena_bytes == 0 || arena_bytes > (uint64_t(1) << 40)) return fail(nullptr, B200DDP_EINVAL, "bad arena size");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(nullptr, B200DDP_ECUDA, "no CUDA device available (%s): this library has no CPU fallback",
                ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  }
  if (device < 0 || device >= ndev) return fail(nullptr, B200DDP_EINVAL, "device %d out of range (%d present)", device, ndev);

  b200ddp_ctx* c = new b200ddp_ctx();
  c->rank = rank;
  c->world = world;
  c->device = device;
  c->arena_bytes = round_up(arena_bytes, 16);
  c->oneshot_max_bytes = env_int("B200DDP_ONESHOT_MAX", c->oneshot_max_bytes);
  timeval tv;
  gettimeofday(&tv, nullptr);
  static std::atomic<uint64_t> seq{0};
  c->nonce = (uint64_t(tv.tv_sec) << 20) ^ uint64_t(tv.tv_usec) ^ (seq.fetch_add(1) << 40) ^ (uint64_t(rank) << 56);

  auto bail = [&](int code) {
    g_init_error = c->err;
    if (c->ev) cudaEventDestroy(c->ev);
    if (c->err_host) cudaFreeHost(c->err_host);
    free_local(c);
    if (c->share_fd >= 0) close(c->share_fd);
    if (c->sock >= 0) close(c->sock);
    delete c;
    return code;
  };

  DeviceGuard g(device);
  ce = cudaFree(nullptr);  // create / bind the primary context on this thread
  if (ce != cudaSuccess) {
    fail(c, B200DDP_ECUDA, "cannot initialise device %d: %s", device, cudaGetErrorString(ce));
    return bail(B200DDP_ECUDA);
  }
  int cc_major = 0;
  cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device);
  if (cc_major != 10) {
    fail(c, B200DDP_EUNSUP, "device %d has compute capability %d.x; this library is built for sm_100a only", device,
         cc_major);
    return bail(B200DDP_EUNSUP);
  }
  int rc = 0;
  if (mem_mode == B200DDP_MEM_IPC) {
    rc = alloc_ipc(c);
  } else {
    rc = alloc_vmm(c);
    if (rc == B200DDP_EUNSUP && mem_mode == B200DDP_MEM_AUTO) rc = alloc_ipc(c);
  }
  if (rc) return bail(rc);
  if (cudaMemset(c->base, 0, c->alloc_bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    fail(c, B200DDP_ECUDA, "cannot clear the arena: %s", cudaGetErrorString(cudaGetLastError()));
    return bail(B200DDP_ECUDA);
  }
  if (cudaHostAlloc(reinterpret_cast<void**>(&c->err_host), 64, cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->err_dev), c->err_host, 0) != cudaSuccess) {
    fail(c, B200DDP_ECUDA, "cannot allocate the mapped error word: %s", cudaGetErrorString(cudaGetLastError()));
    return bail(B200DDP_ECUDA);
  }
  *c->err_host = 0;
  if (cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming) != cudaSuccess ||
      prepare_kernels(world) != cudaSuccess) {
    fail(c, B200DDP_ECUDA, "kernel/event setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    return bail(B200DDP_ECUDA);
  }
  if (c->mem_mode == B200DDP_MEM_VMM) {
    rc = sock_open(c);
    if (rc) return bail(rc);
  }
  c->pt.base[rank] = reinterpret_cast<uint64_t>(c->base);
  if (world == 1) c->opened = true;
  *out = c;
  return 0;
}

int b200ddp_export_handle(b200ddp_ctx* c, void* out, uint64_t cap) {
  if (!c || !out) return B200DDP_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (cap < sizeof(Handle)) return fail(c, B200DDP_EINVAL, "handle buffer too small");
  Handle h{};
  h.magic = kMagic;
  h.rank = c->rank;
  h.world = c->world;
  h.pid = static_cast<int32_t>(getpid());
  h.device = c->device;
  h.mem_mode = c->mem_mode;
  h.base = reinterpret_cast<uint64_t>(c->base);
  h.alloc_bytes = c->alloc_bytes;
  h.nonce = c->nonce;
  if (c->mem_mode == B200DDP_MEM_IPC) {
    DeviceGuard g(c->device);
    RT_CHECK(c, cudaIpcGetMemHandle(&h.ipc, c->base));
  }
  memcpy(out, &h, sizeof h);
  return 0;
}

int b200ddp_open_peers(b200ddp_ctx* c, const void* all_handles, int n) {
  if (!c || !all_handles) return B200DDP_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->opened && c->world > 1) return fail(c, B200DDP_ESTATE, "peers already opened");
  if (n != c->world) return fail(c, B200DDP_EINVAL, "got %d handles for world %d", n, c->world);
  if (c->world == 1) return 0;
  DeviceGuard g(c->device);
  std::vector<Handle> hs(n);
  memcpy(hs.data(), all_handles, sizeof(Handle) * n);
  const int mypid = static_cast<int>(getpid());
  bool any_remote = false;
  for (int q = 0; q < n; ++q) {
    const Handle& h = hs[q];
    if (h.magic != kMagic || h.rank != q || h.world != c->world)
      return fail(c, B200DDP_EINVAL, "handle %d is corrupt (magic %x rank %d world %d)", q, h.magic, h.rank, h.world);
    if (h.mem_mode != c->mem_mode)
      return fail(c, B200DDP_EINVAL, "rank %d uses mem_mode %d but this rank uses %d", q, h.mem_mode, c->mem_mode);
    if (h.alloc_bytes != c->alloc_bytes)
      return fail(c, B200DDP_EINVAL, "rank %d arena size differs (%llu vs %llu)", q, (unsigned long long)h.alloc_bytes,
                  (unsigned long long)c->alloc_bytes);
    if (q != c->rank && h.pid != mypid) any_remote = true;
  }
  // 1. peers living in this process (loop-back worlds): their mapping is directly usable
  for (int q = 0; q < n; ++q) {
    if (q == c->rank || hs[q].pid != mypid) continue;
    if (hs[q].device != c->device) {
      if (c->mem_mode != B200DDP_MEM_IPC)
        return fail(c, B200DDP_EUNSUP, "same-process peers on different devices need mem_mode IPC");
      cudaError_t e = cudaDeviceEnablePeerAccess(hs[q].device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        return fail(c, B200DDP_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", hs[q].device, cudaGetErrorString(e));
      }
      (void)cudaGetLastError();
    }
    c->pt.base[q] = hs[q].base;
  }
  // 2. peers in other processes
  if (any_remote) {
    if (c->mem_mode == B200DDP_MEM_IPC) {
      for (int q = 0; q < n; ++q) {
        if (q == c->rank || hs[q].pid == mypid) continue;
        void* p = nullptr;
        RT_CHECK(c, cudaIpcOpenMemHandle(&p, hs[q].ipc, cudaIpcMemLazyEnablePeerAccess));
        c->pt.base[q] = reinterpret_cast<uint64_t>(p);
        c->peer_is_ipc[q] = true;
      }
    } else {
      Driver& d = driver();
      for (int q = 0; q < n; ++q) {
        if (q == c->rank || hs[q].pid == mypid) continue;
        Msg m{kMagic, kMsgMemFd, c->rank, 0, 1};
        int rc = sock_send(c, hs[q], m, c->share_fd);
        if (rc) return rc;
      }
      int expect = 0;
      for (int q = 0; q < n; ++q)
        if (q != c->rank && hs[q].pid != mypid) ++expect;
      for (int i = 0; i < expect; ++i) {
        Rx rx;
        int rc = sock_recv(c, kMsgMemFd, 0, &rx);
        if (rc) return rc;
        const int q = rx.m.src;
        if (q < 0 || q >= n || q == c->rank || rx.fd < 0 || c->pt.base[q] != 0) {
          if (rx.fd >= 0) close(rx.fd);
          return fail(c, B200DDP_ESYS, "unexpected arena fd message from rank %d", q);
        }
        CUmemGenericAllocationHandle h = 0;
        CUresult r = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<intptr_t>(rx.fd)),
                                                    CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(rx.fd);
        if (r != CUDA_SUCCESS) return fail(c, B200DDP_ECUDA, "import arena of rank %d: %s", q, cu_err(r));
        CUdeviceptr va = 0;
        int rc2 = vmm_map(c, h, c->alloc_bytes, vmm_granularity(c), &va);
        if (rc2) {
          d.MemRelease(h);
          return rc2;
        }
        c->peer_handle[q] = h;
        c->peer_is_vmm[q] = true;
        c->pt.base[q] = static_cast<uint64_t>(va);
      }
      bool all_remote = true;
      for (int q = 0; q < n; ++q)
        if (q != c->rank && hs[q].pid == mypid) all_remote = false;
      if (all_remote) {
        int rc = setup_multicast(c, hs.data());
        if (rc) return rc;
      }
    }
  }
  for (int q = 0; q < n; ++q)
    if (c->pt.base[q] == 0) return fail(c, B200DDP_ESTATE, "rank %d was not mapped", q);
  c->opened = true;
  return 0;
}

int b200ddp_arena_ptr(b200ddp_ctx* c, void** dev_ptr) {
  if (!c || !dev_ptr) return B200DDP_EINVAL;
  *dev_ptr = static_cast<unsigned char*>(c->base) + kSignalBytes;
  return 0;
}
uint64_t b200ddp_arena_bytes(b200ddp_ctx* c) { return c ? c->arena_bytes : 0; }
int b200ddp_has_multicast(b200ddp_ctx* c) { return c && c->has_mc ? 1 : 0; }
int b200ddp_mem_mode(b200ddp_ctx* c) { return c ? c->mem_mode : B200DDP_EINVAL; }

int b200ddp_destroy(b200ddp_ctx* c) {
  if (!c) return B200DDP_EINVAL;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    (void)cudaGetLastError();
    teardown_multicast(c);
    Driver& d = driver();
    for (int q = 0; q < c->world; ++q) {
      if (q == c->rank) continue;
      if (c->peer_is_ipc[q]) cudaIpcCloseMemHandle(reinterpret_cast<void*>(c->pt.base[q]));
      if (c->peer_is_vmm[q]) {
        d.MemUnmap(static_cast<CUdeviceptr>(c->pt.base[q]), c->alloc_bytes);
        d.MemAddressFree(static_cast<CUdeviceptr>(c->pt.base[q]), c->alloc_bytes);
        d.MemRelease(c->peer_handle[q]);
      }
    }
    free_local(c);
    if (c->share_fd >= 0) close(c->share_fd);
    if (c->sock >= 0) close(c->sock);
    for (auto& rx : c->pending)
      if (rx.fd >= 0) close(rx.fd);
    if (c->ev) cudaEventDestroy(c->ev);
    if (c->err_host) cudaFreeHost(c->err_host);
    (void)cudaGetLastError();
  }
  delete c;
  return 0;
}

int b200ddp_flatten_scale(b200ddp_ctx* c, const void* const* grad_ptrs, const uint64_t* off_el,
                          const uint64_t* len_el, int n, int dtype, uint64_t bucket_off, float scale, void* stream) {
  return copy_bucket(c, grad_ptrs, off_el, len_el, n, dtype, bucket_off, scale, false, stream);
}

int b200ddp_unflatten(b200ddp_ctx* c, void* const* grad_ptrs, const uint64_t* off_el, const uint64_t* len_el, int n,
                      int dtype, uint64_t bucket_off, void* stream) {
  return copy_bucket(c, grad_ptrs, off_el, len_el, n, dtype, bucket_off, 1.0f, true, stream);
}

int b200ddp_unflatten_sgd(b200ddp_ctx* c, void* const* grad_ptrs, void* const* param_ptrs, const uint64_t* off_el,
                          const uint64_t* len_el, int n, int dtype, uint64_t bucket_off, float lr, void* stream) {
  if (c && !param_ptrs) return fail(c, B200DDP_EINVAL, "unflatten_sgd: param_ptrs is null");
  return copy_bucket(c, grad_ptrs, off_el, len_el, n, dtype, bucket_off, 1.0f, true, stream, param_ptrs, lr);
}

int b200ddp_pick_algo(b200ddp_ctx* c, uint64_t nbytes) {
  if (!c) return B200DDP_EINVAL;
  const int forced = env_int("B200DDP_ALGO", 0);
  if (forced >= B200DDP_ALGO_ONESHOT && forced <= B200DDP_ALGO_NVLS) {
    if (forced == B200DDP_ALGO_NVLS && !c->has_mc) return B200DDP_ALGO_TWOSHOT;
    return forced;
  }
  if (nbytes <= static_cast<uint64_t>(c->oneshot_max_bytes)) return B200DDP_ALGO_ONESHOT;
  // NVLS moves (1 + 1/ws) S per GPU and direction, the peer-to-peer two-shot 2 (ws-1)/ws S: the switch
  // reduction only pays from ws >= 4 (measured at ws = 2: 393 vs 642 GB/s bus at 256 MiB)
  if (c->has_mc && c->world >= env_int("B200DDP_NVLS_MIN_WORLD", 4)) return B200DDP_ALGO_NVLS;
  return B200DDP_ALGO_TWOSHOT_TMA;
}

int b200ddp_allreduce(b200ddp_ctx* c, uint64_t in_off, uint64_t out_off, uint64_t nbytes, int dtype, int algo,
                      int max_ctas, float scale, void* stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  if (algo == B200DDP_ALGO_AUTO) algo = b200ddp_pick_algo(c, nbytes);
  std::lock_guard<std::mutex> lk(c->mu);
  if (!dtype_size(dtype)) return fail(c, B200DDP_EINVAL, "unknown dtype %d", dtype);
  if ((rc = check_range(c, "allreduce(in)", in_off, nbytes))) return rc;
  if ((rc = check_range(c, "allreduce(out)", out_off, nbytes))) return rc;
  if (nbytes == 0) return 0;
  if (in_off != out_off && in_off < out_off + nbytes && out_off < in_off + nbytes)
    return fail(c, B200DDP_EINVAL, "allreduce: in and out regions overlap partially");
  DeviceGuard g(c->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint64_t nvec = nbytes / 16;
  unsigned char* arena = static_cast<unsigned char*>(c->base) + kSignalBytes;
  if (!(scale == scale) || scale == 0.0f) return fail(c, B200DDP_EINVAL, "allreduce: scale must be a non-zero number");
  if (c->world == 1) {
    if (scale != 1.0f) {
      // a world of one still owes the caller out = scale * in: one flatten launch over the region as one tensor
      TensorTable tt{};
      tt.n = 1;
      tt.ptr[0] = reinterpret_cast<uint64_t>(arena + in_off);
      tt.len_el[0] = nbytes / dtype_size(dtype);
      tt.off_vec[0] = 0;
      tt.off_vec[1] = static_cast<uint32_t>(nvec);
      if (nvec > 0xffffffffull) return fail(c, B200DDP_EINVAL, "bucket larger than 64 GiB");
      RT_CHECK(c, launch_flatten(tt, arena + out_off, 0, static_cast<uint32_t>(nvec), dtype, scale, false, nullptr, s));
      c->launches++;
    } else if (in_off != out_off) {
      RT_CHECK(c, launch_local_copy(arena + out_off, arena + in_off, nvec, 0, s));
      c->launches++;
    }
    return 0;
  }
  const Ctl ctl = make_ctl(c);
  switch (algo) {
    case B200DDP_ALGO_ONESHOT:
      if (in_off == out_off) return fail(c, B200DDP_EINVAL, "one-shot allreduce must be out of place");
      RT_CHECK(c, launch_oneshot(c->pt, ctl, in_off, out_off, nvec, dtype, max_ctas, scale, s));
      break;
    case B200DDP_ALGO_TWOSHOT:
      RT_CHECK(c, launch_twoshot(c->pt, ctl, in_off, out_off, nvec, dtype, max_ctas, scale, s));
      break;
    case B200DDP_ALGO_TWOSHOT_TMA:
      RT_CHECK(c, launch_twoshot_tma(c->pt, ctl, in_off, out_off, nvec, dtype, max_ctas, scale, s));
      break;
    case B200DDP_ALGO_NVLS:
      if (!c->has_mc) return fail(c, B200DDP_EUNSUP, "NVLS requested but no multicast mapping exists");
      RT_CHECK(c, launch_nvls(c->pt, ctl, static_cast<uint64_t>(c->mc_va), in_off, out_off, nvec, dtype, max_ctas, scale,
                              s));
      break;
    default:
      return fail(c, B200DDP_EINVAL, "unknown algorithm %d", algo);
  }
  c->launches++;
  return 0;
}

namespace {
int fused_bucket_impl(b200ddp_ctx* c, void* const* grad_ptrs, void* const* param_ptrs, const uint64_t* off_el,
                      const uint64_t* len_el, int n, int dtype, uint64_t bucket_off, uint64_t bucket_bytes, float scale,
                      float lr, void* stream);
}

int b200ddp_fused_bucket(b200ddp_ctx* c, void* const* grad_ptrs, const uint64_t* off_el, const uint64_t* len_el,
                         int n, int dtype, uint64_t bucket_off, uint64_t bucket_bytes, float scale, void* stream) {
  return fused_bucket_impl(c, grad_ptrs, nullptr, off_el, len_el, n, dtype, bucket_off, bucket_bytes, scale, 0.0f, stream);
}

int b200ddp_fused_bucket_sgd(b200ddp_ctx* c, void* const* grad_ptrs, void* const* param_ptrs, const uint64_t* off_el,
                             const uint64_t* len_el, int n, int dtype, uint64_t bucket_off, uint64_t bucket_bytes,
                             float scale, float lr, void* stream) {
  if (c && !param_ptrs) return fail(c, B200DDP_EINVAL, "fused_bucket_sgd: param_ptrs is null");
  return fused_bucket_impl(c, grad_ptrs, param_ptrs, off_el, len_el, n, dtype, bucket_off, bucket_bytes, scale, lr, stream);
}

}  // extern "C"

namespace {
int fused_bucket_impl(b200ddp_ctx* c, void* const* grad_ptrs, void* const* param_ptrs, const uint64_t* off_el,
                      const uint64_t* len_el, int n, int dtype, uint64_t bucket_off, uint64_t bucket_bytes, float scale,
                      float lr, void* stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  const int esize = dtype_size(dtype);
  if (!esize) return fail(c, B200DDP_EINVAL, "unknown dtype %d", dtype);
  if (n <= 0 || n > kMaxTableTensors || !grad_ptrs || !off_el || !len_el)
    return fail(c, B200DDP_EINVAL, "fused_bucket takes 1..%d tensors (got %d)", kMaxTableTensors, n);
  if (bucket_bytes > kFusedMaxBytes)
    return fail(c, B200DDP_EINVAL, "fused_bucket takes at most %llu bytes", (unsigned long long)kFusedMaxBytes);
  // world > 1: the staging is double buffered (two slots, or the LL protocol's slot sets)
  if ((rc = check_range(c, "fused_bucket", bucket_off, fused_region_bytes(bucket_bytes, c->world)))) return rc;
  TensorTable tt;
  uint32_t vb, ve;
  rc = build_table(c, const_cast<const void* const*>(grad_ptrs), off_el, len_el, 0, n, esize, &tt, &vb, &ve);
  if (rc) return rc;
  if (uint64_t(ve) * 16 > bucket_bytes) return fail(c, B200DDP_EINVAL, "tensor table exceeds bucket_bytes");
  SgdTable st;
  if (param_ptrs && (rc = build_sgd(c, param_ptrs, 0, n, esize, lr, &st))) return rc;
  DeviceGuard g(c->device);
  RT_CHECK(c, launch_fused(tt, c->pt, make_ctl(c), bucket_off, static_cast<uint32_t>(bucket_bytes / 16), dtype, scale,
                           static_cast<cudaStream_t>(stream), param_ptrs ? &st : nullptr));
  c->launches++;
  return 0;
}
}  // namespace

extern "C" {

int b200ddp_broadcast(b200ddp_ctx* c, uint64_t off, uint64_t nbytes, int root, int max_ctas, void* stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  if (root < 0 || root >= c->world) return fail(c, B200DDP_EINVAL, "broadcast root %d out of range", root);
  if ((rc = check_range(c, "broadcast", off, nbytes))) return rc;
  if (c->world == 1 || nbytes == 0) return 0;
  DeviceGuard g(c->device);
  RT_CHECK(c, launch_broadcast(c->pt, make_ctl(c), off, nbytes / 16, root, max_ctas, static_cast<cudaStream_t>(stream)));
  c->launches++;
  return 0;
}

int b200ddp_peer_read(b200ddp_ctx* c, int peer, uint64_t src_off, uint64_t dst_off, uint64_t nbytes, int max_ctas,
                      void* stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  if (peer < 0 || peer >= c->world) return fail(c, B200DDP_EINVAL, "peer_read: rank %d out of range", peer);
  if ((rc = check_range(c, "peer_read(src)", src_off, nbytes))) return rc;
  if ((rc = check_range(c, "peer_read(dst)", dst_off, nbytes))) return rc;
  if (nbytes == 0) return 0;
  DeviceGuard g(c->device);
  const unsigned char* src = reinterpret_cast<const unsigned char*>(c->pt.base[peer]) + kSignalBytes + src_off;
  unsigned char* dst = static_cast<unsigned char*>(c->base) + kSignalBytes + dst_off;
  RT_CHECK(c, launch_local_copy(dst, src, nbytes / 16, max_ctas, static_cast<cudaStream_t>(stream)));
  c->launches++;
  return 0;
}

int b200ddp_launch_graph(void* graph_exec, void* stream, void* done_event) {
  if (!graph_exec) return fail(nullptr, B200DDP_EINVAL, "launch_graph: null graph");
  cudaError_t e = cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), static_cast<cudaStream_t>(stream));
  if (e == cudaSuccess && done_event)
    e = cudaEventRecord(static_cast<cudaEvent_t>(done_event), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(nullptr, B200DDP_ECUDA, "launch_graph failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

int b200ddp_gather_rows(const void* src, void* dst, void* dst2, uint64_t row_bytes, uint64_t split_bytes,
                        uint32_t batch_rows, uint64_t full_rows, void* cursor, void* stream) {
  if (!src || !dst || !cursor || (split_bytes != 0 && !dst2))
    return fail(nullptr, B200DDP_EINVAL, "gather_rows: null pointer");
  if (row_bytes == 0 || (row_bytes & 3) || (split_bytes & 3) || split_bytes >= row_bytes || batch_rows == 0 ||
      full_rows == 0 || row_bytes * batch_rows > (64u << 20) || (reinterpret_cast<uint64_t>(src) & 3) ||
      (reinterpret_cast<uint64_t>(dst) & 3) || (reinterpret_cast<uint64_t>(dst2) & 3) ||
      (reinterpret_cast<uint64_t>(cursor) & 7))
    return fail(nullptr, B200DDP_EINVAL,
                "gather_rows: rows and the split must be whole 4-byte words (split < row), batch <= 64 MiB, aligned pointers");
  cudaError_t e = launch_gather_rows(src, dst, dst2, static_cast<uint32_t>(row_bytes / 4),
                                     static_cast<uint32_t>(split_bytes / 4), batch_rows, full_rows, cursor,
                                     static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(nullptr, B200DDP_ECUDA, "gather_rows launch failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

int b200ddp_clear(b200ddp_ctx* c, uint64_t off, uint64_t nbytes, void* stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  if ((rc = check_range(c, "clear", off, nbytes))) return rc;
  if (nbytes == 0) return 0;
  DeviceGuard g(c->device);
  RT_CHECK(c, cudaMemsetAsync(static_cast<unsigned char*>(c->base) + kSignalBytes + off, 0, nbytes,
                              static_cast<cudaStream_t>(stream)));
  return 0;
}

int b200ddp_barrier(b200ddp_ctx* c, void* stream) {
  int rc = check_ready(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->world == 1) return 0;
  DeviceGuard g(c->device);
  RT_CHECK(c, launch_barrier(c->pt, make_ctl(c), static_cast<cudaStream_t>(stream)));
  c->launches++;
  return 0;
}

int b200ddp_stream_wait(b200ddp_ctx* c, void* waiter, void* signaled) {
  if (!c) return B200DDP_EINVAL;
  std::lock_guard<std::mutex> lk(c->mu);
  if (waiter == signaled) return 0;
  DeviceGuard g(c->device);
  RT_CHECK(c, cudaEventRecord(c->ev, static_cast<cudaStream_t>(signaled)));
  RT_CHECK(c, cudaStreamWaitEvent(static_cast<cudaStream_t>(waiter), c->ev, 0));
  return 0;
}

int b200ddp_poll_error(b200ddp_ctx* c) {
  if (!c) return B200DDP_EINVAL;
  const uint32_t v = *reinterpret_cast<volatile uint32_t*>(c->err_host);
  if (v == 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  *reinterpret_cast<volatile uint32_t*>(c->err_host) = 0;
  return fail(c, B200DDP_ETIMEOUT, "rank %d: a kernel timed out after %u ms waiting for rank %u at an in-kernel barrier",
              c->rank, c->timeout_ms, v - 1);
}

int b200ddp_set_timeout_ms(b200ddp_ctx* c, uint32_t ms) {
  if (!c || ms == 0) return B200DDP_EINVAL;
  c->timeout_ms = ms;
  return 0;
}

uint64_t b200ddp_launch_count(b200ddp_ctx* c) { return c ? c->launches.load() : 0; }

}  // extern "C"
