// Symmetric-heap memory through the CUDA virtual-memory-management driver API, owned by this library:
//
//   arena        cuMemCreate (pinned device memory, exportable as a POSIX file descriptor) -> cuMemAddressReserve / cuMemMap /
//                cuMemSetAccess; every peer imports the descriptor (cuMemImportFromShareableHandle) and maps it the same way,
//                so the fused collectives address any peer's HBM with plain loads / stores over NVLink 5 / NVSwitch
//   multicast    rank 0 creates the NVLS multicast object (cuMulticastCreate) and exports it; every rank adds its device
//                (cuMulticastAddDevice), binds its arena (cuMulticastBindMem) and maps the object: a store to that mapping
//                lands in all arenas, a multimem.ld_reduce is summed inside the switch (collectives.cu: kallreduce_mc, kps_sync)
//
// Descriptors travel between the ranks of a node over abstract AF_UNIX sockets (SCM_RIGHTS), done in
// parallel/symm.py:_create_vmm_backend with the standard library - nothing here depends on torch.distributed._symmetric_memory.
//
// Reference analogue: the pinned `comm_buf_` staging buffers of KVStoreDist (kvstore_dist.h:335-342), the P2P enablement of
// CommDevice (comm.h:732-760: cudaDeviceEnablePeerAccess between every pair) and the EFS mount every instance shares
// (deeplearning.template:453-474) - here one mapped address range per peer.
//
// dlcfn_probe_fabric is the native topology probe (reference: gpu_topology.h:133-196, cudaDeviceGetP2PAttribute per pair).
#include "common.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <unistd.h>

namespace {

template <typename Fn>
Fn driver_fn(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return (Fn)f;
}

struct Drv {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*AddrReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*AddrFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*Granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*McCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*McAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*McBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*McGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*DevGetAttr)(int*, CUdevice_attribute, CUdevice);
  CUresult (*DevGet)(CUdevice*, int);
  bool ok = false, mc_ok = false;
};

Drv& drv() {
  static Drv d = [] {
    Drv x;
    memset(&x, 0, sizeof(x));
    (void)cudaFree(nullptr);      // the driver entry points need an initialised runtime
    x.MemCreate = driver_fn<decltype(x.MemCreate)>("cuMemCreate");
    x.MemRelease = driver_fn<decltype(x.MemRelease)>("cuMemRelease");
    x.MemExport = driver_fn<decltype(x.MemExport)>("cuMemExportToShareableHandle");
    x.MemImport = driver_fn<decltype(x.MemImport)>("cuMemImportFromShareableHandle");
    x.AddrReserve = driver_fn<decltype(x.AddrReserve)>("cuMemAddressReserve");
    x.AddrFree = driver_fn<decltype(x.AddrFree)>("cuMemAddressFree");
    x.MemMap = driver_fn<decltype(x.MemMap)>("cuMemMap");
    x.MemUnmap = driver_fn<decltype(x.MemUnmap)>("cuMemUnmap");
    x.SetAccess = driver_fn<decltype(x.SetAccess)>("cuMemSetAccess");
    x.Granularity = driver_fn<decltype(x.Granularity)>("cuMemGetAllocationGranularity");
    x.McCreate = driver_fn<decltype(x.McCreate)>("cuMulticastCreate");
    x.McAddDevice = driver_fn<decltype(x.McAddDevice)>("cuMulticastAddDevice");
    x.McBindMem = driver_fn<decltype(x.McBindMem)>("cuMulticastBindMem");
    x.McGranularity = driver_fn<decltype(x.McGranularity)>("cuMulticastGetGranularity");
    x.DevGetAttr = driver_fn<decltype(x.DevGetAttr)>("cuDeviceGetAttribute");
    x.DevGet = driver_fn<decltype(x.DevGet)>("cuDeviceGet");
    x.ok = x.MemCreate && x.MemRelease && x.MemExport && x.MemImport && x.AddrReserve && x.AddrFree && x.MemMap && x.MemUnmap &&
           x.SetAccess && x.Granularity && x.DevGetAttr && x.DevGet;
    x.mc_ok = x.ok && x.McCreate && x.McAddDevice && x.McBindMem && x.McGranularity;
    return x;
  }();
  return d;
}

CUmemAllocationProp arena_prop(int dev) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = dev;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

#define DRV_CHECK(call)                                                                               \
  do {                                                                                                \
    CUresult _r = (call);                                                                             \
    if (_r != CUDA_SUCCESS) {                                                                         \
      fprintf(stderr, "[dlcfn vmm] %s failed with CUresult %d (%s:%d)\n", #call, (int)_r, __FILE__, __LINE__); \
      return 20000 + (int)_r;                                                                         \
    }                                                                                                 \
  } while (0)

// map `handle` (size bytes) into this process with read / write access for `dev`
int map_handle(CUmemGenericAllocationHandle handle, size_t bytes, size_t gran, int dev, void** out_ptr) {
  Drv& d = drv();
  CUdeviceptr va = 0;
  DRV_CHECK(d.AddrReserve(&va, bytes, gran, 0, 0));
  CUresult r = d.MemMap(va, bytes, 0, handle, 0);
  if (r != CUDA_SUCCESS) {
    d.AddrFree(va, bytes);
    fprintf(stderr, "[dlcfn vmm] cuMemMap failed with CUresult %d\n", (int)r);
    return 20000 + (int)r;
  }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.SetAccess(va, bytes, &acc, 1);
  if (r != CUDA_SUCCESS) {
    d.MemUnmap(va, bytes);
    d.AddrFree(va, bytes);
    fprintf(stderr, "[dlcfn vmm] cuMemSetAccess failed with CUresult %d\n", (int)r);
    return 20000 + (int)r;
  }
  *out_ptr = (void*)va;
  return 0;
}

}  // namespace

// bit 0: exportable VMM allocations (POSIX fd handles) are supported on `dev`; bit 1: NVLS multicast objects are supported
DLCFN_API int dlcfn_vmm_supported(int dev) {
  Drv& d = drv();
  if (!d.ok) return 0;
  CUdevice cd;
  if (d.DevGet(&cd, dev) != CUDA_SUCCESS) return 0;
  int vmm = 0, fd = 0, mc = 0;
  d.DevGetAttr(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cd);
  d.DevGetAttr(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cd);
  if (d.mc_ok) d.DevGetAttr(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd);
  return ((vmm && fd) ? 1 : 0) | ((vmm && fd && mc) ? 2 : 0);
}

// allocation / mapping granularity that satisfies both the arena and (world > 1) the multicast object
DLCFN_API size_t dlcfn_vmm_granularity(int dev, int world, int with_multicast) {
  Drv& d = drv();
  if (!d.ok) return 0;
  CUmemAllocationProp p = arena_prop(dev);
  size_t g = 0;
  if (d.Granularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || g == 0) return 0;
  if (with_multicast && d.mc_ok && world > 1) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)world;
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (d.McGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
  }
  return g;
}

// New arena of `bytes` (a multiple of the granularity) on `dev`: mapped locally, exported as a file descriptor.
DLCFN_API int dlcfn_vmm_alloc(size_t bytes, int dev, size_t gran, void** out_ptr, unsigned long long* out_handle, int* out_fd) {
  Drv& d = drv();
  if (!d.ok || bytes == 0 || gran == 0 || bytes % gran) return (int)cudaErrorInvalidValue;
  CUmemAllocationProp p = arena_prop(dev);
  CUmemGenericAllocationHandle h;
  DRV_CHECK(d.MemCreate(&h, bytes, &p, 0));
  int fd = -1;
  CUresult r = d.MemExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    d.MemRelease(h);
    fprintf(stderr, "[dlcfn vmm] cuMemExportToShareableHandle failed with CUresult %d\n", (int)r);
    return 20000 + (int)r;
  }
  int rc = map_handle(h, bytes, gran, dev, out_ptr);
  if (rc) {
    close(fd);
    d.MemRelease(h);
    return rc;
  }
  *out_handle = (unsigned long long)h;
  *out_fd = fd;
  return 0;
}

// A peer's arena from its file descriptor: imported and mapped with access for `dev` (the LOCAL device).
DLCFN_API int dlcfn_vmm_import(int fd, size_t bytes, int dev, size_t gran, void** out_ptr, unsigned long long* out_handle) {
  Drv& d = drv();
  if (!d.ok || fd < 0) return (int)cudaErrorInvalidValue;
  CUmemGenericAllocationHandle h;
  DRV_CHECK(d.MemImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  int rc = map_handle(h, bytes, gran, dev, out_ptr);
  if (rc) {
    d.MemRelease(h);
    return rc;
  }
  *out_handle = (unsigned long long)h;
  return 0;
}

// rank 0: the multicast object for `world` devices
DLCFN_API int dlcfn_vmm_mc_create(size_t bytes, int world, unsigned long long* out_mc, int* out_fd) {
  Drv& d = drv();
  if (!d.mc_ok) return (int)cudaErrorNotSupported;
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = (unsigned)world;
  mp.size = bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle mc;
  DRV_CHECK(d.McCreate(&mc, &mp));
  int fd = -1;
  CUresult r = d.MemExport(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    d.MemRelease(mc);
    return 20000 + (int)r;
  }
  *out_mc = (unsigned long long)mc;
  *out_fd = fd;
  return 0;
}

DLCFN_API int dlcfn_vmm_mc_import(int fd, unsigned long long* out_mc) {
  Drv& d = drv();
  if (!d.mc_ok || fd < 0) return (int)cudaErrorNotSupported;
  CUmemGenericAllocationHandle mc;
  DRV_CHECK(d.MemImport(&mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  *out_mc = (unsigned long long)mc;
  return 0;
}

// every rank, BEFORE anyone binds: join the multicast team with the local device
DLCFN_API int dlcfn_vmm_mc_add_device(unsigned long long mc, int dev) {
  Drv& d = drv();
  if (!d.mc_ok) return (int)cudaErrorNotSupported;
  CUdevice cd;
  DRV_CHECK(d.DevGet(&cd, dev));
  DRV_CHECK(d.McAddDevice((CUmemGenericAllocationHandle)mc, cd));
  return 0;
}

// every rank, after ALL ranks have added their device: bind the local arena and map the multicast object
DLCFN_API int dlcfn_vmm_mc_bind_map(unsigned long long mc, unsigned long long arena, size_t bytes, int dev, size_t gran,
                                    void** out_mc_ptr) {
  Drv& d = drv();
  if (!d.mc_ok) return (int)cudaErrorNotSupported;
  DRV_CHECK(d.McBindMem((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)arena, 0, bytes, 0));
  return map_handle((CUmemGenericAllocationHandle)mc, bytes, gran, dev, out_mc_ptr);
}

DLCFN_API int dlcfn_vmm_unmap_release(void* ptr, unsigned long long handle, size_t bytes) {
  Drv& d = drv();
  if (!d.ok) return (int)cudaErrorNotSupported;
  if (ptr) {
    d.MemUnmap((CUdeviceptr)ptr, bytes);
    d.AddrFree((CUdeviceptr)ptr, bytes);
  }
  if (handle) d.MemRelease((CUmemGenericAllocationHandle)handle);
  return 0;
}

// Native fabric probe for the first `ndev` devices: access[i*ndev+j] = peer access supported, perf[...] = relative performance
// rank of the link (lower is better, cudaDevP2PAttrPerformanceRank), atomics[...] = native atomics over the link,
// multicast[i] = NVLS multicast support.  Returns the number of devices probed (<= ndev), or a negative CUDA error.
DLCFN_API int dlcfn_probe_fabric(int ndev, int* access, int* perf, int* atomics, int* multicast) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) return -1;
  if (ndev > count) ndev = count;
  Drv& d = drv();
  for (int i = 0; i < ndev; ++i) {
    if (multicast) {
      int mc = 0;
      CUdevice cd;
      if (d.mc_ok && d.DevGet(&cd, i) == CUDA_SUCCESS) d.DevGetAttr(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd);
      multicast[i] = mc;
    }
    for (int j = 0; j < ndev; ++j) {
      int a = (i == j), pr = 0, at = (i == j);
      if (i != j) {
        cudaDeviceGetP2PAttribute(&a, cudaDevP2PAttrAccessSupported, i, j);
        cudaDeviceGetP2PAttribute(&pr, cudaDevP2PAttrPerformanceRank, i, j);
        cudaDeviceGetP2PAttribute(&at, cudaDevP2PAttrNativeAtomicSupported, i, j);
      }
      if (access) access[i * ndev + j] = a;
      if (perf) perf[i * ndev + j] = pr;
      if (atomics) atomics[i * ndev + j] = at;
    }
  }
  (void)cudaGetLastError();
  return ndev;
}
