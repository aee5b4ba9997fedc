// NVSwitch multicast ("NVLS") memory for the gradient bucket: C ABI of include/avconnector_b200.h, section
// "multicast bucket".  Host code only (CUDA driver VMM + multicast API, resolved at run time so that the library
// still loads on a box without a driver).  With the bucket of every rank bound to one multicast object, the comm warps
// of the fused dW + all-reduce GEMM use multimem.ld_reduce (sum of every rank's copy, added inside the switch) and
// multimem.st (one store that lands in every rank's copy) instead of `world` peer loads and `world` peer stores.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include "../../include/avconnector_b200.h"

extern "C" int avc_set_error_(int code, const char* msg);  // avc_capi.cu: records the thread-local message

namespace {

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return avc_set_error_(code, buf);
}

template <typename Fn>
Fn entry(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    p = nullptr;
  return reinterpret_cast<Fn>(p);
}

struct Driver {
  PFN_cuMemCreate create = entry<PFN_cuMemCreate>("cuMemCreate");
  PFN_cuMemRelease release = entry<PFN_cuMemRelease>("cuMemRelease");
  PFN_cuMemAddressReserve reserve = entry<PFN_cuMemAddressReserve>("cuMemAddressReserve");
  PFN_cuMemAddressFree addr_free = entry<PFN_cuMemAddressFree>("cuMemAddressFree");
  PFN_cuMemMap map = entry<PFN_cuMemMap>("cuMemMap");
  PFN_cuMemUnmap unmap = entry<PFN_cuMemUnmap>("cuMemUnmap");
  PFN_cuMemSetAccess set_access = entry<PFN_cuMemSetAccess>("cuMemSetAccess");
  PFN_cuMemExportToShareableHandle export_handle = entry<PFN_cuMemExportToShareableHandle>("cuMemExportToShareableHandle");
  PFN_cuMemImportFromShareableHandle import_handle =
      entry<PFN_cuMemImportFromShareableHandle>("cuMemImportFromShareableHandle");
  PFN_cuMulticastCreate mc_create = entry<PFN_cuMulticastCreate>("cuMulticastCreate");
  PFN_cuMulticastAddDevice mc_add_device = entry<PFN_cuMulticastAddDevice>("cuMulticastAddDevice");
  PFN_cuMulticastBindMem mc_bind_mem = entry<PFN_cuMulticastBindMem>("cuMulticastBindMem");
  PFN_cuMulticastUnbind mc_unbind = entry<PFN_cuMulticastUnbind>("cuMulticastUnbind");
  PFN_cuMulticastGetGranularity mc_granularity = entry<PFN_cuMulticastGetGranularity>("cuMulticastGetGranularity");
  PFN_cuDeviceGet device_get = entry<PFN_cuDeviceGet>("cuDeviceGet");
  PFN_cuDeviceGetAttribute device_attr = entry<PFN_cuDeviceGetAttribute>("cuDeviceGetAttribute");
  bool ok() const {
    return device_get && device_attr && create && release && reserve && addr_free && map && unmap && set_access && export_handle && import_handle &&
           mc_create && mc_add_device && mc_bind_mem && mc_unbind && mc_granularity;
  }
};

const Driver* driver() {
  static Driver d;
  return d.ok() ? &d : nullptr;
}

int cu_fail(CUresult r, const char* what) { return fail(AVC_ERR_CUDA, "%s failed with CUresult %d", what, static_cast<int>(r)); }

CUmulticastObjectProp mc_prop(int world, uint64_t bytes) {
  CUmulticastObjectProp p;
  memset(&p, 0, sizeof(p));
  p.numDevices = static_cast<unsigned>(world);
  p.size = bytes;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

// reserve a VA range, map `handle` into it, give the current device read / write access
int map_handle(const Driver* d, CUmemGenericAllocationHandle handle, uint64_t bytes, int device, void** out,
               const char* what) {
  CUdeviceptr va = 0;
  CUresult r = d->reserve(&va, bytes, 0, 0, 0);
  if (r != CUDA_SUCCESS) return cu_fail(r, "cuMemAddressReserve");
  r = d->map(va, bytes, 0, handle, 0);
  if (r != CUDA_SUCCESS) {
    d->addr_free(va, bytes);
    return fail(AVC_ERR_CUDA, "cuMemMap (%s) failed with CUresult %d", what, static_cast<int>(r));
  }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d->set_access(va, bytes, &acc, 1);
  if (r != CUDA_SUCCESS) {
    d->unmap(va, bytes);
    d->addr_free(va, bytes);
    return fail(AVC_ERR_CUDA, "cuMemSetAccess (%s) failed with CUresult %d", what, static_cast<int>(r));
  }
  *out = reinterpret_cast<void*>(va);
  return AVC_OK;
}

}  // namespace

extern "C" {

int avc_mc_supported(int32_t device, int32_t* supported) {
  if (supported == nullptr) return fail(AVC_ERR_INVALID, "mc_supported: null output");
  *supported = 0;
  const Driver* d = driver();
  if (d == nullptr) return AVC_OK;
  cudaFree(nullptr);  // the driver API needs an initialised context
  CUdevice dev = 0;
  int v = 0;
  if (d->device_get(&dev, device) != CUDA_SUCCESS) return AVC_OK;
  if (d->device_attr(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return AVC_OK;
  *supported = v != 0;
  return AVC_OK;
}

int avc_mc_padded_bytes(int32_t world, uint64_t min_bytes, uint64_t* padded_bytes) {
  const Driver* d = driver();
  if (d == nullptr) return fail(AVC_ERR_UNSUPPORTED, "multicast: driver entry points not available");
  if (world < 1 || min_bytes == 0 || padded_bytes == nullptr) return fail(AVC_ERR_INVALID, "mc_padded_bytes: bad argument");
  cudaFree(nullptr);  // make sure the primary context exists
  CUmulticastObjectProp p = mc_prop(world, min_bytes);
  size_t gran = 0;
  CUresult r = d->mc_granularity(&gran, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS || gran == 0) return cu_fail(r, "cuMulticastGetGranularity");
  *padded_bytes = (min_bytes + gran - 1) / gran * gran;
  return AVC_OK;
}

int avc_mc_create(int32_t world, uint64_t padded_bytes, uint64_t* mc_handle, int32_t* fd) {
  const Driver* d = driver();
  if (d == nullptr) return fail(AVC_ERR_UNSUPPORTED, "multicast: driver entry points not available");
  if (mc_handle == nullptr || fd == nullptr) return fail(AVC_ERR_INVALID, "mc_create: null output");
  cudaFree(nullptr);
  CUmulticastObjectProp p = mc_prop(world, padded_bytes);
  CUmemGenericAllocationHandle h = 0;
  CUresult r = d->mc_create(&h, &p);
  if (r != CUDA_SUCCESS) return cu_fail(r, "cuMulticastCreate");
  int out_fd = -1;
  r = d->export_handle(&out_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    d->release(h);
    return cu_fail(r, "cuMemExportToShareableHandle (multicast object)");
  }
  *mc_handle = h;
  *fd = out_fd;
  return AVC_OK;
}

int avc_mc_import(int32_t fd, uint64_t* mc_handle) {
  const Driver* d = driver();
  if (d == nullptr) return fail(AVC_ERR_UNSUPPORTED, "multicast: driver entry points not available");
  if (mc_handle == nullptr || fd < 0) return fail(AVC_ERR_INVALID, "mc_import: bad argument");
  cudaFree(nullptr);
  CUmemGenericAllocationHandle h = 0;
  CUresult r = d->import_handle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) return cu_fail(r, "cuMemImportFromShareableHandle (multicast object)");
  *mc_handle = h;
  return AVC_OK;
}

int avc_mc_add_device(uint64_t mc_handle) {
  const Driver* d = driver();
  if (d == nullptr) return fail(AVC_ERR_UNSUPPORTED, "multicast: driver entry points not available");
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(AVC_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  CUresult r = d->mc_add_device(mc_handle, dev);
  if (r != CUDA_SUCCESS) return cu_fail(r, "cuMulticastAddDevice");
  return AVC_OK;
}

int avc_mc_bucket_alloc(uint64_t mc_handle, uint64_t padded_bytes, avc_mc_bucket* out) {
  const Driver* d = driver();
  if (d == nullptr) return fail(AVC_ERR_UNSUPPORTED, "multicast: driver entry points not available");
  if (out == nullptr || padded_bytes == 0) return fail(AVC_ERR_INVALID, "mc_bucket_alloc: bad argument");
  memset(out, 0, sizeof(*out));
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(AVC_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle mem = 0;
  CUresult r = d->create(&mem, padded_bytes, &ap, 0);
  if (r != CUDA_SUCCESS) return cu_fail(r, "cuMemCreate");
  r = d->mc_bind_mem(mc_handle, 0, mem, 0, padded_bytes, 0);
  if (r != CUDA_SUCCESS) {
    d->release(mem);
    return cu_fail(r, "cuMulticastBindMem");
  }
  void* uc = nullptr;
  void* mc = nullptr;
  if (int rc = map_handle(d, mem, padded_bytes, dev, &uc, "bucket")) {
    d->release(mem);
    return rc;
  }
  if (int rc = map_handle(d, mc_handle, padded_bytes, dev, &mc, "multicast object")) {
    d->unmap(reinterpret_cast<CUdeviceptr>(uc), padded_bytes);
    d->addr_free(reinterpret_cast<CUdeviceptr>(uc), padded_bytes);
    d->release(mem);
    return rc;
  }
  e = cudaMemset(uc, 0, padded_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(AVC_ERR_CUDA, "mc_bucket_alloc memset: %s", cudaGetErrorString(e));
  out->ptr = uc;
  out->mc_ptr = mc;
  out->bytes = padded_bytes;
  out->mem_handle = mem;
  out->mc_handle = mc_handle;
  return AVC_OK;
}

int avc_mc_bucket_free(avc_mc_bucket* b) {
  const Driver* d = driver();
  if (d == nullptr || b == nullptr || b->ptr == nullptr) return AVC_OK;
  cudaDeviceSynchronize();
  int dev = 0;
  cudaGetDevice(&dev);
  d->unmap(reinterpret_cast<CUdeviceptr>(b->mc_ptr), b->bytes);
  d->addr_free(reinterpret_cast<CUdeviceptr>(b->mc_ptr), b->bytes);
  d->mc_unbind(b->mc_handle, dev, 0, b->bytes);
  d->unmap(reinterpret_cast<CUdeviceptr>(b->ptr), b->bytes);
  d->addr_free(reinterpret_cast<CUdeviceptr>(b->ptr), b->bytes);
  d->release(b->mem_handle);
  d->release(b->mc_handle);
  memset(b, 0, sizeof(*b));
  return AVC_OK;
}

}  // extern "C"
