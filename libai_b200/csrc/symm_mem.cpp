// Symmetric (peer-mapped) device memory for the in-kernel NVLink collectives.
//
// Every rank of a group allocates a buffer of the same size with cudaMalloc, exports a CUDA IPC
// handle, and maps the buffers of all peers into its own address space.  Kernels then receive a
// table of peer pointers and issue plain ld/st/red on them — the transfers ride NVLink 5 through
// the NVSwitch (one process per GPU, single node).  The python side
// (libai_b200/parallel/symm_mem.py) exchanges the 64-byte handles with torch.distributed.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>
#include <torch/library.h>
#include <torch/torch.h>

#include <cstring>
#include <mutex>
#include <unordered_map>

namespace {

using at::Tensor;

#define LB_CUDA_CHECK(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    TORCH_CHECK(_e == cudaSuccess, "CUDA error in symm_mem: ", cudaGetErrorString(_e), " at ", #expr); \
  } while (0)

std::mutex g_mu;
std::unordered_map<int64_t, void*> g_opened;  // peer ptr -> ptr (for closing)

// Allocate `nbytes` of zero-initialised device memory outside the caching allocator (IPC handles
// must refer to the base of a cudaMalloc allocation). Returns a uint8 tensor that frees on destruction.
Tensor symm_alloc(int64_t nbytes) {
  int dev = 0;
  LB_CUDA_CHECK(cudaGetDevice(&dev));
  void* ptr = nullptr;
  LB_CUDA_CHECK(cudaMalloc(&ptr, static_cast<size_t>(nbytes)));
  LB_CUDA_CHECK(cudaMemset(ptr, 0, static_cast<size_t>(nbytes)));
  auto opts = at::TensorOptions().dtype(at::kByte).device(at::kCUDA, dev);
  return at::from_blob(ptr, {nbytes}, [](void* p) { cudaFree(p); }, opts);
}

// 64-byte IPC handle of an allocation made by symm_alloc (as a CPU uint8 tensor)
Tensor symm_export(const Tensor& buf) {
  cudaIpcMemHandle_t h;
  LB_CUDA_CHECK(cudaIpcGetMemHandle(&h, buf.data_ptr()));
  Tensor out = at::empty({static_cast<int64_t>(sizeof(h))}, at::kByte);
  std::memcpy(out.data_ptr(), &h, sizeof(h));
  return out;
}

// Map a peer's allocation; returns the device pointer (valid in this process) as int64
int64_t symm_open(const Tensor& handle) {
  TORCH_CHECK(handle.numel() == static_cast<int64_t>(sizeof(cudaIpcMemHandle_t)), "bad IPC handle size");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle.data_ptr(), sizeof(h));
  void* ptr = nullptr;
  LB_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
  std::lock_guard<std::mutex> lk(g_mu);
  g_opened[reinterpret_cast<int64_t>(ptr)] = ptr;
  return reinterpret_cast<int64_t>(ptr);
}

void symm_close(int64_t ptr) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_opened.find(ptr);
  if (it != g_opened.end()) {
    cudaIpcCloseMemHandle(it->second);
    g_opened.erase(it);
  }
}

// View `nbytes` at device pointer `ptr` (own or peer memory) as a tensor of the given dtype/shape
Tensor symm_view(int64_t ptr, at::IntArrayRef sizes, at::ScalarType dtype, int64_t device_index) {
  auto opts = at::TensorOptions().dtype(dtype).device(at::kCUDA, device_index);
  return at::from_blob(reinterpret_cast<void*>(ptr), sizes, [](void*) {}, opts);
}

bool can_access_peer(int64_t dev, int64_t peer) {
  int ok = 0;
  cudaDeviceCanAccessPeer(&ok, static_cast<int>(dev), static_cast<int>(peer));
  return ok != 0;
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(libai_b200, m) {
  m.def("symm_alloc(int nbytes) -> Tensor", &symm_alloc);
  m.def("symm_export(Tensor buf) -> Tensor", &symm_export);
  m.def("symm_open(Tensor handle) -> int", &symm_open);
  m.def("symm_close(int ptr) -> ()", &symm_close);
  m.def("symm_view(int ptr, int[] sizes, ScalarType dtype, int device_index) -> Tensor", &symm_view);
  m.def("can_access_peer(int dev, int peer) -> bool", &can_access_peer);
}
