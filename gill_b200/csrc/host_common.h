// Host-side helpers shared by every translation unit of libgillb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cerrno>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <utility>

namespace gb {

// thread-local error string returned by gillb200_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define GB_CHECK_ARG(cond, ...)                       \
  do {                                                \
    if (!(cond)) return gb::set_err(-EINVAL, __VA_ARGS__); \
  } while (0)

#define GB_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) return gb::set_err(-EIO, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// dims/strides innermost-first; strides in bytes for dims 1..rank-1. 16-bit elements, SWIZZLE_128B.
int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool is_bf16);

// general form: dtype 0 = bf16, 1 = fp16, 2 = fp32; swizzle_bytes 128 / 64 / 32
// elem_strides: optional per-dimension traversal strides (nullptr = all 1)
int encode_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes, const uint32_t* elem_strides);

int num_sms();

// cudaFuncSetAttribute (dynamic shared memory opt-in) is per device: a launcher configures its kernel once on every
// device the process touches, not once per process.
struct PerDeviceOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// Launch, optionally (GILLB200_PDL=1) with the programmatic-dependent-launch attribute: the kernel must call
// gb::pdl_wait() before its first global-memory access. Captured into CUDA graphs as programmatic dependency edges.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Same, for the kernels around the GEMMs (norms, attention, elementwise): GILLB200_PDL_LIGHT=1 gives only THEM the
// programmatic-launch attribute -- the GEMM / conv launches (stream-K flags) stay fully serialised.
bool pdl_light_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_light(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() || pdl_light_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// number of kernel launches issued by this library (bench.py reports it as gpu_launches)
extern long long g_launch_count;
#define GB_COUNT_LAUNCH(n) (gb::g_launch_count += (n))

}  // namespace gb
