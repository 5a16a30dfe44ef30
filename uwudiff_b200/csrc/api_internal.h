// Internal helpers shared by the C-ABI translation units (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/uwu_b200.h"

namespace uwu {

// thread-local message returned by uwu_last_error()
void set_error(const char* fmt, ...);

// Encode a tiled TMA descriptor for a bf16 tensor (rank 2..4). dims/strides innermost first,
// strides in BYTES for dims 1..rank-1. Returns 0 on success, else sets the error string.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle128);

int sm_count();
bool pdl_enabled();  // programmatic dependent launch for the TMEM kernels (UWU_PDL=0 disables)

extern std::atomic<long long> g_launches;

// launch with the programmatic stream-serialization attribute (see pdl_trigger / pdl_wait in common.cuh)
template <typename Arg>
inline cudaError_t launch_pdl(void (*kernel)(Arg), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const Arg& arg) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, arg);
}

}  // namespace uwu

#define UWU_CHECK_ARG(cond, ...)         \
    do {                                 \
        if (!(cond)) {                   \
            uwu::set_error(__VA_ARGS__); \
            return UWU_ERR_INVALID;      \
        }                                \
    } while (0)

#define UWU_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            uwu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                    \
        }                                                                                      \
    } while (0)

#define UWU_CHECK_LAUNCH()                                   \
    do {                                                     \
        uwu::g_launches.fetch_add(1, std::memory_order_relaxed); \
        UWU_CHECK_CUDA(cudaGetLastError());                  \
    } while (0)
