// C-ABI plumbing: error string, version, TMA descriptor encoding through the driver entry point
// (no link-time dependency on libcuda), SM count cache.
#include <cstdlib>
#include "api_internal.h"

#include <atomic>
#include <string.h>

namespace uwu {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle128) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver / GPU?)");
        return UWU_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
        set_error("TMA base pointer %p is not 16-byte aligned", base);
        return UWU_ERR_INVALID;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (box[i] == 0 || box[i] > 256) {
            set_error("TMA box dim %d = %u out of range", i, box[i]);
            return UWU_ERR_INVALID;
        }
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        if (gstr[i] % 16 != 0) {
            set_error("TMA stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)gstr[i]);
            return UWU_ERR_INVALID;
        }
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu box %u,%u)", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return UWU_ERR_INVALID;
    }
    return 0;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("UWU_PDL");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        n = prop.multiProcessorCount;
    }
    return n;
}

std::atomic<long long> g_launches{0};

}  // namespace uwu

extern "C" const char* uwu_last_error(void) { return uwu::g_err; }
extern "C" int uwu_version(void) { return 100; }
extern "C" int64_t uwu_launch_count(void) { return (int64_t)uwu::g_launches.load(); }
