// Host-side plumbing shared by all translation units of libnlc_b200: error reporting, the context object
// and the driver entry point used to encode TMA tensor maps (resolved at run time so the library links
// against the CUDA runtime only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nlc_b200.h"

namespace nlc {

int set_error(int code, const char* fmt, ...);

#define NLC_CHECK_CUDA(expr)                                                                          \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return nlc::set_error(NLC_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                                  __FILE__, __LINE__);                                                \
    } while (0)

#define NLC_CHECK_LAUNCH()                                                                            \
    do {                                                                                              \
        cudaError_t _e = cudaGetLastError();                                                          \
        if (_e != cudaSuccess)                                                                        \
            return nlc::set_error(NLC_ECUDA, "kernel launch failed: %s (%s:%d)",                      \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);                        \
    } while (0)

#define NLC_REQUIRE(cond, ...)                                        \
    do {                                                              \
        if (!(cond)) return nlc::set_error(NLC_EINVAL, __VA_ARGS__);  \
    } while (0)

// Operand dtype helpers: 16-bit operands (bf16 / fp16) against fp32 containers (tf32-rounded / plain), and the one-bit
// "format" flag the kernels take: fp32 containers -> round to tf32 (NLC_F32), 16-bit -> fp16 instead of bf16 (NLC_F16).
inline bool dtype_valid(int dt) { return dt == NLC_F32 || dt == NLC_BF16 || dt == NLC_F32X3 || dt == NLC_F16; }
inline bool dtype_is16(int dt) { return dt == NLC_BF16 || dt == NLC_F16; }
inline int dtype_fmt(int dt) { return dt == NLC_F32 || dt == NLC_F16; }

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: one flag per (kernel instantiation, device),
// not one per process, so a second context on another GPU configures its own copy of the function.
struct PerDeviceFlag {
    bool done[64] = {};
    bool& operator[](int dev) { return done[dev & 63]; }
};

// Launch with programmatic stream serialization (ptx.cuh: pdl_wait / pdl_trigger; the kernel MUST call pdl_wait() before it
// touches global memory).  Inside a captured CUDA graph the attribute becomes a programmatic dependency edge.  Opt-in
// (NLC_PDL=1): by default the same kernels are launched in ordinary stream order, where pdl_wait() is a no-op.
bool pdl_enabled();
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<A&&>(args)...);
}

}  // namespace nlc

// The entry points that configure / launch kernels with opt-in shared memory run on the CURRENT device: it must be the
// context's (one nlc_ctx per (process, device); the caller selects the device, as torch.cuda.set_device does).
#define NLC_REQUIRE_DEVICE(ctx)                                                                                   \
    do {                                                                                                          \
        int _cur = -1;                                                                                            \
        NLC_CHECK_CUDA(cudaGetDevice(&_cur));                                                                     \
        if (_cur != (ctx)->device)                                                                                \
            return nlc::set_error(NLC_EINVAL, "%s: the current CUDA device is %d but the context belongs to %d", \
                                  __func__, _cur, (ctx)->device);                                                 \
    } while (0)

struct nlc_ctx {
    int device;
    int sm_count;
    nlc::encode_tiled_fn encode_tiled;
    int use_cta_pairs;  // tcgen05 cta_group::2 conv kernel for large layers (NLC_CTA_PAIRS=0 disables)
    int use_slab;       // halo-slab 3x3 kernel (conv_slab.cu): 0 off, 1 layers with 128 output channels, 2 every eligible layer
    int attn_onepass;   // fused attention with one pass over the keys (online softmax); NLC_ATTN_ONEPASS=0: the two-pass kernel
    int use_splitk;     // split-K of the small-M convolution launches (NLC_SPLITK=0 disables)
    void* splitk_ws;    // its workspace: one fixed allocation (captured graphs hold the address); one stream at a time
    size_t splitk_bytes;
    int use_tma_epi;    // 16-bit conv epilogues through TMA (residual block by tensor load, output by tensor store); NLC_TMA_EPI=0 disables
};
