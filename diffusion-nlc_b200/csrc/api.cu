// Context lifetime and error reporting of the C ABI (include/nlc_b200.h).
#include <string.h>

#include <stdlib.h>

#include "common.h"

namespace nlc {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("NLC_PDL");  // opt-in: measured neutral inside the CUDA-graph replay
        return e && e[0] == '1';
    }();
    return on;
}

}  // namespace nlc

extern "C" {

const char* nlc_last_error(void) { return nlc::g_err; }

int nlc_abi_version(void) { return 1; }

nlc_ctx* nlc_create(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        nlc::set_error(NLC_ECUDA, "nlc_create: CUDA device %d not available (%d visible)", device, ndev);
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        nlc::set_error(NLC_ECUDA, "nlc_create: cudaGetDeviceProperties failed");
        return nullptr;
    }
    if (prop.major != 10) {
        nlc::set_error(NLC_ENOTSUP, "nlc_create: device %d is sm_%d%d; this library is built for sm_100a only",
                       device, prop.major, prop.minor);
        return nullptr;
    }
    cudaSetDevice(device);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        nlc::set_error(NLC_ECUDA, "nlc_create: cuTensorMapEncodeTiled not found in the driver");
        return nullptr;
    }
    nlc_ctx* ctx = new nlc_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->encode_tiled = reinterpret_cast<nlc::encode_tiled_fn>(fn);
    const char* e = getenv("NLC_CTA_PAIRS");
    ctx->use_cta_pairs = !(e && e[0] == '0');
    e = getenv("NLC_SLAB");
    ctx->use_slab = e ? atoi(e) : 1;
    e = getenv("NLC_SPLITK");
    ctx->use_splitk = !(e && e[0] == '0');
    ctx->splitk_ws = nullptr, ctx->splitk_bytes = 0;
    e = getenv("NLC_ATTN_ONEPASS");
    ctx->attn_onepass = !(e && e[0] == '0');
    e = getenv("NLC_TMA_EPI");
    ctx->use_tma_epi = e ? atoi(e) : 1;
    return ctx;
}

void nlc_destroy(nlc_ctx* ctx) {
    if (ctx && ctx->splitk_ws) cudaFree(ctx->splitk_ws);
    delete ctx;
}

int nlc_sm_count(nlc_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int nlc_ctx_set(nlc_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return nlc::set_error(NLC_EINVAL, "nlc_ctx_set: null argument");
    if (!strcmp(key, "cta_pairs")) {
        ctx->use_cta_pairs = value != 0;
    } else if (!strcmp(key, "slab")) {
        if (value < 0 || value > 2) return nlc::set_error(NLC_EINVAL, "nlc_ctx_set: slab must be 0, 1 or 2");
        ctx->use_slab = value;
    } else if (!strcmp(key, "splitk")) {
        ctx->use_splitk = value != 0;
    } else if (!strcmp(key, "attn_onepass")) {
        ctx->attn_onepass = value != 0;
    } else if (!strcmp(key, "tma_epi")) {
        if (value < 0 || value > 2) return nlc::set_error(NLC_EINVAL, "nlc_ctx_set: tma_epi must be 0, 1 or 2");
        ctx->use_tma_epi = value;
    } else {
        return nlc::set_error(NLC_EINVAL, "nlc_ctx_set: unknown key '%s'", key);
    }
    return NLC_OK;
}

}  // extern "C"
