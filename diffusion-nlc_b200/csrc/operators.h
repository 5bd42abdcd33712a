// Shared between operators.cu (A, A^T, A^+, projection) and ddnm.cu (DDNM+ terms and the fused DDNM reverse step):
// the operator object and the transform / GEMM launchers both files build their pipelines from.
#pragma once
#include "common.h"

struct nlc_op {
    nlc_ctx* ctx;
    int task, C, R, ratio, m;
    int64_t ydim;
    int* idx_a;   // INPAINT: kept[k] = pixel*3+c ; WHCS: invperm[q]
    int* idx_b;   // INPAINT: pos2k[pixel*3+c] (-1 = missing)
    int* idx_c;   // INPAINT: the same table in image order, pos2k_planar[c*HW + pixel] (128-bit loads next to the image)
    int n_kept;
    float u, s;
    float* v0;    // COLOR: 3 ; SR_AVG: r*r
    float* Vfull; // COLOR / SR_AVG: the whole K x K V_small, row-major (Lambda / Lambda_noise rotate with it)
    float* Vfull_host;  // the same on the host (handed to the kernels by value for K <= 16)
    int K;        // needle length: 3 (colour) or r*r
    float *Us, *Vs, *mult, *pinv;  // SEPARABLE (left factors; also right factors unless Us2 / Vs2 are set)
    float *Us2, *Vs2;              // SEPARABLE: right factors (== Us / Vs for one-kernel operators)
    float* lam_s;                  // SEPARABLE: [m*m] singular value per spectral position for Lambda (nullptr: none)
    bool own2;
    int64_t nx;                    // GENERAL: columns of A (Vs is [nx,nx], Us [m,m], v0 the m singular values)
};

namespace nlc {

// out = alpha * base + beta * value + g1 * add1 + g2 * add2   (terms with a null pointer are absent; base == nullptr
// leaves the plain transform)
struct Epilogue {
    const float* base = nullptr;
    float alpha = 1.f, beta = 1.f;
    const float* add1 = nullptr;
    float g1 = 0.f;
    const float* add2 = nullptr;
    float g2 = 0.f;
    long long add2_stride = 0;  // elements between the samples of add2 (0 = dense C*R*R): a [B,6,R,R] network output
    // WH-CS: work on the transform output itself (spectral position q = r*R + c, kept when invperm[q] < m), before the
    // terms above:  ymeas   -> value = kept ? value - ymeas[b][invperm[q]*C + c] : 0   (the projection's residual)
    //               ygather -> kept values go to ygather[b][invperm[q]*C + c], nothing else is written (A x)
    const int* invperm = nullptr;
    int m = 0;
    const float* ymeas = nullptr;
    float* ygather = nullptr;
};

#ifdef __CUDACC__
// Shared by the two FWHT implementations: what happens to transform value `v` of plane (b, c), entry (r, col), flat index
// o.  Returns false when nothing is to be stored at o (the gather wrote elsewhere).
__device__ __forceinline__ bool fwht_epilogue(const Epilogue& e, float& v, int b, int c, int C, int R, int r, int col,
                                              size_t o) {
    if (e.invperm) {
        const int j = e.invperm[r * R + col];
        const size_t yo = (static_cast<size_t>(b) * e.m + j) * C + c;
        if (e.ygather) {
            if (j < e.m) e.ygather[yo] = v;
            return false;
        }
        if (e.ymeas) v = j < e.m ? v - e.ymeas[yo] : 0.f;
    }
    if (e.base) v = e.alpha * e.base[o] + e.beta * v;
    if (e.add1) v += e.g1 * e.add1[o];
    if (e.add2) {
        const size_t o2 = e.add2_stride ? static_cast<size_t>(b) * e.add2_stride + (static_cast<size_t>(c) * R + r) * R + col : o;
        v += e.g2 * e.add2[o2];
    }
    return true;
}
#endif

// One-kernel 2-D FWHT for R = 64 .. 512 (fwht_cluster.cu): a thread-block cluster holds one plane in distributed shared
// memory.  Returns NLC_OK, or 1 when the size is not covered (the caller falls back to the two-kernel transform).
// Experimental (opt-in with NLC_FWHT_CLUSTER=1): correct and bit-identical, but slower than the two-kernel path.
int fwht2d_cluster(const float* in, float* out, const Epilogue& epi, int planes, int C, int R, cudaStream_t st);

// orthonormal 2-D fast Walsh-Hadamard transform of B*C planes (the reference's 1-D FWHT over R^2 entries)
int fwht2d(nlc_op* op, const float* in, float* out, const Epilogue& epi, int B, cudaStream_t st);

// Cm[b] = epi( A[b] * Bm[b] ): strided batched fp32 GEMM; epi = (* mult[(b % nch)]) (- sub) (alpha*base + beta*v) (+ add)
int launch_gemm(cudaStream_t st, int batch, int M, int N, int K, const float* A, long long sab, long long sai,
                long long sak, const float* Bm, long long sbb, long long sbk, long long sbj, float* Cm, const float* mult,
                int nch, const float* sub, const float* base, float alpha = 1.f, float beta = -1.f,
                const float* add = nullptr);

int separable_A(nlc_op* op, const float* x, int B, float* y, float* ws, const float* sub, cudaStream_t st);

inline unsigned blocks_for(long long n, int bs = 256) { return static_cast<unsigned>((n + bs - 1) / bs); }

}  // namespace nlc
