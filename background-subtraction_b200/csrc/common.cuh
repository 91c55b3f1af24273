// common.cuh -- shared device/host definitions for the B200 (sm_100a) LSD / group-sparse RPCA solver.
//
// Device data layout: the reference's Fortran-order m x n matrix
// (pixels x frames, /root/reference/inexact_alm_lsd.py:84-88,225) is byte-for-byte a row-major
// [n frames][m pixels] array; on the device every matrix is float32 [n][ld] with the pixel index
// contiguous, ld = m rounded up to 32, pad columns kept at zero.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace bsub {

constexpr int kMaxIterLog = 512;

// Solver state that lives on the device for the whole run (no host round trips inside the loop).
struct DevState {
    // scalars fixed at init
    double lambda;          // 1/(sqrt(max(m_global, n)) * delta)
    double non_block_lambda;
    double norm_two;        // ||D||_2
    double norm_rowsum;     // max_p sum_f |D[p,f]|   (NumPy induced inf-norm)
    double dual_norm;
    double normD2;          // ||D||_F^2
    double rho, tol, mu_scale;
    // evolving
    double mu;              // mu_k used by the current iteration
    double mu_iter;         // mu of the iteration whose shrink pass ran last (mu itself has already been advanced)
    double thresh;          // 1/mu
    double zz;              // sum Z^2 of the last shrink pass
    double err;
    unsigned long long nnzS;
    float maxS;             // max |S| after the last shrink pass
    int iter;               // iterations started (== reference iter_out)
    int sv, svp;            // sv = number of singular values the NEXT eig call looks at; svp = rank kept
    int sv_used;            // sv the last eig call looked at
    int svp_L;              // rank of the last COMPLETED iteration (what L is built from)
    int done;               // 0 running, 1 converged, 2 max_iter, 3 rank-0 break
    int converged;
    int max_iter, round005d, d, use_sv_prediction, break_on_rank0;
    int eig_info;           // diagnostics of the eigen solver (bisection rounds etc.)
    long long eig_clk[16];  // [0..5] clock64() at the phase boundaries of the last eig_kernel (CTA 0); [8..13] cycles per tridiagonalisation sub-phase
    // int8 (tcgen05) Gram path: W of the NEXT iteration is written by the shrink pass as 32-bit fixed point
    double wq_scale;        // S: q = rint(W * 2^31 / S), power of two
    double wq_scale_next;   // S for the slices the coming shrink pass writes
    double wmax;            // max |W| seen by the last shrink pass (over the W it produced)
    double wm_local;        // this rank's max |W_next| of the running iteration (-1: no slices written); never all-reduced
    int gram_mode;          // 0: fp64 DMMA Gram from D,S,Y   1: int8 tcgen05 Gram from the slices
    int wq_saturated;       // the last shrink pass clipped a slice -> fall back to the DMMA Gram once
    int use_i8;             // configuration: int8 path enabled
    // warm-started eigensolver (eig.cu): rows 0..eig_p-1 of Z hold orthonormal vectors of the previous iteration
    int eig_p;
    int eig_fast_iters;     // iterations of this solve that took the warm-started path
    double eig_gb;          // last certificate: ||G - X theta X^T||_F * mu^2  (must be < 1)
    // accuracy bookkeeping of the int8 Gram: once the bound of its truncation error (all-reduced with the Gram) stops being small
    // against the threshold (1/mu)^2, the rest of the solve uses the fp64 Gram
    double gram_err;        // last bound seen by the eigensolver, relative to (1/mu)^2
    int force_dmma;
    // shrink_flat.cu skips the store of S in iterations that cannot be the last one (S is not part of the recursion there: it
    // can be rebuilt from D, Y and the digit planes of W_next).  s_stale: S in HBM is older than the state.
    int s_stale, s_stale_next;
};

// which iterations the single-pass kernel of shrink_flat.cu takes (evaluated identically by it and by the kernels it relieves):
// the digit planes of this W exist (so project.cu has produced T), not the first iteration, rank <= 8
constexpr int kFlatMaxRank = 8;
__device__ __forceinline__ bool shrink_flat_takes(const DevState* st) {
    return st->gram_mode == 1 && st->iter > 1 && st->svp <= kFlatMaxRank;
}

struct IterLog {
    int iter, svp, sv, pad;
    double err, mu;
    unsigned long long nnz;
};

// Host-visible mirror (mapped pinned memory, written by the device control kernel).
struct HostMirror {
    volatile int iter;
    volatile int done;
    volatile int converged;
    volatile int svp;
    volatile double err;
};

#define BSUB_CUDA_CHECK(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            bsub::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)

void set_error(const char* fmt, ...);

// cudaFuncSetAttribute is per (function, device): remember per device whether a launcher has set its attributes
// (a benign race between host threads only repeats an idempotent call)
inline bool first_call_on_device(unsigned long long* seen) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (*seen & bit) return false;
    *seen |= bit;
    return true;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of doubles; result valid in thread 0. scratch: >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = (lane < nw) ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = -1e300;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = (lane < nw) ? scratch[lane] : -1e300;
        r = warp_max(r);
    }
    return r;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming 128-bit load that does not allocate in L1 (data is touched once per pass)
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace bsub
