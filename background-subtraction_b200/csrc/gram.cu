// gram.cu -- pass A of the ALM iteration: G = W W^T (frames x frames) with
//   W = D - S + Y/mu   formed on the fly in fp32   (/root/reference/inexact_alm_lsd.py:131)
// accumulated in fp64 on the FP64 tensor pipe (DMMA m8n8k4), which is what the Gram-route
// singular-value thresholding needs (the small eigenvalues near (1/mu)^2 sit ~1e-10
// below sigma_1^2, so products and sums must be fp64-grade).
//
// Work decomposition:
//   * the symmetric output is cut into 32x32 "macro tiles" (bi <= bj); one warp owns one macro tile
//     = 4x4 DMMA blocks = 32 fp64 accumulators per thread;
//   * a CTA has 14 MMA warps (14 macro tiles) + 2 producer warps; ceil(ntasks/14) CTA "types" cover
//     the whole upper triangle, CTAs of different type that share a k-slot walk the same pixel chunks
//     at the same time so the re-reads hit L2;
//   * K (= pixels) is split across k-slots; every CTA writes its partial tile to a scratch buffer and a
//     second kernel sums the k-slots in a fixed order (deterministic, no atomics).
#include "common.cuh"
#include "kernels.h"

namespace bsub {

constexpr int GR_KC = 32;           // pixels per load step (128 B per frame row)
// The W tile is kept in shared memory as fp64 (converted once by the producer warps): F2F.F64.F32 runs at
// 1/4 rate, and every fragment is read by up to 14 warps, so converting at fragment-load time made the kernel
// conversion-bound (ncu r1a: DMMA pipe 35 % active).  Row stride kc+4 doubles = 4 (mod 16): the 16 lanes of a
// half-warp (r = 0..3, c = 0..3) hit 16 distinct 8-byte banks.
constexpr int GR_MMA_WARPS = 14;
constexpr int GR_PROD_WARPS = 2;
constexpr int GR_THREADS = 32 * (GR_MMA_WARPS + GR_PROD_WARPS);

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct GramArgs {
    const float* D; const float* S; const float* Y;   // S == nullptr -> W = D (init pass)
    long long ld;
    int n, npad, ntasks, ntype, gridK, kc, lds;        // kc = pixels per stage (32, 16 or 8), lds = kc + 4
    long long nchunks;                                  // ld / kc
    const int2* tasks;                                  // (bi, bj), bi <= bj
    const DevState* st;                                 // may be nullptr (stand-alone use)
    float inv_mu_override;                              // used when st == nullptr
    double* partial;                                    // [gridK][ntasks][32*32]
};

// Fill one stage: rows = frames (zero beyond n), kc pixels each, stored as fp64.
__device__ __forceinline__ void gram_fill_stage(double* buf, const GramArgs& a, long long chunk, float inv_mu, int tp) {
    const int qpr = a.kc >> 2;            // float4 slots per row (8, 4 or 2)
    const int q = tp % qpr;
    const int r0 = tp / qpr;
    const int rstep = 64 / qpr;           // rows covered by the 64 producer threads per pass
    const long long p0 = chunk * a.kc + 4 * q;
    const bool combo = (a.S != nullptr);
    for (int f = r0; f < a.npad; f += 4 * rstep) {
        float4 d[4], s[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int ff = f + rstep * u;
            d[u] = make_float4(0.f, 0.f, 0.f, 0.f); s[u] = d[u]; y[u] = d[u];
            if (ff < a.n) {
                long long off = (long long)ff * a.ld + p0;
                d[u] = ldg4_stream(a.D + off);
                if (combo) { s[u] = ldg4_stream(a.S + off); y[u] = ldg4_stream(a.Y + off); }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int ff = f + rstep * u;
            if (ff < a.npad) {
                // same fp32 expression as the shrink pass so that both passes see the same W
                const float wx = (d[u].x - s[u].x) + y[u].x * inv_mu;
                const float wy = (d[u].y - s[u].y) + y[u].y * inv_mu;
                const float wz = (d[u].z - s[u].z) + y[u].z * inv_mu;
                const float ww = (d[u].w - s[u].w) + y[u].w * inv_mu;
                double* dst = buf + (size_t)ff * a.lds + 4 * q;
                *reinterpret_cast<double2*>(dst) = make_double2((double)wx, (double)wy);
                *reinterpret_cast<double2*>(dst + 2) = make_double2((double)wz, (double)ww);
            }
        }
    }
}

__global__ void __launch_bounds__(GR_THREADS, 1) gram_dmma_kernel(GramArgs a) {
    if (a.st != nullptr && a.st->done) return;
    extern __shared__ __align__(16) double gram_smem[];
    double* bufs[2] = {gram_smem, gram_smem + (size_t)a.npad * a.lds};

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int type = blockIdx.x % a.ntype;
    const int kslot = blockIdx.x / a.ntype;
    const bool producer = warp >= GR_MMA_WARPS;
    const int task = type * GR_MMA_WARPS + warp;
    const bool has_task = !producer && task < a.ntasks;
    int bi = 0, bj = 0;
    if (has_task) { int2 t = a.tasks[task]; bi = t.x; bj = t.y; }
    const bool diag = (bi == bj);
    float inv_mu = 0.f;
    if (a.S != nullptr) inv_mu = (a.st != nullptr) ? (float)(1.0 / a.st->mu) : a.inv_mu_override;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    long long niter = 0;
    if (kslot < a.nchunks) niter = (a.nchunks - kslot + a.gridK - 1) / a.gridK;
    const int tp = threadIdx.x - GR_MMA_WARPS * 32;   // producer thread id 0..63

    if (producer && niter > 0) gram_fill_stage(bufs[0], a, kslot, inv_mu, tp);
    __syncthreads();
    const int fr = lane >> 2, fc = lane & 3;
    for (long long it = 0; it < niter; ++it) {
        const int cur = (int)(it & 1);
        if (producer) {
            if (it + 1 < niter) gram_fill_stage(bufs[cur ^ 1], a, kslot + (it + 1) * a.gridK, inv_mu, tp);
        } else if (has_task) {
            const double* rowA = bufs[cur] + (size_t)(bi * 32 + fr) * a.lds + fc;
            const double* rowB = bufs[cur] + (size_t)(bj * 32 + fr) * a.lds + fc;
            const int nks = a.kc >> 2, rs8 = 8 * a.lds;
#pragma unroll 2
            for (int ks = 0; ks < nks; ++ks) {
                double af[4], bf[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    af[i] = rowA[i * rs8 + ks * 4];
                    bf[i] = rowB[i * rs8 + ks * 4];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (!diag || i <= j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        __syncthreads();
    }
    if (has_task) {
        double* out = a.partial + ((size_t)kslot * a.ntasks + task) * 1024;
        const int row = lane >> 2, col = 2 * (lane & 3);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
                *reinterpret_cast<double2*>(out + (i * 8 + row) * 32 + j * 8 + col) = v;
            }
    }
}

// Sum the k-slot partials in a fixed order and scatter the symmetric result into G[npad][npad].
__global__ void gram_reduce_kernel(const double* __restrict__ partial, const int2* __restrict__ tasks, int ntasks,
                                   int gridK, int npad, double* __restrict__ G, const DevState* st) {
    if (st != nullptr && st->done) return;
    const int task = blockIdx.x;
    const int2 t = tasks[task];
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) {
        const int r = e >> 5, c = e & 31;
        if (t.x == t.y && r > c) continue;
        double s = 0.0;
        for (int k = 0; k < gridK; ++k) s += partial[((size_t)k * ntasks + task) * 1024 + e];
        const int i = t.x * 32 + r, j = t.y * 32 + c;
        G[(size_t)i * npad + j] = s;
        G[(size_t)j * npad + i] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------
GramPlan make_gram_plan(int n, long long ld, int num_sms) {
    GramPlan p;
    p.n = n;
    p.npad = ((n + 31) / 32) * 32;
    p.nb = p.npad / 32;
    p.ntasks = p.nb * (p.nb + 1) / 2;
    p.ntype = (p.ntasks + GR_MMA_WARPS - 1) / GR_MMA_WARPS;
    // largest stage width whose two fp64 stages fit in shared memory
    p.kc = GR_KC;
    while (p.kc > 8 && (size_t)2 * p.npad * (p.kc + 4) * sizeof(double) > 200 * 1024) p.kc >>= 1;
    p.nchunks = ld / p.kc;
    long long gk = num_sms / p.ntype;
    if (gk < 1) gk = 1;
    if (gk > p.nchunks) gk = p.nchunks;
    if (gk < 1) gk = 1;
    p.gridK = (int)gk;
    p.smem_bytes = (size_t)2 * p.npad * (p.kc + 4) * sizeof(double);
    p.partial_elems = (size_t)p.gridK * p.ntasks * 1024;
    return p;
}

void fill_gram_tasks(const GramPlan& p, int2* host_tasks) {
    // order tasks so that the macro tiles of one CTA type share frame rows as much as possible
    int t = 0;
    for (int bi = 0; bi < p.nb; ++bi)
        for (int bj = bi; bj < p.nb; ++bj) host_tasks[t++] = make_int2(bi, bj);
}

int launch_gram(const GramPlan& p, const float* D, const float* S, const float* Y, long long ld,
                const int2* dev_tasks, const DevState* st, float inv_mu_override, double* partial, double* G,
                cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        BSUB_CUDA_CHECK(cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (p.smem_bytes > 200 * 1024) { set_error("gram: n=%d too large for the shared-memory W tile", p.n); return -1; }
    GramArgs a;
    a.D = D; a.S = S; a.Y = Y; a.ld = ld; a.n = p.n; a.npad = p.npad; a.ntasks = p.ntasks; a.ntype = p.ntype;
    a.gridK = p.gridK; a.kc = p.kc; a.lds = p.kc + 4; a.nchunks = p.nchunks; a.tasks = dev_tasks; a.st = st; a.inv_mu_override = inv_mu_override;
    a.partial = partial;
    gram_dmma_kernel<<<p.gridK * p.ntype, GR_THREADS, p.smem_bytes, stream>>>(a);
    BSUB_CUDA_CHECK(cudaGetLastError());
    gram_reduce_kernel<<<p.ntasks, 256, 0, stream>>>(partial, dev_tasks, p.ntasks, p.gridK, p.npad, G, st);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace bsub
