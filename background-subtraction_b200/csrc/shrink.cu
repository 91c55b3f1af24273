// shrink.cu -- pass B of the ALM iteration, one fused bandwidth-bound kernel (20 B per matrix element):
//
//   T   = Vr^T W,  W = D - S + Y/mu                    (projection on the kept right singular vectors)
//   L   = VC T      = U (sigma - 1/mu) V^T             (/root/reference/inexact_alm_lsd.py:147, utils.py:185-186)
//   G_S = D - L + Y/mu                                 (:150)
//   S   = prox_{lambda/mu}(G_S)   over the 3x3 pixel tiles of each frame       (:153-155, prox_flat :71-79;
//                                  closed form: clip |G_S| at the level theta of the tile's l1-ball projection)
//   Z   = D - L - S ;  Y += mu Z ;  sum Z^2            (:162-167)
//
// Work decomposition: a tile is 3 image columns x R rows (whole 3x3 groups) of ALL
// frames.  A thread-block cluster of Cf CTAs owns a tile, CTA `rank` handling frames [rank*nf, rank*nf+nf):
// it streams its D,S,Y slice in once (float4, coalesced), keeps D and Y in shared memory, accumulates its
// partial T, the Cf partials are exchanged through L2 around ONE hardware cluster barrier, and everything after
// that (L, prox, dual update, stores) is local to the CTA.  The only HBM traffic is read D,S,Y + write S,Y.
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace bsub {

constexpr int SH_KC = 8;      // singular vectors handled per register chunk
constexpr int SH_KR = 4;      // vectors per cross-warp reduction round
constexpr int SH_THREADS = 256;

struct ShrinkArgs {
    const float* D; float* S; float* Y; float* T; float* U;
    float* tpart; const float* Vr; const float* VC; int vstride;
    long long ld, m;
    int n, rows, cols, R, P, NQ, NFL, Cf, nf;
    int ntile_r; long long ntiles; int nclusters;
    const DevState* st;
    double* part_zz; unsigned long long* part_nnz; float* part_max;
    int mode, vec;
};

__device__ __forceinline__ void cswap_desc(float& a, float& b) {
    float hi = fmaxf(a, b), lo = fminf(a, b);
    a = hi; b = lo;
}

// Clip level of the l1-ball projection: theta >= 0 with sum_i max(a_i - theta, 0) = z, for a_i >= 0, sum a > z.
__device__ __forceinline__ float clip_level9(const float* a_in, float z) {
    float u[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) u[i] = a_in[i];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8 - i; ++j) cswap_desc(u[j], u[j + 1]);
    const float inv[9] = {1.f, 0.5f, 1.f / 3.f, 0.25f, 0.2f, 1.f / 6.f, 1.f / 7.f, 0.125f, 1.f / 9.f};
    float cs = 0.f, theta = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        cs += u[k];
        float t = (cs - z) * inv[k];
        if (u[k] > t) theta = t;
    }
    return theta;
}

struct Quad { long long p; int valid; };   // first pixel index, number of valid rows (0..4), -1 = column outside

__global__ void __launch_bounds__(SH_THREADS, 2) shrink_kernel(ShrinkArgs a) {
    const DevState* st = a.st;
    if (st->done) return;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int rank = blockIdx.x % a.Cf, cl = blockIdx.x / a.Cf;
    const int f0 = rank * a.nf;
    const int nfr = max(0, min(a.nf, a.n - f0));
    const int P = a.P, NQ = a.NQ, NFL = a.NFL, R = a.R;
    const int r = st->svp;
    const double mu_d = st->mu;
    const float inv_mu = (float)(1.0 / mu_d), mu_f = (float)mu_d;
    const float lamq = (float)(st->lambda / mu_d);
    const int nchunk = (r + SH_KC - 1) / SH_KC;

    extern __shared__ __align__(16) float ssm[];
    float* Ds = ssm;                               // [nf][P]
    float* Ys = Ds + (size_t)a.nf * P;             // [nf][P]
    float* scr = Ys + (size_t)a.nf * P;            // [NFL][SH_KR][P]
    float* Tp = scr + (size_t)NFL * SH_KR * P;     // [SH_KC][P]
    float* Vs = Tp + (size_t)SH_KC * P;            // [nf][SH_KC]
    __shared__ double redd[32];

    const int qd = tid % NQ, fl = tid / NQ;
    const bool tact = fl < NFL;
    const int RQ = R / 4;
    const int qc = qd / RQ, qi = (qd - qc * RQ) * 4;

    double zz_acc = 0.0;
    unsigned long long nnz_acc = 0ull;
    float max_acc = 0.f;

    int slot = 0;
    for (long long tl = cl; tl < a.ntiles; tl += a.nclusters, slot ^= 1) {
        const int tcx = (int)(tl / a.ntile_r), trx = (int)(tl - (long long)tcx * a.ntile_r);
        const int j0 = 3 * tcx, i0 = trx * R;
        // this thread's quad
        const int qj = j0 + qc, qrow = i0 + qi;
        int qvalid = 0;
        if (qj < a.cols) qvalid = max(0, min(4, a.rows - qrow));
        const long long qp = (long long)qj * a.rows + qrow;
        float* tslot = a.tpart + (((size_t)slot * a.nclusters + cl) * a.Cf) * (size_t)a.n * P;   // [Cf][n][P]

        // ---------------- sweep 1: stream D,S,Y in, stage D and Y, accumulate partial T ----------------
        for (int kc = 0; kc < max(nchunk, 1); ++kc) {
            const int k0 = kc * SH_KC;
            if (r > 0) {
                for (int idx = tid; idx < nfr * SH_KC; idx += SH_THREADS) {
                    int f = idx / SH_KC, k = idx - f * SH_KC;
                    Vs[idx] = (k0 + k < r) ? a.Vr[(size_t)(f0 + f) * a.vstride + k0 + k] : 0.f;
                }
            }
            __syncthreads();
            float acc[SH_KC][4];
#pragma unroll
            for (int k = 0; k < SH_KC; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
            if (tact) {
#pragma unroll 2
                for (int f = fl; f < nfr; f += NFL) {
                    const long long off = (long long)(f0 + f) * a.ld + qp;
                    float4 d4, y4, s4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (kc == 0) {
                        d4 = s4; y4 = s4;
                        if (a.vec) {
                            if (qvalid == 4) { d4 = ldg4_stream(a.D + off); y4 = ldg4_stream(a.Y + off); s4 = ldg4_stream(a.S + off); }
                        } else {
                            float dd[4] = {0, 0, 0, 0}, yy[4] = {0, 0, 0, 0}, sx[4] = {0, 0, 0, 0};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (e < qvalid) { dd[e] = a.D[off + e]; yy[e] = a.Y[off + e]; sx[e] = a.S[off + e]; }
                            d4 = make_float4(dd[0], dd[1], dd[2], dd[3]);
                            y4 = make_float4(yy[0], yy[1], yy[2], yy[3]);
                            s4 = make_float4(sx[0], sx[1], sx[2], sx[3]);
                        }
                        *reinterpret_cast<float4*>(Ds + (size_t)f * P + 4 * qd) = d4;
                        *reinterpret_cast<float4*>(Ys + (size_t)f * P + 4 * qd) = y4;
                    } else {
                        d4 = *reinterpret_cast<const float4*>(Ds + (size_t)f * P + 4 * qd);
                        y4 = *reinterpret_cast<const float4*>(Ys + (size_t)f * P + 4 * qd);
                        if (a.vec) { if (qvalid == 4) s4 = ldg4(a.S + off); }
                        else {
                            float sx[4] = {0, 0, 0, 0};
#pragma unroll
                            for (int e = 0; e < 4; ++e) if (e < qvalid) sx[e] = a.S[off + e];
                            s4 = make_float4(sx[0], sx[1], sx[2], sx[3]);
                        }
                    }
                    if (r > 0) {
                        float4 w;
                        w.x = (d4.x - s4.x) + y4.x * inv_mu;
                        w.y = (d4.y - s4.y) + y4.y * inv_mu;
                        w.z = (d4.z - s4.z) + y4.z * inv_mu;
                        w.w = (d4.w - s4.w) + y4.w * inv_mu;
                        const float4 va = *reinterpret_cast<const float4*>(Vs + f * SH_KC);
                        const float4 vb = *reinterpret_cast<const float4*>(Vs + f * SH_KC + 4);
                        const float vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
                        for (int k = 0; k < SH_KC; ++k) {
                            acc[k][0] = fmaf(vv[k], w.x, acc[k][0]);
                            acc[k][1] = fmaf(vv[k], w.y, acc[k][1]);
                            acc[k][2] = fmaf(vv[k], w.z, acc[k][2]);
                            acc[k][3] = fmaf(vv[k], w.w, acc[k][3]);
                        }
                    }
                }
            }
            if (r > 0) {
                // cross-warp reduction over the NFL frame lanes, SH_KR vectors per round
#pragma unroll
                for (int kr0 = 0; kr0 < SH_KC; kr0 += SH_KR) {
                    if (tact) {
#pragma unroll
                        for (int k = 0; k < SH_KR; ++k)
                            *reinterpret_cast<float4*>(scr + ((size_t)(fl * SH_KR + k)) * P + 4 * qd) =
                                make_float4(acc[kr0 + k][0], acc[kr0 + k][1], acc[kr0 + k][2], acc[kr0 + k][3]);
                    }
                    __syncthreads();
                    for (int idx = tid; idx < SH_KR * NQ; idx += SH_THREADS) {
                        const int k = idx / NQ, q = idx - k * NQ;
                        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int l = 0; l < NFL; ++l) {
                            const float4 v = *reinterpret_cast<const float4*>(scr + ((size_t)(l * SH_KR + k)) * P + 4 * q);
                            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                        }
                        // publish this CTA's partial for vector k0+kr0+k
                        if (k0 + kr0 + k < r)
                            *reinterpret_cast<float4*>(tslot + ((size_t)rank * a.n + (k0 + kr0 + k)) * P + 4 * q) = s;
                    }
                    __syncthreads();
                }
            }
        }
        // ---------------- exchange: one cluster barrier, partials through L2 ----------------
        cluster.sync();

        // ---------------- sweep 2a: a = D - L in place ----------------
        for (int kc = 0; kc < nchunk; ++kc) {
            const int k0 = kc * SH_KC;
            for (int idx = tid; idx < SH_KC * NQ; idx += SH_THREADS) {
                const int k = idx / NQ, q = idx - k * NQ;
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + k < r) {
                    for (int rk = 0; rk < a.Cf; ++rk) {
                        const float4 v = __ldcg(reinterpret_cast<const float4*>(tslot + ((size_t)rk * a.n + (k0 + k)) * P + 4 * q));
                        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                    }
                    if (rank == 0) {      // keep T for the final materialisation of L
                        const int c2 = q / RQ, i2 = (q - c2 * RQ) * 4;
                        const int j2 = j0 + c2, row2 = i0 + i2;
                        if (j2 < a.cols) {
                            const int nv = max(0, min(4, a.rows - row2));
                            float* tp = a.T + (size_t)(k0 + k) * a.ld + (long long)j2 * a.rows + row2;
                            if (a.vec) { if (nv == 4) stg4(tp, s); }
                            else { const float sv[4] = {s.x, s.y, s.z, s.w}; for (int e = 0; e < nv; ++e) tp[e] = sv[e]; }
                        }
                    }
                }
                *reinterpret_cast<float4*>(Tp + (size_t)k * P + 4 * q) = s;
            }
            for (int idx = tid; idx < nfr * SH_KC; idx += SH_THREADS) {
                int f = idx / SH_KC, k = idx - f * SH_KC;
                Vs[idx] = (k0 + k < r) ? a.VC[(size_t)(f0 + f) * a.vstride + k0 + k] : 0.f;
            }
            __syncthreads();
            if (tact) {
                float4 t4[SH_KC];
#pragma unroll
                for (int k = 0; k < SH_KC; ++k) t4[k] = *reinterpret_cast<const float4*>(Tp + (size_t)k * P + 4 * qd);
                for (int f = fl; f < nfr; f += NFL) {
                    const float4 va = *reinterpret_cast<const float4*>(Vs + f * SH_KC);
                    const float4 vb = *reinterpret_cast<const float4*>(Vs + f * SH_KC + 4);
                    const float vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
                    float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < SH_KC; ++k) {
                        l.x = fmaf(vv[k], t4[k].x, l.x); l.y = fmaf(vv[k], t4[k].y, l.y);
                        l.z = fmaf(vv[k], t4[k].z, l.z); l.w = fmaf(vv[k], t4[k].w, l.w);
                    }
                    float4* dp = reinterpret_cast<float4*>(Ds + (size_t)f * P + 4 * qd);
                    float4 d4 = *dp;
                    d4.x -= l.x; d4.y -= l.y; d4.z -= l.z; d4.w -= l.w;
                    *dp = d4;
                }
            }
            __syncthreads();
        }
        if (nchunk == 0) __syncthreads();

        // ---------------- sweep 2b: prox + dual update, in shared memory ----------------
        if (a.mode == SHRINK_FLAT3) {
            const int NG = R / 3;
            for (int it = tid; it < nfr * NG; it += SH_THREADS) {
                const int f = it / NG, g = it - f * NG;
                float* dsp = Ds + (size_t)f * P + 3 * g;
                float* ysp = Ys + (size_t)f * P + 3 * g;
                float av[9], yv[9], x[9], ax[9];
                float sabs = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int dr = 0; dr < 3; ++dr) {
                        const int e = c * 3 + dr;
                        av[e] = dsp[c * R + dr];
                        yv[e] = ysp[c * R + dr];
                        x[e] = av[e] + yv[e] * inv_mu;          // G_S
                        ax[e] = fabsf(x[e]);
                        sabs += ax[e];
                    }
                float theta = 0.f;
                const bool nz = sabs > lamq;
                if (nz) theta = clip_level9(ax, lamq);
                float zl = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int dr = 0; dr < 3; ++dr) {
                        const int e = c * 3 + dr;
                        const float s = nz ? copysignf(fminf(ax[e], theta), x[e]) : 0.f;
                        const float z = av[e] - s;                 // Z = D - L - S
                        dsp[c * R + dr] = s;
                        ysp[c * R + dr] = fmaf(mu_f, z, yv[e]);    // Y += mu Z
                        zl = fmaf(z, z, zl);
                        nnz_acc += (s != 0.f);
                        max_acc = fmaxf(max_acc, fabsf(s));
                    }
                zz_acc += (double)zl;
            }
        } else if (a.mode == SHRINK_L1) {
            for (int it = tid; it < nfr * P; it += SH_THREADS) {
                const float av = Ds[it], yv = Ys[it];
                const float x = av + yv * inv_mu;
                const float s = copysignf(fmaxf(fabsf(x) - lamq, 0.f), x);
                const float z = av - s;
                Ds[it] = s; Ys[it] = fmaf(mu_f, z, yv);
                zz_acc += (double)(z * z);
                nnz_acc += (s != 0.f);
                max_acc = fmaxf(max_acc, fabsf(s));
            }
        } else {   // SHRINK_SPILL: hand G_S to a separate prox, dual update happens in launch_dual_update
            for (int it = tid; it < nfr * P; it += SH_THREADS) Ds[it] = Ds[it] + Ys[it] * inv_mu;
        }
        __syncthreads();

        // ---------------- sweep 2c: coalesced stores ----------------
        if (tact) {
            float* out0 = (a.mode == SHRINK_SPILL) ? a.U : a.S;
            for (int f = fl; f < nfr; f += NFL) {
                const long long off = (long long)(f0 + f) * a.ld + qp;
                const float4 s4 = *reinterpret_cast<const float4*>(Ds + (size_t)f * P + 4 * qd);
                const float4 y4 = *reinterpret_cast<const float4*>(Ys + (size_t)f * P + 4 * qd);
                if (a.vec) {
                    if (qvalid == 4) { stg4(out0 + off, s4); if (a.mode != SHRINK_SPILL) stg4(a.Y + off, y4); }
                } else {
                    const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
                    for (int e = 0; e < qvalid; ++e) { out0[off + e] = sv[e]; if (a.mode != SHRINK_SPILL) a.Y[off + e] = yv[e]; }
                }
            }
        }
        __syncthreads();
    }

    // ---------------- per-CTA partial statistics (summed in fixed order by control_post) ----------------
    double zt = block_sum(zz_acc, redd);
    if (tid == 0) a.part_zz[blockIdx.x] = zt;
    double nt = block_sum((double)nnz_acc, redd);
    if (tid == 0) a.part_nnz[blockIdx.x] = (unsigned long long)(nt + 0.5);
    double mt = block_max((double)max_acc, redd);
    if (tid == 0) a.part_max[blockIdx.x] = (float)mt;
}

// -------------------------------------------------------------------------------------------------------------
ShrinkPlan make_shrink_plan(int n, int rows, int cols, long long ld, int num_sms, int R_hint, int Cf_hint) {
    ShrinkPlan p;
    p.n = n; p.rows = rows; p.cols = cols; p.ld = ld; p.m = (long long)rows * cols;
    p.threads = SH_THREADS; p.kr = SH_KR;
    // tile rows: multiple of 12 (whole 3x3 groups, whole float4s)
    int R = R_hint > 0 ? R_hint : 48;
    R = ((R + 11) / 12) * 12;
    int rows12 = ((rows + 11) / 12) * 12;
    if (R > rows12) R = rows12;
    if (R > 252) R = 252;           // NQ = 3R/4 must stay <= SH_THREADS
    auto smem_for = [&](int Rr, int Cf) {
        int P = 3 * Rr, NQ = P / 4, NFL = SH_THREADS / NQ;
        int nf = (n + Cf - 1) / Cf;
        size_t fl = (size_t)2 * nf * P + (size_t)NFL * SH_KR * P + (size_t)SH_KC * P + (size_t)nf * SH_KC;
        return fl * sizeof(float);
    };
    int Cf = Cf_hint > 0 ? Cf_hint : 1;
    if (Cf_hint <= 0) {
        // aim for <= ~72 KB per CTA (3 CTAs per SM); clusters of up to 8 CTAs split the frames
        while (Cf < 8 && smem_for(R, Cf) > 72 * 1024) Cf *= 2;
    }
    while (smem_for(R, Cf) > 110 * 1024 && R > 12) R -= 12;   // last resort: narrower tiles
    p.R = R; p.P = 3 * R; p.Cf = Cf; p.nf = (n + Cf - 1) / Cf;
    p.smem_bytes = smem_for(R, Cf);
    p.ntile_r = (rows + R - 1) / R;
    p.ntile_c = (cols + 2) / 3;
    p.ntiles = (long long)p.ntile_r * p.ntile_c;
    int occ = (int)((220 * 1024) / (p.smem_bytes + 1024));
    if (occ < 1) occ = 1;
    if (occ > 4) occ = 4;
    long long ncl = ((long long)num_sms * occ) / Cf;
    if (ncl < 1) ncl = 1;
    if (ncl > p.ntiles) ncl = p.ntiles;
    p.grid_clusters = (int)ncl;
    p.nparts = p.grid_clusters * Cf;
    p.tpart_floats = (size_t)2 * p.grid_clusters * Cf * (size_t)n * p.P;
    return p;
}

int launch_shrink(const ShrinkPlan& p, ShrinkBuffers b, const DevState* st, int mode, cudaStream_t stream) {
    static unsigned long long attr_devs = 0;      // one bit per device: the attribute is per (function, device)
    if (first_call_on_device(&attr_devs)) {
        BSUB_CUDA_CHECK(cudaFuncSetAttribute(shrink_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    if (p.smem_bytes > 200 * 1024) { set_error("shrink: tile does not fit shared memory (n=%d)", p.n); return -1; }
    ShrinkArgs a;
    a.D = b.D; a.S = b.S; a.Y = b.Y; a.T = b.T; a.U = b.U; a.tpart = b.tpart; a.Vr = b.Vr; a.VC = b.VC;
    a.vstride = b.vstride; a.ld = p.ld; a.m = p.m; a.n = p.n; a.rows = p.rows; a.cols = p.cols; a.R = p.R; a.P = p.P;
    a.NQ = p.P / 4; a.NFL = SH_THREADS / a.NQ; a.Cf = p.Cf; a.nf = p.nf; a.ntile_r = p.ntile_r; a.ntiles = p.ntiles;
    a.nclusters = p.grid_clusters; a.st = st; a.part_zz = b.part_zz; a.part_nnz = b.part_nnz; a.part_max = b.part_max;
    a.mode = mode; a.vec = (p.rows % 4 == 0) ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.grid_clusters * p.Cf); cfg.blockDim = dim3(SH_THREADS);
    cfg.dynamicSmemBytes = p.smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = p.Cf; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    BSUB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, shrink_kernel, a));
    return 0;
}

}  // namespace bsub
