// elementwise.cu -- the streaming kernels around the two big passes: init norms, Y0, per-iteration control,
// dtype/layout conversion at the boundary, materialisation of L, and the second phase of the two-phase shrink.
#include "common.cuh"
#include "kernels.h"

namespace bsub {

constexpr int EW_THREADS = 256;

static inline int ew_grid(long long work, int per_block, int cap = 148 * 8) {
    long long g = (work + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// ---- max_p sum_f |D[p,f]|  (NumPy's induced inf-norm of the m x n matrix; /root/reference/inexact_alm_lsd.py:110,
//      SURVEY Q1).  comm_max[0] must be zero on entry; non-negative doubles order like their bit patterns.
__global__ void rowsum_max_kernel(const float* __restrict__ D, long long ld, long long m, int n, double* comm_max) {
    __shared__ double red[32];
    double best = 0.0, dmax = 0.0;
    const long long nq = ld / 4;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int f = 0; f < n; ++f) {
            const float4 d = ldg4_stream(D + (size_t)f * ld + 4 * q);
            acc.x += fabsf(d.x); acc.y += fabsf(d.y); acc.z += fabsf(d.z); acc.w += fabsf(d.w);
            dmax = fmax(dmax, (double)fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
            if ((f & 63) == 63) { a0 += acc.x; a1 += acc.y; a2 += acc.z; a3 += acc.w; acc = make_float4(0.f, 0.f, 0.f, 0.f); }
        }
        a0 += acc.x; a1 += acc.y; a2 += acc.z; a3 += acc.w;
        best = fmax(best, fmax(fmax(a0, a1), fmax(a2, a3)));
    }
    best = block_max(best, red);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(comm_max), (unsigned long long)__double_as_longlong(best));
    dmax = block_max(dmax, red);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(comm_max + 2), (unsigned long long)__double_as_longlong(dmax));
}

int launch_rowsum_max(const float* D, long long ld, long long m, int n, double* comm_max, cudaStream_t s) {
    rowsum_max_kernel<<<ew_grid(ld / 4, EW_THREADS), EW_THREADS, 0, s>>>(D, ld, m, n, comm_max);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- Y0 = D / max(||D||_2, rowsum/lambda),  S0 = 0   (/root/reference/inexact_alm_lsd.py:108-120)
__global__ void init_Y_kernel(const float* __restrict__ D, float* __restrict__ Y, float* __restrict__ S, long long total4,
                              const DevState* st) {
    const double inv = 1.0 / st->dual_norm;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (long long)gridDim.x * blockDim.x) {
        const float4 d = ldg4_stream(D + 4 * q);
        float4 y;
        y.x = (float)((double)d.x * inv); y.y = (float)((double)d.y * inv);
        y.z = (float)((double)d.z * inv); y.w = (float)((double)d.w * inv);
        stg4(Y + 4 * q, y);
        stg4(S + 4 * q, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

int launch_init_Y(const float* D, float* Y, float* S, long long ld, int n, const DevState* st, cudaStream_t s) {
    const long long total4 = ld * n / 4;
    init_Y_kernel<<<ew_grid(total4, EW_THREADS * 4), EW_THREADS, 0, s>>>(D, Y, S, total4, st);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- per-iteration control after the shrink pass (/root/reference/inexact_alm_lsd.py:164-177).
// phase bit 1: gather this rank's per-CTA partials (fixed order) into the communication buffer
//              comm_tail[0] = sum Z^2, comm_tail[1] = nnz(S), comm_tail[2] = max |S| (local)
// phase bit 4: advance mu and the digit-plane bookkeeping (local quantities only).
// phase bit 2: finish the iteration from the (all-reduced) communication buffer: err, log line, stop flags.
__global__ void control_post_kernel(DevState* st, const double* part_zz, const unsigned long long* part_nnz,
                                    const float* part_max, int nparts, double* comm_tail, IterLog* log,
                                    HostMirror* mirror, int phase, const float* part_wmax, int nwmax) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    if (!st->done && (phase & 1)) {
        // one warp: lane-strided partial sums, then a fixed shuffle tree (deterministic for a given launch geometry)
        double zz = 0.0, nz = 0.0; float mx = 0.f;
        for (int i = lane; i < nparts; i += 32) { zz += part_zz[i]; nz += (double)part_nnz[i]; mx = fmaxf(mx, part_max[i]); }
        // int8 Gram bookkeeping: were slices of the next W written (no negative marker), and with head-room?
        float wm = 0.f; int missing = 0;
        const bool track = st->use_i8 && part_wmax != nullptr && nwmax > 0;
        if (track)
            for (int i = lane; i < nwmax; i += 32) { const float v = part_wmax[i]; missing |= (v < 0.f); wm = fmaxf(wm, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            zz += __shfl_xor_sync(0xffffffffu, zz, o); nz += __shfl_xor_sync(0xffffffffu, nz, o);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
            missing |= __shfl_xor_sync(0xffffffffu, missing, o);
        }
        if (lane == 0) {
            comm_tail[0] = zz; comm_tail[1] = nz; comm_tail[2] = (double)mx;
            // a clipped digit pass while S was not being stored leaves this rank with neither valid planes nor a current S: the
            // flag travels with the other scalars (summed over the ranks), so every rank stops with done == 5 in the same
            // iteration and the drivers repeat the solve with S stored every time (solver.cu bsub_run, dist.ShardedLSD)
            comm_tail[3] = (track && !missing && wm >= 1.0e38f && st->s_stale_next) ? 1.0 : 0.0;
            // scale and validity of the slices are per rank (every rank turns its own partial Gram into doubles), so this
            // stays out of the all-reduced tail
            st->wm_local = (track && !missing) ? (double)wm : -1.0;
        }
    }
    if (lane != 0) return;
    if (!st->done) {
        if (phase & 4) {
            // bookkeeping that does not need the reduced residual: mu <- rho mu (inexact_alm_lsd.py:164) and the fixed-point
            // scale of the digit planes.  A pixel-sharded run enqueues the next Gram right after this, so that the 4 scalars of
            // this iteration travel in the same all-reduce message as the next Gram.
            st->mu_iter = st->mu;
            st->mu = fmin(st->mu * st->rho, st->mu * 1e7);
            st->s_stale = st->s_stale_next; st->s_stale_next = 0;      // did this iteration's shrink pass leave S in HBM untouched?
            if (st->use_i8) {
                const double wm = st->wm_local;
                if (wm >= 0.0 && wm < 1.0e38 && wm > 0.0) {          // slices written, nothing clipped
                    st->gram_mode = 1; st->wq_saturated = 0;
                    st->wq_scale = st->wq_scale_next;              // scale of the slices that now exist
                    st->wq_scale_next = exp2(ceil(log2(4.0 * wm))); // head-room x4 for the W of the following pass
                    st->wmax = wm;
                } else {
                    st->gram_mode = 0; st->wq_saturated = (wm >= 1.0e38);
                    if (wm >= 1.0e38) st->wq_scale_next = st->wq_scale_next * 16.0;
                }
                if (st->force_dmma) st->gram_mode = 0;            // the int8 Gram has become too coarse for the shrinking threshold (eig.cu)

            }
        }
        if (phase & 2) {
            const double zz = comm_tail[0];
            const double err = sqrt(zz / st->normD2);
            st->zz = zz; st->err = err;
            st->nnzS = (unsigned long long)(comm_tail[1] + 0.5);
            st->maxS = (float)comm_tail[2];
            st->svp_L = st->svp;
            const int it = st->iter;
            if (it >= 1 && it <= kMaxIterLog) {
                IterLog& l = log[it - 1];
                l.iter = it; l.svp = st->svp; l.sv = st->sv_used; l.pad = 0; l.err = err; l.mu = st->mu_iter; l.nnz = st->nnzS;
            }
            if (comm_tail[3] > 0.0) { st->done = 5; }               // unrecoverable digit saturation somewhere: restart (see phase 1)
            else if (err < st->tol) { st->done = 1; st->converged = 1; }
            else if (it >= st->max_iter) { st->done = 2; }
        }
    }
    if (mirror != nullptr && (phase & 2)) {
        mirror->iter = st->iter; mirror->converged = st->converged; mirror->svp = st->svp; mirror->err = st->err;
        __threadfence_system();
        mirror->done = st->done;
        __threadfence_system();
    }
}

int launch_control_post(DevState* st, const double* part_zz, const unsigned long long* part_nnz, const float* part_max,
                        int nparts, double* comm_tail, IterLog* log, HostMirror* mirror, int phase, const float* part_wmax, int nwmax,
                        cudaStream_t s) {
    control_post_kernel<<<1, 32, 0, s>>>(st, part_zz, part_nnz, part_max, nparts, comm_tail, log, mirror, phase, part_wmax, nwmax);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- boundary conversions ---------------------------------------------------------------------------------
__global__ void convert_f64_kernel(const double* __restrict__ src, long long src_ld, float* __restrict__ dst, long long ld,
                                   long long m, int n) {
    for (int f = blockIdx.y; f < n; f += gridDim.y)
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < ld; p += (long long)gridDim.x * blockDim.x)
            dst[(size_t)f * ld + p] = (p < m) ? (float)src[(size_t)f * src_ld + p] : 0.f;
}
int launch_convert_f64(const double* src, long long src_ld, float* dst, long long ld, long long m, int n, cudaStream_t s) {
    dim3 g(ew_grid(ld, EW_THREADS * 4, 1024), n < 512 ? n : 512);
    convert_f64_kernel<<<g, EW_THREADS, 0, s>>>(src, src_ld, dst, ld, m, n);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
__global__ void export_f64_kernel(const float* __restrict__ src, long long ld, double* __restrict__ dst, long long dst_ld,
                                  long long m, int n) {
    for (int f = blockIdx.y; f < n; f += gridDim.y)
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (long long)gridDim.x * blockDim.x)
            dst[(size_t)f * dst_ld + p] = (double)src[(size_t)f * ld + p];
}
int launch_export_f64(const float* src, long long ld, double* dst, long long dst_ld, long long m, int n, cudaStream_t s) {
    dim3 g(ew_grid(m, EW_THREADS * 4, 1024), n < 512 ? n : 512);
    export_f64_kernel<<<g, EW_THREADS, 0, s>>>(src, ld, dst, dst_ld, m, n);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
__global__ void copy_f32_kernel(const float* __restrict__ src, long long src_ld, float* __restrict__ dst, long long ld,
                                long long m, int n) {
    for (int f = blockIdx.y; f < n; f += gridDim.y)
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < ld; p += (long long)gridDim.x * blockDim.x)
            dst[(size_t)f * ld + p] = (p < m) ? src[(size_t)f * src_ld + p] : 0.f;
}
int launch_copy_f32(const float* src, long long src_ld, float* dst, long long ld, long long m, int n, cudaStream_t s) {
    dim3 g(ew_grid(ld, EW_THREADS * 4, 1024), n < 512 ? n : 512);
    copy_f32_kernel<<<g, EW_THREADS, 0, s>>>(src, src_ld, dst, ld, m, n);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- uint8 frames -> D: the LSD() pre-processing on the device (/root/reference/inexact_alm_lsd.py:211-225,
// normalizeImage /root/reference/utils.py:220-223): (x - min) * 1/(max - min) - mean, evaluated in fp64, stored fp32.
// sum_minmax: [0] = sum (u64), [1] = min, [2] = max (as u64)
__global__ void u8_stats_kernel(const unsigned char* __restrict__ src, long long count, unsigned long long* out) {
    __shared__ double red[32];
    unsigned long long sum = 0; unsigned int mn = 255, mx = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        unsigned int v = src[i];
        sum += v; mn = min(mn, v); mx = max(mx, v);
    }
    double s = block_sum((double)sum, red);          // exact: < 2^53
    if (threadIdx.x == 0) atomicAdd(out, (unsigned long long)(s + 0.5));
    double a = block_max(-(double)mn, red);
    if (threadIdx.x == 0) atomicMin(out + 1, (unsigned long long)(-a + 0.5));
    double b = block_max((double)mx, red);
    if (threadIdx.x == 0) atomicMax(out + 2, (unsigned long long)(b + 0.5));
}
int launch_u8_stats(const unsigned char* src, long long count, unsigned long long* sum_minmax, cudaStream_t s) {
    u8_stats_kernel<<<ew_grid(count, EW_THREADS * 16), EW_THREADS, 0, s>>>(src, count, sum_minmax);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}
__global__ void u8_to_D_kernel(const unsigned char* __restrict__ src, float* __restrict__ D, long long ld, long long m, int n,
                               double lo, double scale, double mean) {
    for (int f = blockIdx.y; f < n; f += gridDim.y)
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < ld; p += (long long)gridDim.x * blockDim.x)
            D[(size_t)f * ld + p] = (p < m) ? (float)(((double)src[(size_t)f * m + p] - lo) * scale - mean) : 0.f;
}
int launch_u8_to_D(const unsigned char* src, float* D, long long ld, long long m, int n, double lo, double scale, double mean,
                   cudaStream_t s) {
    dim3 g(ew_grid(ld, EW_THREADS * 4, 1024), n < 512 ? n : 512);
    u8_to_D_kernel<<<g, EW_THREADS, 0, s>>>(src, D, ld, m, n, lo, scale, mean);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ---- L = VC T   (svd_reconstruct, /root/reference/utils.py:185-186, from the kept projection T = Vr^T W)
// One thread owns a pixel quad and walks the frames; T rows for the quad stay in registers 8 at a time.
template <bool DUAL>
__global__ void __launch_bounds__(EW_THREADS) lowrank_kernel(const float* __restrict__ T, const float* __restrict__ VC, int vstride,
                                                             const DevState* st, float* __restrict__ L, long long ld, int n,
                                                             const float* __restrict__ D, const float* __restrict__ Snew,
                                                             float* __restrict__ S, float* __restrict__ Y,
                                                             double* part_zz, unsigned long long* part_nnz, float* part_max) {
    __shared__ double red[32];
    if (DUAL && st->done) return;
    const int r = DUAL ? st->svp : st->svp_L;
    const float mu_f = (float)st->mu;
    const long long nq = ld / 4;
    // frames are split over blockIdx.y so that small problems still fill the GPU
    const int fy0 = (int)(((long long)n * blockIdx.y) / gridDim.y), fy1 = (int)(((long long)n * (blockIdx.y + 1)) / gridDim.y);
    double zz = 0.0; unsigned long long nnz = 0; float mx = 0.f;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
        for (int k0 = 0; k0 < (r > 0 ? r : 1); k0 += 8) {
            float4 t4[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                t4[k] = (k0 + k < r) ? __ldg(reinterpret_cast<const float4*>(T + (size_t)(k0 + k) * ld + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const bool first = (k0 == 0), last = (k0 + 8 >= r);
            for (int f = fy0; f < fy1; ++f) {
                float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* vc = VC + (size_t)f * vstride + k0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float v = (k0 + k < r) ? __ldg(vc + k) : 0.f;
                    l.x = fmaf(v, t4[k].x, l.x); l.y = fmaf(v, t4[k].y, l.y); l.z = fmaf(v, t4[k].z, l.z); l.w = fmaf(v, t4[k].w, l.w);
                }
                const size_t off = (size_t)f * ld + 4 * q;
                if (!first) { const float4 o = *reinterpret_cast<const float4*>(L + off); l.x += o.x; l.y += o.y; l.z += o.z; l.w += o.w; }
                if (!DUAL || !last) { stg4(L + off, l); }
                if (DUAL && last) {
                    const float4 d = ldg4_stream(D + off), sn = ldg4_stream(Snew + off);
                    float4 y = *reinterpret_cast<const float4*>(Y + off);
                    const float z0 = (d.x - l.x) - sn.x, z1 = (d.y - l.y) - sn.y, z2 = (d.z - l.z) - sn.z, z3 = (d.w - l.w) - sn.w;
                    y.x = fmaf(mu_f, z0, y.x); y.y = fmaf(mu_f, z1, y.y); y.z = fmaf(mu_f, z2, y.z); y.w = fmaf(mu_f, z3, y.w);
                    stg4(Y + off, y); stg4(S + off, sn);
                    zz += (double)(z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3);
                    nnz += (sn.x != 0.f) + (sn.y != 0.f) + (sn.z != 0.f) + (sn.w != 0.f);
                    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(sn.x), fabsf(sn.y)), fmaxf(fabsf(sn.z), fabsf(sn.w))));
                }
            }
        }
    }
    if (DUAL) {
        const int b = blockIdx.y * gridDim.x + blockIdx.x;
        double zt = block_sum(zz, red);
        if (threadIdx.x == 0) part_zz[b] = zt;
        double nt = block_sum((double)nnz, red);
        if (threadIdx.x == 0) part_nnz[b] = (unsigned long long)(nt + 0.5);
        double mt = block_max((double)mx, red);
        if (threadIdx.x == 0) part_max[b] = (float)mt;
    }
}

static dim3 lowrank_grid(long long ld, int n, int max_blocks) {
    long long gx = (ld / 4 + EW_THREADS - 1) / EW_THREADS;
    if (gx < 1) gx = 1;
    int gy = 1;
    if (gx > max_blocks) gx = max_blocks;
    while (gx * gy < max_blocks / 2 && gy * 2 <= n) gy *= 2;
    if (gx * gy > max_blocks) gy = (int)(max_blocks / gx > 0 ? max_blocks / gx : 1);
    return dim3((unsigned)gx, (unsigned)gy);
}

int launch_materialize_L(const float* T, const float* VC, int vstride, const DevState* st, float* L, long long ld, long long m,
                         int n, cudaStream_t s) {
    dim3 g = lowrank_grid(ld, n, 148 * 8);
    lowrank_kernel<false><<<g, EW_THREADS, 0, s>>>(T, VC, vstride, st, L, ld, n, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                  nullptr, nullptr);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Second phase of the two-phase shrink (generic groups / overlapping graph / l2 blocks): S_new is given by a
// separate prox kernel; L is recomputed from T (partial sums for svp > 8 go through Lscratch), then
// Z = D - L - S_new, Y += mu Z, sum Z^2.
int launch_dual_update(const float* D, const float* Snew, float* S, float* Y, const float* T, const float* VC, int vstride,
                          const DevState* st, float* Lscratch, long long ld, int n, double* part_zz,
                          unsigned long long* part_nnz, float* part_max, int nparts, cudaStream_t s) {
    dim3 g = lowrank_grid(ld, n, nparts);
    // unused tail partials must not contribute
    BSUB_CUDA_CHECK(cudaMemsetAsync(part_zz, 0, sizeof(double) * nparts, s));
    BSUB_CUDA_CHECK(cudaMemsetAsync(part_nnz, 0, sizeof(unsigned long long) * nparts, s));
    BSUB_CUDA_CHECK(cudaMemsetAsync(part_max, 0, sizeof(float) * nparts, s));
    lowrank_kernel<true><<<g, EW_THREADS, 0, s>>>(T, VC, vstride, st, Lscratch, ld, n, D, Snew, S, Y, part_zz, part_nnz, part_max);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace bsub
