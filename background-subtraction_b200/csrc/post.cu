// post.cu -- the stages on either side of the decomposition (SURVEY.md section 8f rows 3 and 4, and the morphology of row 1),
// as device kernels with a C ABI (include/bsub_b200.h):
//
//   * bsub_resize_dev            cv2.resize INTER_AREA / INTER_CUBIC of every frame      /root/reference/utils.py:129-136
//   * bsub_cc_label_dev/_stats   8-connected components per frame + area / bbox / sums   utils.py:404-420, motion_saliency_check.py:19-49
//   * bsub_filter_sparse_map_dev connected-component size filter of a binary video        utils.py:404-420
//   * bsub_scube_product_dev, bsub_conv1d_reflect_dev   |xt| * |yt| / sum and the separable 3-D Gaussian of computeSCube.py:40-50,82-92
//   * bsub_morph_disk_dev        binary dilation / erosion by a disk (apply_morph_ops)    lsd_improvement.py:323-335
//
// All of them are byte / index work on images of a few MB per frame: HBM-bound streaming kernels with coalesced accesses along
// the contiguous pixel index.  Image-stack convention of the solver: frame f, image column j, image row i
// lives at f*ld + j*rows + i.  The resize and cube kernels take explicit element strides because their callers hold the
// reference's own [h][w][t] / [t][h][w] arrays.
#include <math.h>
#include <limits.h>
#include <vector>
#include "common.cuh"
#include "../../include/bsub_b200.h"

namespace bsub {

static cudaStream_t post_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static unsigned grid_for(long long work, int block, int cap = 148 * 16) {
    long long g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (unsigned)g;
}

// ------------------------------------------------------------------------------------------------------------------
// resize.  cv2.resize(src, dsize, interpolation): scale = src/dst per axis.
// INTER_AREA follows OpenCV's computeResizeAreaTab: destination cell d covers the source interval [d*scale, (d+1)*scale);
// source cells that are cut take the covered fraction, everything is divided by the cell width min(scale, ssize - d*scale).
// The tables (taps per destination index, CSR) are built by the launcher on the host in double and used for all frames.
// INTER_CUBIC: Keys kernel A = -0.75 at fx = (d + 0.5)*scale - 0.5, four taps, border replicate.
// ------------------------------------------------------------------------------------------------------------------
struct ResizeTab { std::vector<int> ptr, idx; std::vector<float> w; };

static void area_tab(int ssize, int dsize, ResizeTab& t) {
    const double scale = 1.0 / ((double)dsize / ssize);      // cv::resize: inv_scale = dsize/ssize, scale = 1/inv_scale
    t.ptr.assign(1, 0);
    for (int d = 0; d < dsize; ++d) {
        const double fsx1 = d * scale, fsx2 = fsx1 + scale;
        const double cell = std::min(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        if (sx1 - fsx1 > 1e-3) { t.idx.push_back(sx1 - 1); t.w.push_back((float)((sx1 - fsx1) / cell)); }
        for (int sx = sx1; sx < sx2; ++sx) { t.idx.push_back(sx); t.w.push_back((float)(1.0 / cell)); }
        if (fsx2 - sx2 > 1e-3) { t.idx.push_back(sx2); t.w.push_back((float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell)); }
        t.ptr.push_back((int)t.idx.size());
    }
}

static void cubic_tab(int ssize, int dsize, ResizeTab& t) {
    const double scale = 1.0 / ((double)dsize / ssize), A = -0.75;
    t.ptr.assign(1, 0);
    for (int d = 0; d < dsize; ++d) {
        double fx = (d + 0.5) * scale - 0.5;
        const int sx = (int)floor(fx);
        fx -= sx;
        float fxf = (float)fx;                              // OpenCV interpolates with a float fraction
        double c[4];
        const double x0 = fxf + 1.0, x1 = fxf, x2 = 1.0 - fxf;
        c[0] = ((A * x0 - 5 * A) * x0 + 8 * A) * x0 - 4 * A;
        c[1] = ((A + 2) * x1 - (A + 3)) * x1 * x1 + 1;
        c[2] = ((A + 2) * x2 - (A + 3)) * x2 * x2 + 1;
        c[3] = 1.0 - c[0] - c[1] - c[2];
        for (int k = 0; k < 4; ++k) {
            int s = sx - 1 + k;
            s = s < 0 ? 0 : (s >= ssize ? ssize - 1 : s);   // BORDER_REPLICATE
            t.idx.push_back(s); t.w.push_back((float)c[k]);
        }
        t.ptr.push_back((int)t.idx.size());
    }
}

// one thread per destination pixel; x (destination column index of the contiguous axis) is the fastest thread index
__global__ void __launch_bounds__(256) resize_kernel(const float* __restrict__ src, long long s_f, long long s_y, long long s_x, float* __restrict__ dst,
                                                     long long d_f, long long d_y, long long d_x, int n, int dh, int dw, const int* __restrict__ yptr,
                                                     const int* __restrict__ yidx, const float* __restrict__ yw, const int* __restrict__ xptr,
                                                     const int* __restrict__ xidx, const float* __restrict__ xw, int x_fast) {
    const long long total = (long long)n * dh * dw;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int f, y, x;
        if (x_fast) { x = (int)(t % dw); const long long r = t / dw; y = (int)(r % dh); f = (int)(r / dh); }
        else        { y = (int)(t % dh); const long long r = t / dh; x = (int)(r % dw); f = (int)(r / dw); }
        const float* sf = src + (long long)f * s_f;
        float acc = 0.f;
        for (int a = yptr[y]; a < yptr[y + 1]; ++a) {
            const float* row = sf + (long long)yidx[a] * s_y;
            float racc = 0.f;
            for (int b = xptr[x]; b < xptr[x + 1]; ++b) racc = fmaf(__ldg(row + (long long)xidx[b] * s_x), xw[b], racc);
            acc = fmaf(racc, yw[a], acc);
        }
        dst[(long long)f * d_f + (long long)y * d_y + (long long)x * d_x] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// connected components, 8-connectivity, one image per frame: union-find with atomicMin links towards the smaller pixel index
// (the root of a component is its smallest pixel index p = j*rows + i), then full path compression.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cc_find(const int* L, int x) {
    const volatile int* V = L;                        // other threads may be re-linking roots (cc_merge_kernel): no cached reads
    int y = V[x];
    while (y != x) { x = y; y = V[x]; }
    return x;
}
__device__ __forceinline__ void cc_union(int* L, int a, int b) {
    for (;;) {
        a = cc_find(L, a); b = cc_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[b], a);          // b was a root when read; if it still is, it now points to a
        if (old == b) return;
        b = old;
    }
}

__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t* __restrict__ mask, long long ld_mask, int* __restrict__ L, long long ld, int m, int n) {
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        L[(long long)f * ld + p] = mask[(long long)f * ld_mask + p] ? p : -1;
    }
}

__global__ void __launch_bounds__(256) cc_merge_kernel(int* __restrict__ L, long long ld, int rows, int cols, int n) {
    const int m = rows * cols;
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        int* Lf = L + (long long)f * ld;
        if (Lf[p] < 0) continue;
        const int j = p / rows, i = p - j * rows;
        // the four neighbours with a smaller pixel index: (i-1, j), (i-1, j-1), (i, j-1), (i+1, j-1)
        if (i > 0 && Lf[p - 1] >= 0) cc_union(Lf, p, p - 1);
        if (j > 0) {
            const int q = p - rows;
            if (Lf[q] >= 0) cc_union(Lf, p, q);
            if (i > 0 && Lf[q - 1] >= 0) cc_union(Lf, p, q - 1);
            if (i + 1 < rows && Lf[q + 1] >= 0) cc_union(Lf, p, q + 1);
        }
    }
}

// root of every foreground pixel; optionally the area of every component, accumulated at its root
__global__ void __launch_bounds__(256) cc_compress_kernel(int* __restrict__ L, long long ld, int m, int n, int* __restrict__ area) {
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        int* Lf = L + (long long)f * ld;
        if (Lf[p] < 0) continue;
        const int r = cc_find(Lf, p);
        if (area != nullptr) atomicAdd(&area[(long long)f * ld + r], 1);
        // no write here: other threads still walk the links (roots never change any more, so a separate flatten pass follows)
    }
}
__global__ void __launch_bounds__(256) cc_flatten_kernel(const int* __restrict__ L, int* __restrict__ R, long long ld, int m, int n) {
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        const int* Lf = L + (long long)f * ld;
        R[(long long)f * ld + p] = (Lf[p] < 0) ? -1 : cc_find(Lf, p);
    }
}

// consecutive ids 1..K per frame in the order of the roots' pixel index; one CTA per frame (block-wide scan over the pixels)
__global__ void __launch_bounds__(1024) cc_number_kernel(const int* __restrict__ R, long long ld, int m, int* __restrict__ ids /*[n][ld] scratch*/,
                                                         int* __restrict__ num_labels) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* Rf = R + (long long)f * ld;
    int* If = ids + (long long)f * ld;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += blockDim.x) {
        const int p = base + threadIdx.x;
        const int flag = (p < m && Rf[p] == p) ? 1 : 0;
        int v = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
            wsum[lane] = w;
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? wsum[warp - 1] : 0) + v - flag;
        if (flag) If[p] = before + 1;
        __syncthreads();
        if (threadIdx.x == 0) carry += wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) num_labels[f] = carry;
}
__global__ void __launch_bounds__(256) cc_relabel_kernel(const int* __restrict__ R, const int* __restrict__ ids, int* __restrict__ labels, long long ld,
                                                         int m, int n) {
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        const int r = R[(long long)f * ld + p];
        labels[(long long)f * ld + p] = (r < 0) ? 0 : ids[(long long)f * ld + r];
    }
}

// per-component statistics (cv2.connectedComponentsWithStats columns + the weight sum of compute_groups_per_frame)
__global__ void __launch_bounds__(256) cc_stats_kernel(const int* __restrict__ labels, long long ld, int rows, int cols, int n,
                                                       const int* __restrict__ offsets, const float* __restrict__ weight, long long w_f, long long w_j,
                                                       long long w_i, int* __restrict__ area,
                                                       int* __restrict__ box /*[total][5]: min i, max i, min j, max j, first 2x2 block*/,
                                                       double* __restrict__ wsum) {
    const int m = rows * cols;
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        const int l = labels[(long long)f * ld + p];
        if (l <= 0) continue;
        const int k = offsets[f] + l - 1;
        const int j = p / rows, i = p - j * rows;
        atomicAdd(&area[k], 1);
        atomicMin(&box[5 * k + 0], i); atomicMax(&box[5 * k + 1], i);
        atomicMin(&box[5 * k + 2], j); atomicMax(&box[5 * k + 3], j);
        // OpenCV numbers the components in the raster order of the 2x2 block that first meets them (block-based two-pass
        // labelling, labels flattened in creation order): the smallest block index is the sort key that reproduces its numbering
        atomicMin(&box[5 * k + 4], (i >> 1) * ((cols + 1) >> 1) + (j >> 1));
        if (weight != nullptr) atomicAdd(&wsum[k], (double)weight[(long long)f * w_f + (long long)j * w_j + (long long)i * w_i]);
    }
}
__global__ void __launch_bounds__(256) cc_box_init_kernel(int* __restrict__ area, int* __restrict__ box, double* __restrict__ wsum, int total) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        area[k] = 0; box[5 * k + 0] = INT_MAX; box[5 * k + 1] = -1; box[5 * k + 2] = INT_MAX; box[5 * k + 3] = -1; box[5 * k + 4] = INT_MAX;
        if (wsum != nullptr) wsum[k] = 0.0;
    }
}

__global__ void __launch_bounds__(256) cc_filter_kernel(const int* __restrict__ R, const int* __restrict__ area, long long ld, int m, int n, int size_thresh,
                                                        uint8_t* __restrict__ out, long long ld_out) {
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        const int r = R[(long long)f * ld + p];
        out[(long long)f * ld_out + p] = (r >= 0 && area[(long long)f * ld + r] > size_thresh) ? 1 : 0;
    }
}

// labels 1..K of a frame -> table[offsets[f] + label - 1] (uint8; e.g. the block ids of bsub_set_blocks), 0 stays 0
__global__ void __launch_bounds__(256) cc_remap_kernel(const int* __restrict__ labels, long long ld, int m, int n, const int* __restrict__ offsets,
                                                       const uint8_t* __restrict__ table, uint8_t* __restrict__ out, long long ld_out) {
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        const int l = labels[(long long)f * ld + p];
        out[(long long)f * ld_out + p] = (l > 0) ? table[offsets[f] + l - 1] : 0;
    }
}

// roots into R (scratch int32 [n][ld]); L is a second scratch of the same size
static int cc_roots(const uint8_t* mask, long long ld_mask, int rows, int cols, int n, long long ld, int* L, int* R, int* area, cudaStream_t st) {
    const int m = rows * cols;
    const unsigned g = grid_for((long long)m * n, 256);
    cc_init_kernel<<<g, 256, 0, st>>>(mask, ld_mask, L, ld, m, n);
    cc_merge_kernel<<<g, 256, 0, st>>>(L, ld, rows, cols, n);
    if (area != nullptr) {
        BSUB_CUDA_CHECK(cudaMemsetAsync(area, 0, sizeof(int) * (size_t)ld * n, st));
        cc_compress_kernel<<<g, 256, 0, st>>>(L, ld, m, n, area);
    }
    cc_flatten_kernel<<<g, 256, 0, st>>>(L, R, ld, m, n);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// computeSCube: cube[t][h][w] = |xt[w][h][t]| * |yt[h][w][t]|, its sum (fp64, fixed order: per-CTA partials, then one CTA),
// and a 1-D correlation along one axis with scipy.ndimage's 'reflect' boundary (d c b a | a b c d | d c b a).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scube_product_kernel(const float* __restrict__ xt, const float* __restrict__ yt, float* __restrict__ cube, int T,
                                                            int H, int W, double* __restrict__ partial) {
    __shared__ double red[32];
    const long long total = (long long)T * H * W;
    double acc = 0.0;
    // thread index runs over (h, w, t) with t fastest for the reads (both inputs are contiguous in t); the transposed
    // write goes through the L2 (the cube is a few tens of MB)
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(e % T);
        const long long r = e / T;
        const int w = (int)(r % W), h = (int)(r / W);
        const float v = fabsf(xt[((long long)w * H + h) * T + t]) * fabsf(yt[((long long)h * W + w) * T + t]);
        cube[((long long)t * H + h) * W + w] = v;
        acc += (double)v;
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ partial, int count, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partial[i];
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = s;
}

// dst[.., i, ..] = scale * sum_k w[k] * src[.., reflect(i + k - shift), ..] along the axis of length `len` and element stride
// `stride`; `inner` = product of the faster axes' extents (the element index is outer*len*inner + i*inner + in)
__global__ void __launch_bounds__(256) conv1d_reflect_kernel(const float* __restrict__ src, float* __restrict__ dst, long long total, int len,
                                                             long long inner, const float* __restrict__ w, int taps, int shift,
                                                             const double* __restrict__ inv_scale_src) {
    const float scale = (inv_scale_src != nullptr) ? (float)(1.0 / inv_scale_src[0]) : 1.f;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long in = e % inner, r = e / inner;
        const int i = (int)(r % len);
        const long long base = (r / len) * len * inner + in;
        float acc = 0.f;
        for (int k = 0; k < taps; ++k) {
            int s = i + k - shift;
            // 'reflect' with period 2*len (kernels longer than the axis wrap more than once)
            const int per = 2 * len;
            s %= per; if (s < 0) s += per;
            if (s >= len) s = per - 1 - s;
            acc = fmaf(w[k], src[base + (long long)s * inner], acc);
        }
        dst[e] = acc * scale;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// binary morphology with a disk footprint (skimage.morphology.disk(r): di^2 + dj^2 <= r^2), out-of-image pixels ignored
// (for a disk this equals scipy's 'reflect' boundary, the default of skimage's dilation / erosion).
//   pass 1: run[p] = distance along the image column (the contiguous index i) to the nearest pixel of the target value
//           (dilation: nearest set pixel; erosion: nearest clear pixel), capped at r + 1
//   pass 2: dilation: out = any dj: run[i, j + dj] <= hw(dj);   erosion: out = all dj: run[i, j + dj] > hw(dj)
// ------------------------------------------------------------------------------------------------------------------
constexpr int MORPH_RMAX = 255;
__global__ void __launch_bounds__(256) morph_run_kernel(const uint8_t* __restrict__ src, long long ld_src, uint8_t* __restrict__ run, long long ld, int rows,
                                                        int cols, int n, int r, int target) {
    const int m = rows * cols;
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        const int j = p / rows, i = p - j * rows;
        const uint8_t* col = src + (long long)f * ld_src + (long long)j * rows;
        int d = r + 1;
        for (int k = 0; k <= r; ++k) {
            const bool up = (i - k >= 0) && ((col[i - k] != 0) == (target != 0));
            const bool dn = (i + k < rows) && ((col[i + k] != 0) == (target != 0));
            if (up || dn) { d = k; break; }
        }
        run[(long long)f * ld + p] = (uint8_t)d;
    }
}
__global__ void __launch_bounds__(256) morph_apply_kernel(const uint8_t* __restrict__ run, long long ld, uint8_t* __restrict__ dst, long long ld_dst, int rows,
                                                          int cols, int n, int r, const uint8_t* __restrict__ hw /*[2r+1]*/, int erode) {
    const int m = rows * cols;
    const long long total = (long long)m * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / m), p = (int)(t - (long long)f * m);
        const int j = p / rows, i = p - j * rows;
        const uint8_t* rf = run + (long long)f * ld;
        bool hit = false;                     // a pixel of the target value inside the disk
        const int j0 = max(0, j - r), j1 = min(cols - 1, j + r);
        for (int jj = j0; jj <= j1 && !hit; ++jj) hit = rf[(long long)jj * rows + i] <= hw[jj - j + r];
        dst[(long long)f * ld_dst + p] = erode ? (hit ? 0 : 1) : (hit ? 1 : 0);
    }
}

}  // namespace bsub

using namespace bsub;

extern "C" {

int bsub_resize_dev(const float* src, int64_t src_stride_f, int64_t src_stride_y, int64_t src_stride_x, int32_t src_h, int32_t src_w, int32_t n,
                    float* dst, int64_t dst_stride_f, int64_t dst_stride_y, int64_t dst_stride_x, int32_t dst_h, int32_t dst_w, int32_t interp,
                    void* stream) {
    if (!src || !dst || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1 || n < 1) { set_error("bsub_resize_dev: bad argument"); return -1; }
    if (interp != 0 && interp != 1) { set_error("bsub_resize_dev: interp must be 0 (INTER_AREA) or 1 (INTER_CUBIC)"); return -1; }
    if (interp == 0 && (dst_h > src_h || dst_w > src_w)) { set_error("bsub_resize_dev: INTER_AREA is implemented for shrinking only (the reference switches to INTER_CUBIC for ratio >= 1)"); return -1; }
    cudaStream_t st = post_stream(stream);
    ResizeTab ty, tx;
    if (interp == 0) { area_tab(src_h, dst_h, ty); area_tab(src_w, dst_w, tx); }
    else             { cubic_tab(src_h, dst_h, ty); cubic_tab(src_w, dst_w, tx); }
    // one device buffer for the six tables
    const size_t ni = ty.ptr.size() + ty.idx.size() + tx.ptr.size() + tx.idx.size(), nf = ty.w.size() + tx.w.size();
    std::vector<int> hi; hi.reserve(ni);
    hi.insert(hi.end(), ty.ptr.begin(), ty.ptr.end()); hi.insert(hi.end(), ty.idx.begin(), ty.idx.end());
    hi.insert(hi.end(), tx.ptr.begin(), tx.ptr.end()); hi.insert(hi.end(), tx.idx.begin(), tx.idx.end());
    std::vector<float> hf; hf.reserve(nf);
    hf.insert(hf.end(), ty.w.begin(), ty.w.end()); hf.insert(hf.end(), tx.w.begin(), tx.w.end());
    int* di = nullptr; float* df = nullptr;
    BSUB_CUDA_CHECK(cudaMalloc(&di, sizeof(int) * ni));
    if (cudaMalloc(&df, sizeof(float) * nf) != cudaSuccess) { cudaFree(di); set_error("bsub_resize_dev: cudaMalloc failed"); return -1; }
    int rc = 0;
    if (cudaMemcpyAsync(di, hi.data(), sizeof(int) * ni, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(df, hf.data(), sizeof(float) * nf, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (rc == 0) {
        const int* yptr = di; const int* yidx = yptr + ty.ptr.size(); const int* xptr = yidx + ty.idx.size(); const int* xidx = xptr + tx.ptr.size();
        const float* yw = df; const float* xw = df + ty.w.size();
        const int x_fast = (dst_stride_x <= dst_stride_y) ? 1 : 0;
        resize_kernel<<<grid_for((long long)n * dst_h * dst_w, 256), 256, 0, st>>>(src, src_stride_f, src_stride_y, src_stride_x, dst, dst_stride_f,
                                                                                  dst_stride_y, dst_stride_x, n, dst_h, dst_w, yptr, yidx, yw, xptr, xidx,
                                                                                  xw, x_fast);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) rc = -1;    // the tables are freed below
    }
    cudaFree(di); cudaFree(df);
    if (rc != 0) set_error("bsub_resize_dev: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int bsub_cc_label_dev(const uint8_t* mask, int64_t ld_mask, int32_t rows, int32_t cols, int32_t n, int32_t* labels, int64_t ld, int32_t* num_labels,
                      int32_t* scratch /* 2 x [n][ld] int32 */, void* stream) {
    if (!mask || !labels || !num_labels || !scratch || rows < 1 || cols < 1 || n < 1 || (long long)rows * cols > ld || (long long)rows * cols > ld_mask ||
        (long long)rows * cols > INT_MAX) { set_error("bsub_cc_label_dev: bad argument"); return -1; }
    cudaStream_t st = post_stream(stream);
    const int m = rows * cols;
    int* L = scratch; int* R = scratch + (size_t)ld * n;
    if (cc_roots(mask, ld_mask, rows, cols, n, ld, L, R, nullptr, st) != 0) return -1;
    cc_number_kernel<<<n, 1024, 0, st>>>(R, ld, m, L /* ids at the roots */, num_labels);
    cc_relabel_kernel<<<grid_for((long long)m * n, 256), 256, 0, st>>>(R, L, labels, ld, m, n);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int bsub_cc_stats_dev(const int32_t* labels, int64_t ld, int32_t rows, int32_t cols, int32_t n, const int32_t* offsets, int32_t total,
                      const float* weight, int64_t w_stride_f, int64_t w_stride_j, int64_t w_stride_i, int32_t* area, int32_t* box, double* wsum,
                      void* stream) {
    if (!labels || !offsets || !area || !box || total < 0 || (weight != nullptr && wsum == nullptr)) { set_error("bsub_cc_stats_dev: bad argument"); return -1; }
    if (total == 0) return 0;
    cudaStream_t st = post_stream(stream);
    cc_box_init_kernel<<<grid_for(total, 256), 256, 0, st>>>(area, box, wsum, total);
    cc_stats_kernel<<<grid_for((long long)rows * cols * n, 256), 256, 0, st>>>(labels, ld, rows, cols, n, offsets, weight, w_stride_f, w_stride_j,
                                                                              w_stride_i, area, box, wsum);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int bsub_cc_remap_dev(const int32_t* labels, int64_t ld, int64_t m, int32_t n, const int32_t* offsets, const uint8_t* table, uint8_t* out, int64_t ld_out,
                      void* stream) {
    if (!labels || !offsets || !table || !out || m < 1 || n < 1 || m > ld || m > ld_out || m > INT_MAX) { set_error("bsub_cc_remap_dev: bad argument"); return -1; }
    cc_remap_kernel<<<grid_for((long long)m * n, 256), 256, 0, post_stream(stream)>>>(labels, ld, (int)m, n, offsets, table, out, ld_out);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int bsub_filter_sparse_map_dev(const uint8_t* mask, int64_t ld_mask, int32_t rows, int32_t cols, int32_t n, int32_t size_thresh, uint8_t* out,
                               int64_t ld_out, int32_t* scratch /* 3 x [n][ld] int32, ld = rows*cols */, void* stream) {
    if (!mask || !out || !scratch || rows < 1 || cols < 1 || n < 1 || (long long)rows * cols > ld_mask || (long long)rows * cols > ld_out ||
        (long long)rows * cols > INT_MAX) { set_error("bsub_filter_sparse_map_dev: bad argument"); return -1; }
    cudaStream_t st = post_stream(stream);
    const int m = rows * cols;
    const long long ld = m;
    int* L = scratch; int* R = L + (size_t)ld * n; int* area = R + (size_t)ld * n;
    if (cc_roots(mask, ld_mask, rows, cols, n, ld, L, R, area, st) != 0) return -1;
    cc_filter_kernel<<<grid_for((long long)m * n, 256), 256, 0, st>>>(R, area, ld, m, n, size_thresh, out, ld_out);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int bsub_scube_product_dev(const float* xt, const float* yt, float* cube, int32_t T, int32_t H, int32_t W, double* sum_out,
                           double* scratch /* >= 2048 doubles */, void* stream) {
    if (!xt || !yt || !cube || !sum_out || !scratch || T < 1 || H < 1 || W < 1) { set_error("bsub_scube_product_dev: bad argument"); return -1; }
    cudaStream_t st = post_stream(stream);
    const unsigned g = grid_for((long long)T * H * W, 256, 2048);
    scube_product_kernel<<<g, 256, 0, st>>>(xt, yt, cube, T, H, W, scratch);
    sum_partials_kernel<<<1, 256, 0, st>>>(scratch, (int)g, sum_out);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int bsub_conv1d_reflect_dev(const float* src, float* dst, int64_t outer, int32_t len, int64_t inner, const float* weights_dev, int32_t taps,
                            int32_t shift, const double* divide_by_dev, void* stream) {
    if (!src || !dst || src == dst || !weights_dev || outer < 1 || len < 1 || inner < 1 || taps < 1) { set_error("bsub_conv1d_reflect_dev: bad argument"); return -1; }
    const long long total = (long long)outer * len * inner;
    conv1d_reflect_kernel<<<grid_for(total, 256), 256, 0, post_stream(stream)>>>(src, dst, total, len, inner, weights_dev, taps, shift, divide_by_dev);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int bsub_morph_disk_dev(const uint8_t* src, int64_t ld_src, uint8_t* dst, int64_t ld_dst, int32_t rows, int32_t cols, int32_t n, int32_t radius,
                        int32_t erode, uint8_t* scratch /* [n][rows*cols] + 2*radius+1 bytes */, void* stream) {
    if (!src || !dst || !scratch || rows < 1 || cols < 1 || n < 1 || radius < 0 || radius >= MORPH_RMAX || (long long)rows * cols > ld_src ||
        (long long)rows * cols > ld_dst || (long long)rows * cols > INT_MAX) { set_error("bsub_morph_disk_dev: bad argument (radius < 255)"); return -1; }
    cudaStream_t st = post_stream(stream);
    const int m = rows * cols;
    uint8_t* run = scratch;
    uint8_t* hw_dev = scratch + (size_t)m * n;
    std::vector<uint8_t> hw((size_t)2 * radius + 1);
    for (int d = -radius; d <= radius; ++d) {
        int h = (int)floor(sqrt((double)radius * radius - (double)d * d));
        while ((long long)(h + 1) * (h + 1) + (long long)d * d <= (long long)radius * radius) ++h;     // exact integer test (disk(): X^2 + Y^2 <= r^2)
        while ((long long)h * h + (long long)d * d > (long long)radius * radius) --h;
        hw[(size_t)(d + radius)] = (uint8_t)h;
    }
    BSUB_CUDA_CHECK(cudaMemcpyAsync(hw_dev, hw.data(), hw.size(), cudaMemcpyHostToDevice, st));
    BSUB_CUDA_CHECK(cudaStreamSynchronize(st));                                // hw is a stack object
    const unsigned g = grid_for((long long)m * n, 256);
    morph_run_kernel<<<g, 256, 0, st>>>(src, ld_src, run, m, rows, cols, n, radius, erode ? 0 : 1);
    morph_apply_kernel<<<g, 256, 0, st>>>(run, m, dst, ld_dst, rows, cols, n, radius, hw_dev, erode ? 1 : 0);
    BSUB_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
