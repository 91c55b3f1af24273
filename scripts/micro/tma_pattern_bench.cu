// Micro-benchmark (round 2): how fast can one B200 move the shrink pass's data with TMA boxes of a given shape?
// Arrays: float [n][cols][rows] (rows contiguous).  Per stage a persistent CTA loads NIN boxes {R rows, NC cols, FC frames} and
// stores NOUT boxes from the same shared memory (no compute).  Prints GB/s for a list of shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_pattern_bench tma_pattern_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    unsigned long long spins = 0;
    while (!mbar_try(b, par)) { if (++spins > (1ull << 26)) { printf("watchdog: block %d thread %d bar %p par %u\n", blockIdx.x, threadIdx.x, b, par); __trap(); } }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Args { int rows, cols, n, R, NC, FC, NS, nin, nout, ntr, ntc, ncf, order; };

__global__ void __launch_bounds__(96, 1) bench_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
                                                      const __grid_constant__ CUtensorMap mC, const __grid_constant__ CUtensorMap mD,
                                                      const __grid_constant__ CUtensorMap mE, Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t slot = ((size_t)a.R * a.NC * a.FC * 4 + 127) / 128 * 128;
    const size_t stage = 3 * slot;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.NS * stage);
    uint64_t* freeb = full + a.NS;
    if (threadIdx.x == 0) { for (int s = 0; s < a.NS; ++s) { mbar_init(&full[s], 1); mbar_init(&freeb[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const long long ntiles = (long long)a.ntr * a.ntc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = (ntiles * blockIdx.x) / gridDim.x, t1 = (ntiles * (blockIdx.x + 1)) / gridDim.x;
    if (warp == 0 && lane == 0) {
        long long q = 0;
        if (a.order == 0) {
            for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
                const int tc = (int)(tl / a.ntr), tr = (int)(tl % a.ntr);
                for (int c = 0; c < a.ncf; ++c, ++q) {
                    const int s = (int)(q % a.NS);
                    if (q >= a.NS) mbar_wait(&freeb[s], (uint32_t)(((q / a.NS) - 1) & 1));
                    unsigned char* b = smem + (size_t)s * stage;
                    mbar_expect_tx(&full[s], (uint32_t)(a.nin * (size_t)a.R * a.NC * a.FC * 4));
                    tma_load_3d(b, &mA, &full[s], tr * a.R, tc * a.NC, c * a.FC);
                    if (a.nin > 1) tma_load_3d(b + slot, &mB, &full[s], tr * a.R, tc * a.NC, c * a.FC);
                }
            }
        } else {
            for (int c = 0; c < a.ncf; ++c)
                for (long long tl = t0; tl < t1; ++tl, ++q) {
                    const int tc = (int)(tl / a.ntr), tr = (int)(tl % a.ntr);
                    const int s = (int)(q % a.NS);
                    if (q >= a.NS) mbar_wait(&freeb[s], (uint32_t)(((q / a.NS) - 1) & 1));
                    unsigned char* b = smem + (size_t)s * stage;
                    mbar_expect_tx(&full[s], (uint32_t)(a.nin * (size_t)a.R * a.NC * a.FC * 4));
                    tma_load_3d(b, &mA, &full[s], tr * a.R, tc * a.NC, c * a.FC);
                    if (a.nin > 1) tma_load_3d(b + slot, &mB, &full[s], tr * a.R, tc * a.NC, c * a.FC);
                }
        }
    } else if (warp == 1 && lane == 0) {
        long long q = 0;
        auto one = [&](int tc, int tr, int c) {
            const int s = (int)(q % a.NS);
            mbar_wait(&full[s], (uint32_t)((q / a.NS) & 1));
            unsigned char* b = smem + (size_t)s * stage;
            if (a.nout > 0) tma_store_3d(&mC, b, tr * a.R, tc * a.NC, c * a.FC);
            if (a.nout > 1) tma_store_3d(&mD, b + slot, tr * a.R, tc * a.NC, c * a.FC);
            if (a.nout > 2) tma_store_3d(&mE, b + 2 * slot, tr * a.R, tc * a.NC, c * a.FC);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&freeb[s]);
            ++q;
        };
        if (a.order == 0) {
            for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x)
                for (int c = 0; c < a.ncf; ++c) one((int)(tl / a.ntr), (int)(tl % a.ntr), c);
        } else {
            for (int c = 0; c < a.ncf; ++c)
                for (long long tl = t0; tl < t1; ++tl) one((int)(tl / a.ntr), (int)(tl % a.ntr), c);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_enc get_enc() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    return (PFN_enc)fn;
}
static void make_map(CUtensorMap* m, void* base, int rows, int cols, int n, int R, int NC, int FC) {
    static PFN_enc enc = get_enc();
    cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)cols, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)rows * 4, (cuuint64_t)rows * cols * 4};
    cuuint32_t box[3] = {(cuuint32_t)R, (cuuint32_t)NC, (cuuint32_t)FC}, es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map encode failed %d (R=%d NC=%d FC=%d)\n", (int)r, R, NC, FC); exit(1); }
}

int main(int argc, char** argv) {
    const int rows = 1080, cols = 1920, n = 300;
    const size_t bytes = (size_t)rows * cols * n * 4;
    setvbuf(stdout, nullptr, _IONBF, 0);
    float* buf[5];
    for (int i = 0; i < 5; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], i, bytes)); }
    CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    struct Shape { int R, NC, FC, nin, nout, order; };
    std::vector<Shape> shapes = {
        {48, 3, 16, 2, 3, 0}, {48, 3, 16, 2, 3, 1}, {48, 3, 16, 2, 2, 0}, {48, 3, 16, 2, 2, 1}, {216, 3, 4, 2, 2, 0}, {216, 3, 4, 2, 2, 1},
        {48, 3, 16, 1, 0, 0}, {48, 3, 16, 1, 0, 1},
    };
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // reference: plain device-to-device copy
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(buf[2], buf[0], bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) { printf("cudaMemcpy D2D: %.3f ms  %.0f GB/s (read+write)\n", ms, 2.0 * bytes / ms / 1e6); fflush(stdout); }
    }
    for (const Shape& sh : shapes) {
        Args a; a.rows = rows; a.cols = cols; a.n = n; a.R = sh.R; a.NC = sh.NC; a.FC = sh.FC; a.nin = sh.nin; a.nout = sh.nout; a.order = sh.order;
        a.ntr = (rows + sh.R - 1) / sh.R; a.ntc = (cols + sh.NC - 1) / sh.NC; a.ncf = (n + sh.FC - 1) / sh.FC;
        const size_t slot = ((size_t)sh.R * sh.NC * sh.FC * 4 + 127) / 128 * 128;
        int NS = (int)((220 * 1024) / (3 * slot)); if (NS > 12) NS = 12; if (NS < 2) { printf("shape too big\n"); continue; }
        a.NS = NS;
        CUtensorMap m[5];
        for (int i = 0; i < 5; ++i) make_map(&m[i], buf[i], rows, cols, n, sh.R, sh.NC, sh.FC);
        const size_t smem = (size_t)NS * 3 * slot + 2 * NS * 8 + 64;
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            bench_kernel<<<148, 96, smem>>>(m[0], m[1], m[2], m[3], m[4], a);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double moved = (double)(sh.nin + sh.nout) * bytes;
        printf("order %d box {%3d rows, %2d cols, %2d frames} = %6.1f KB/slot, %2d stages, %d in / %d out: %.3f ms  %.0f GB/s  (row run %d B)\n", sh.order, sh.R, sh.NC, sh.FC,
               slot / 1024.0, NS, sh.nin, sh.nout, best, moved / best / 1e6, sh.R * 4);
        fflush(stdout);
    }
    return 0;
}
