// bf16 1x1-convolution GEMM on the 5th-generation tensor cores (sm_100a).
//
//   C[M,N] = act( A[M,K] . W[N,K]^T + bias[N] ) (+ residual[M,N])        A, W, C, residual bf16; accumulate fp32
//
// A is the NHWC activation matrix (M = batch*H*W rows, K = input channels, K-major) and W the
// BN-folded weight matrix [N][K] (K-major) of an expand / project / head convolution of the
// reference's EfficientNet-B0 (model.py:63-72; SURVEY.md Appendix A "GEMM view").
//
// Design (one persistent CTA per SM, warp-specialised, no cluster):
//   warp 0  TMA producer: cp.async.bulk.tensor 2D loads of a 128 x 64 A tile and an N x 64 W tile
//           (SWIZZLE_128B; out-of-bounds rows/columns are zero-filled by TMA, so ragged M, K=16..1152
//           and N=16..256 need no padding copies) into a multi-stage smem ring, mbarrier complete_tx.
//   warp 1  MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N, K=16)
//           per 16-wide K step from smem descriptors; tcgen05.commit releases the smem stage and, after
//           the last K block, publishes the TMEM accumulator.  Two accumulators (2 x N columns of TMEM)
//           let the epilogue of tile i overlap the MMAs of tile i+1.
//   warp 2  allocates / frees TMEM.
//   warps 4-7  epilogue: tcgen05.ld 32 lanes x 16 columns -> registers, + bias, swish, + residual,
//           pack to bf16, 16-byte stores to C (row-contiguous NHWC).
// These layers are HBM-bound (arithmetic intensity 10-250 flop/B): the design goal is
// to stream A exactly once at full bandwidth; the tensor pipe has >10x headroom.
#include "dfd_internal.cuh"
#include <cuda.h>

#define BLOCK_M 128
#define BLOCK_K 64
#define GEMM_THREADS 256
#define A_STAGE_BYTES (BLOCK_M * BLOCK_K * 2)

// ---------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO unused for swizzled K-major (1), SBO = 8 rows * 128 B = 1024 B, version 1 (sm_100), layout SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ float swish_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

struct GemmParams {
    int M, N, K;
    int n_pad;            // UMMA N (multiple of 16, <= 256)
    int n_blocks;         // ceil(N / n_pad)
    int num_tiles;        // m_blocks * n_blocks
    int stages;
    int act;
    const float* bias;
    const __nv_bfloat16* residual;
    __nv_bfloat16* C;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * 8 + 4];     // full[8], empty[8], tmem_full[2], tmem_empty[2]
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b_stage_bytes = (uint32_t)p.n_pad * BLOCK_K * 2;
    const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
    const uint32_t tfull0 = smem_u32(&bars[16]), tempty0 = smem_u32(&bars[18]);
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * (uint32_t)p.n_pad) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int m_blk = tile / p.n_blocks, n_blk = tile % p.n_blocks;
                for (int kb = 0; kb < num_kb; kb++) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + A_STAGE_BYTES;
                    mbar_expect_tx(full0 + 8 * stage, stage_bytes);
                    tma_load_2d(sa, &map_a, kb * BLOCK_K, m_blk * BLOCK_M, full0 + 8 * stage);
                    tma_load_2d(sb, &map_b, kb * BLOCK_K, n_blk * p.n_pad, full0 + 8 * stage);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_pad);
                for (int kb = 0; kb < num_kb; kb++) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + A_STAGE_BYTES;
                    const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
                    const int krem = p.K - kb * BLOCK_K;
                    const int ksteps = krem >= BLOCK_K ? BLOCK_K / 16 : (krem + 15) / 16;
                    for (int k = 0; k < ksteps; k++)
                        tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit(empty0 + 8 * stage);                 // frees the smem stage when the MMAs retire
                    if (kb == num_kb - 1) tc_commit(tfull0 + 8 * acc);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;                                   // TMEM lane quarter = warp % 4
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.n_blocks, n_blk = tile % p.n_blocks;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tc_fence_after();
            const int m = m_blk * BLOCK_M + ew * 32 + lane;
            const int n_base = n_blk * p.n_pad;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.n_pad);
            const bool row_ok = m < p.M;
            __nv_bfloat16* crow = p.C + (size_t)m * p.N;
            const __nv_bfloat16* rrow = p.residual ? p.residual + (size_t)m * p.N : nullptr;
            for (int c = 0; c < p.n_pad; c += 16) {
                uint32_t r[16];
                tc_ld16(taddr + c, r);
                tc_ld_wait();
                const int n0 = n_base + c;
                if (row_ok && n0 < p.N) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int n = n0 + h * 8;
                        if (n + 8 <= p.N) {
                            float v[8];
                            const float4 b0 = __ldg((const float4*)(p.bias + n)), b1 = __ldg((const float4*)(p.bias + n + 4));
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                v[j] = __uint_as_float(r[h * 8 + j]) + bb[j];
                                if (p.act) v[j] = swish_fast(v[j]);
                            }
                            if (rrow) {
                                const uint4 rv = *(const uint4*)(rrow + n);
                                const __nv_bfloat162* rh = (const __nv_bfloat162*)&rv;
#pragma unroll
                                for (int j = 0; j < 4; j++) { float2 f = __bfloat1622float2(rh[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
                            }
                            uint4 o;
                            __nv_bfloat162* oh = (__nv_bfloat162*)&o;
#pragma unroll
                            for (int j = 0; j < 4; j++) oh[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                            *(uint4*)(crow + n) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int get_encode(dfd_ctx* ctx) {
    if (g_encode) return DFD_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DFD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    DFD_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, DFD_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode = (PFN_encodeTiled)fn;
    return DFD_OK;
}

// 2D bf16 row-major [rows][cols] tensor, box [box_rows][64], 128-byte swizzle
static int make_map(dfd_ctx* ctx, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return DFD_ERR_CUDA; }
    return DFD_OK;
}

static bool g_enabled = true;
bool dfd_gemm_bf16_enabled() { return g_enabled; }
void dfd_gemm_free(dfd_ctx*) {}

int dfd_gemm_bf16(dfd_ctx* ctx, const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias,
                  const __nv_bfloat16* residual, __nv_bfloat16* C, int M, int N, int K, int act, cudaStream_t st) {
    int rc = get_encode(ctx);
    if (rc) return rc;
    DFD_REQUIRE(K % 8 == 0 && N % 8 == 0, DFD_ERR_INVALID, "gemm: K and N must be multiples of 8");
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.act = act; p.bias = bias; p.residual = residual; p.C = C;
    // N tiling: the smallest number of equal UMMA-N blocks (multiples of 16, <= 256) covering N
    int nb = (N + 255) / 256;
    int n_pad = ((N + nb - 1) / nb + 15) / 16 * 16;
    p.n_pad = n_pad; p.n_blocks = (N + n_pad - 1) / n_pad;
    const int m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
    p.num_tiles = m_blocks * p.n_blocks;
    const int stage_bytes = A_STAGE_BYTES + n_pad * BLOCK_K * 2;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    if (stages > 2 * num_kb && stages > 4) stages = 2 * num_kb > 4 ? 2 * num_kb : 4;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        DFD_CUDA(cudaFuncSetAttribute(k_gemm_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        attr_set = true;
    }
    CUtensorMap ma, mb;
    if ((rc = make_map(ctx, &ma, A, (uint64_t)M, (uint64_t)K, BLOCK_M))) return rc;
    if ((rc = make_map(ctx, &mb, W, (uint64_t)N, (uint64_t)K, (uint32_t)n_pad))) return rc;
    int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
    k_gemm_tcgen05<<<grid, GEMM_THREADS, smem, st>>>(ma, mb, p);
    DFD_LAUNCH_CHECK("k_gemm_tcgen05", st);
    return DFD_OK;
}

// ---------------------------------------------------------------------------------------------
// self-test against a CUDA-core reference (used by tests/test_gpu_gemm.py through dfd_gemm_selftest)
__global__ void k_gemm_ref(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, const __nv_bfloat16* res,
                           float* C, int M, int N, int K, int act) {
    int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; k++) acc = fmaf(__bfloat162float(A[(size_t)m * K + k]), __bfloat162float(W[(size_t)n * K + k]), acc);
    acc += bias[n];
    if (act) acc = acc / (1.0f + expf(-acc));
    if (res) acc += __bfloat162float(res[(size_t)m * N + n]);
    C[(size_t)m * N + n] = acc;
}

__global__ void k_fill_bf16(__nv_bfloat16* x, size_t n, uint32_t seed, float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    x[i] = __float2bfloat16_rn(((float)(h & 0xffff) / 32768.0f - 1.0f) * scale);
}

__global__ void k_fill_f32(float* x, size_t n, uint32_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = (uint32_t)i * 2246822519u ^ seed;
    h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13;
    x[i] = (float)(h & 0xffff) / 65536.0f - 0.5f;
}

__global__ void k_maxerr(const __nv_bfloat16* c, const float* ref, size_t n, float* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float e = fabsf(__bfloat162float(c[i]) - ref[i]) / fmaxf(1.0f, fabsf(ref[i]));
    if (!(e == e)) e = 1e30f;
    atomicMax((int*)out, __float_as_int(e));        // non-negative floats order like ints
}

extern "C" int dfd_gemm_selftest(dfd_ctx* ctx, int M, int N, int K, int act, int with_residual, double* max_err_host,
                                 void* stream) {
    if (!ctx) return DFD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16 *A, *W, *R, *C;
    float *bias, *ref, *err;
    DFD_CUDA(cudaMalloc(&A, (size_t)M * K * 2));
    DFD_CUDA(cudaMalloc(&W, (size_t)N * K * 2));
    DFD_CUDA(cudaMalloc(&R, (size_t)M * N * 2));
    DFD_CUDA(cudaMalloc(&C, (size_t)M * N * 2));
    DFD_CUDA(cudaMalloc(&bias, (size_t)N * 4));
    DFD_CUDA(cudaMalloc(&ref, (size_t)M * N * 4));
    DFD_CUDA(cudaMalloc(&err, 4));
    DFD_CUDA(cudaMemsetAsync(err, 0, 4, st));
    DFD_CUDA(cudaMemsetAsync(C, 0xff, (size_t)M * N * 2, st));
    k_fill_bf16<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, st>>>(A, (size_t)M * K, 11u, 1.0f);
    k_fill_bf16<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(W, (size_t)N * K, 22u, 0.25f);
    k_fill_bf16<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(R, (size_t)M * N, 33u, 1.0f);
    k_fill_f32<<<(N + 255) / 256, 256, 0, st>>>(bias, (size_t)N, 44u);
    int rc = dfd_gemm_bf16(ctx, A, W, bias, with_residual ? R : nullptr, C, M, N, K, act, st);
    if (rc == DFD_OK) {
        k_gemm_ref<<<dim3((N + 127) / 128, M), 128, 0, st>>>(A, W, bias, with_residual ? R : nullptr, ref, M, N, K, act);
        k_maxerr<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(C, ref, (size_t)M * N, err);
        float e = 0.f;
        cudaError_t ce = cudaMemcpyAsync(&e, err, 4, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { ctx->err = std::string("gemm selftest: ") + cudaGetErrorString(ce); rc = DFD_ERR_CUDA; }
        if (max_err_host) *max_err_host = (double)e;
    }
    cudaFree(A); cudaFree(W); cudaFree(R); cudaFree(C); cudaFree(bias); cudaFree(ref); cudaFree(err);
    return rc;
}
