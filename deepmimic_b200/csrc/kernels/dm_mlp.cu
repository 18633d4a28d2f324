// Policy network of the batched rollout: the actor of the reference's PPO / AMP agents,
//   a = un-normalise( W2^T relu( W1^T relu( W0^T normalise(s) + b0 ) + b1 ) + b2 )        (R/learning/nets/fc_2layers_1024units.py,
//   R/learning/pg_agent.py:140-160, R/learning/normalizer.py), 227 -> 1024 -> 512 -> 28 for humanoid3d,
// as one operand-preparation launch + three launches of one sm_100a GEMM kernel built on the 5th-generation tensor cores:
//   * tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 128 / 32 / 64, K = 16 per instruction) issued by one thread, accumulators in TMEM (fp32),
//   * BOTH operands live in global memory already tiled in the shared-memory operand layout (canonical K-major, no swizzle: 8 x 16-byte core
//     matrices): the weights are tiled once on the host, the activations are written in that layout by the producing launch (the preparation
//     kernel for the observations, the previous layer's epilogue otherwise).  A K-chunk of either operand is therefore ONE contiguous block
//     that the TMA unit brings in as a bulk copy (cp.async.bulk ... mbarrier::complete_tx),
//   * warp-specialised 4-stage pipeline without block-wide barriers in the main loop: thread 0 = TMA producer (waits on a stage's "empty"
//     mbarrier, arms "full" with the byte count, issues the two bulk copies), thread 32 = MMA issuer (waits on "full", issues 4 K-steps x (hi, lo)
//     MMAs, tcgen05.commit -> "empty"), all four warps = epilogue after the last commit,
//   * every weight is carried as fp16 hi + fp16 lo (w = hi + lo to 2^-22): two MMAs per K-step make the weights exact to fp32 level, so the only
//     rounding beyond the fp32 reference is the fp16 rounding of the activations (measured action error < 1e-3, tests/test_mlp_gpu.py),
//   * epilogue: tcgen05.ld of the accumulator rows (one TMEM lane = one environment), bias, ReLU, fp16 tiles for the next layer (a warp writes
//     512 contiguous bytes per 8 columns); the last layer adds the exploration noise and un-normalises into the DeepMimic action layout (fp32).
// This is the one dense contraction next to the hot path (the simulation itself has none); it replaces cuBLAS / eager torch in the rollout shim.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace dmk {

constexpr int kMlpBM = 128;        // environments per CTA (= TMEM lanes)
constexpr int kMlpBK = 64;         // K elements per chunk (8 core matrices of 8 fp16)
constexpr int kMlpThreads = 128;   // 4 warps: warp w owns TMEM lanes [32 w, 32 w + 32)
constexpr int kMlpStages = 4;
constexpr int kMlpATile = kMlpBM * kMlpBK;   // halves per activation tile (16 KB): [k8][row group][row in group][8 halves]

struct MlpPrepParams {
    const float* obs;          // [M x in_dim] fp32 observations
    const float* in_mean;      // normaliser mean / 1/std, in_dim entries
    const float* in_istd;
    float in_clip;
    int in_dim, M, NC;         // NC = padded K / 64
    __half* tiles;             // [m tiles][NC][kMlpATile]
};
struct MlpGemmParams {
    const __half* a_tiles;     // [m tiles][K / 64][kMlpATile] fp16 activations in operand layout
    const __half* w_tiles;     // [n tiles][K / 64][hi | lo][BN x 64] in operand layout
    const float* bias;         // [N padded]
    __half* out_tiles;         // !LAST: [m tiles][N / 64][kMlpATile]
    float* actions;            // LAST: [M x out_dim] fp32
    const float* out_mean;     // LAST: action un-normalisation a * std + mean
    const float* out_std;
    const float* noise;        // LAST, optional: [M x out_dim] added in normalised action space (exploration), may be null
    int out_dim;
    int M, K, N;               // rows, padded K (multiple of 64), padded N (multiple of BN and, for !LAST, of 64)
};

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a protocol error traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor, canonical K-major layout without swizzle (cute::UMMA::SmemDescriptor, version 1):
//   core matrix = 8 rows x 16 bytes, rows 16 bytes apart; SBO = bytes between 8-row groups, LBO = bytes between the two 16-byte K slices
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) | (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
// instruction descriptor of tcgen05.mma.kind::f16: fp16 x fp16 -> fp32, both operands K-major, M = 128 (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) { return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(kMlpBM >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
        "%28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace

// Observations -> normalised, clipped fp16 activations in operand layout.  grid = (m tiles, K chunks), 128 threads: 8 consecutive threads cover
// 64 consecutive inputs of one row (coalesced 256-byte reads), a thread writes one 16-byte core-matrix row.
__global__ void __launch_bounds__(kMlpThreads) dm_mlp_prep_kernel(MlpPrepParams P) {
    const int m0 = blockIdx.x * kMlpBM, c = blockIdx.y;
    __half* tile = P.tiles + (static_cast<size_t>(blockIdx.x) * P.NC + c) * kMlpATile;
#pragma unroll
    for (int i = 0; i < (kMlpBM * 8) / kMlpThreads; ++i) {
        const int u = threadIdx.x + i * kMlpThreads, row = u >> 3, k8 = u & 7;
        const int grow = m0 + row, k = c * kMlpBK + k8 * 8;
        __align__(16) __half h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float x = 0.f;
            if (grow < P.M && k + e < P.in_dim) {
                x = (P.obs[static_cast<size_t>(grow) * P.in_dim + k + e] - P.in_mean[k + e]) * P.in_istd[k + e];
                x = fminf(fmaxf(x, -P.in_clip), P.in_clip);
            }
            h[e] = __float2half_rn(x);
        }
        *reinterpret_cast<uint4*>(tile + ((k8 * (kMlpBM / 8) + (row >> 3)) * 64 + (row & 7) * 8)) = *reinterpret_cast<const uint4*>(h);
    }
}

// C[M x N] = act(A[M x K] W + b); grid = (M / 128, N / BN), block = 128 threads,
// dynamic shared memory = 4 stages x (A 16 KB + W hi/lo 2 x BN x 128 B) + 1 KB (barriers, TMEM slot)
template <int BN, bool LAST>
__global__ void __launch_bounds__(kMlpThreads, 1) dm_mlp_gemm_kernel(MlpGemmParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int kABytes = kMlpATile * 2;               // 16 KB
    constexpr int kWBytes = 2 * BN * kMlpBK * 2;         // hi + lo
    constexpr int kStage = kABytes + kWBytes;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + kMlpStages * kStage);   // [stages] both operands of the stage have landed
    uint64_t* bar_empty = bar_full + kMlpStages;                                         // [stages] the MMAs reading the stage have completed
    uint64_t* bar_acc = bar_empty + kMlpStages;                                          // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int mt = blockIdx.x, m0 = mt * kMlpBM, nt = blockIdx.y, n0 = nt * BN;
    const int NC = P.K / kMlpBK;
    constexpr int kTmemCols = BN < 32 ? 32 : BN;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kMlpStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        // ---- TMA producer: one bulk copy per operand and K-chunk (both are contiguous blocks in operand layout)
        const __half* a_src = P.a_tiles + static_cast<size_t>(mt) * NC * kMlpATile;
        const __half* w_src = P.w_tiles + static_cast<size_t>(nt) * NC * (kWBytes / 2);
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
            const int s = c % kMlpStages;
            if (c >= kMlpStages) mbar_wait(&bar_empty[s], ((c / kMlpStages) - 1) & 1);
            uint8_t* sA = smem_raw + s * kStage;
            mbar_expect_tx(&bar_full[s], kStage);
            bulk_g2s(sA, a_src + static_cast<size_t>(c) * kMlpATile, kABytes, &bar_full[s]);
            bulk_g2s(sA + kABytes, w_src + static_cast<size_t>(c) * (kWBytes / 2), kWBytes, &bar_full[s]);
        }
    } else if (tid == 32) {
        // ---- MMA issuer: 4 K-steps x (hi, lo) per chunk, accumulating in TMEM
        const uint32_t idesc = umma_idesc_f16(BN);
        constexpr uint32_t kALbo = (kMlpBM / 8) * 128, kBLbo = (BN / 8) * 128;
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
            const int s = c % kMlpStages;
            mbar_wait(&bar_full[s], (c / kMlpStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_u32(smem_raw + s * kStage), b0 = a0 + kABytes;
#pragma unroll
            for (int j = 0; j < kMlpBK / 16; ++j) {
                const uint64_t ad = umma_desc(a0 + j * 2 * kALbo, kALbo, 128);
                umma_f16(tmem_base, ad, umma_desc(b0 + j * 2 * kBLbo, kBLbo, 128), idesc, (c > 0 || j > 0) ? 1u : 0u);
                umma_f16(tmem_base, ad, umma_desc(b0 + BN * kMlpBK * 2 + j * 2 * kBLbo, kBLbo, 128), idesc, 1u);
            }
            umma_commit(&bar_empty[s]);      // arrives when the MMAs issued so far have completed: the stage may be refilled
        }
        umma_commit(bar_acc);
    }
    // ---- all MMAs done (a commit tracks every MMA issued before it)
    mbar_wait(bar_acc, 0);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: thread = one environment (TMEM lane), 32 columns at a time
    const int r = warp * 32 + (tid & 31), row = m0 + r;
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + j * 32, v);
        if constexpr (LAST) {
            if (row < P.M) {
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int n = n0 + j * 32 + e;
                    if (n < P.out_dim) {
                        float a = __uint_as_float(v[e]) + P.bias[n];
                        if (P.noise) a += P.noise[static_cast<size_t>(row) * P.out_dim + n];
                        P.actions[static_cast<size_t>(row) * P.out_dim + n] = a * P.out_std[n] + P.out_mean[n];
                    }
                }
            }
        } else {
            // next layer's operand tiles: K index = this layer's column; rows past M carry relu(bias) (never read back as results)
            __align__(16) __half h[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) h[e] = __float2half_rn(fmaxf(__uint_as_float(v[e]) + P.bias[n0 + j * 32 + e], 0.f));
            const int n = n0 + j * 32, kc = n >> 6, k8b = (n & 63) >> 3;
            __half* tile = P.out_tiles + (static_cast<size_t>(mt) * (P.N >> 6) + kc) * kMlpATile;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(tile + (((k8b + q) * (kMlpBM / 8) + (r >> 3)) * 64 + (r & 7) * 8)) = reinterpret_cast<const uint4*>(h)[q];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

int dm_mlp_smem_bytes(int bn) { return kMlpStages * (kMlpATile * 2 + 2 * bn * kMlpBK * 2) + 1024; }

template __global__ void dm_mlp_gemm_kernel<128, false>(MlpGemmParams);
template __global__ void dm_mlp_gemm_kernel<32, true>(MlpGemmParams);
template __global__ void dm_mlp_gemm_kernel<64, true>(MlpGemmParams);   // action sizes 33..64 (dog3d: 58)

}  // namespace dmk
