// tcgen05 / TMEM / TMA GEMM for the Conformer encoder (and decoder prefill):
//
//     out = epilogue( A[M,K] (fp16, K-major) x W[N,K]^T (fp16, K-major) ), fp32 accumulate in TMEM
//
// Replaces the reference's F.linear / nn.Linear / Conv1d(k=1) call sites
// (nnet/attention.py:623,739,932-936,1344; Conformer.py:126-157; TransformerASR.py:308-316).
//
// One 128 x BN output tile per CTA, 192 threads:
//   warp 0      : TMA producer (one elected lane), 128B-swizzled K-major tiles of 64 halfs
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16)
//   warps 2..5  : epilogue; warp w drains TMEM lanes 32*(w%4).. with tcgen05.ld 32x32b.x32
// smem ring of STAGES (A 16 KB + B BN*128 B) guarded by full/empty mbarriers; the MMA
// completion is published with tcgen05.commit.  Two CTAs fit per SM (3 x 32 KB stages,
// 128 TMEM columns each) so one CTA's epilogue overlaps the other's main loop.
#include "common.cuh"
#include "sbk_internal.h"

namespace sbk {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
};

// Apply the epilogue to 32 consecutive accumulator columns of one row.
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& e, const uint32_t (&acc)[32], int row, int col0,
                                               int M, int N) {
    if (row >= M || col0 >= N) return;
    const bool full = (col0 + 32 <= N);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
    if (e.bias != nullptr) {
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < N) v[j] += __ldg(e.bias + col0 + j);
        }
    }
    switch (e.mode) {
        case EPI_F16: {
            if (e.act == ACT_SILU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
            } else if (e.act == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
            }
            __half* o = reinterpret_cast<__half*>(e.out) + static_cast<size_t>(row) * e.ldo + col0;
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    __half2 h0 = __floats2half2_rn(v[j], v[j + 1]), h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
                    __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]), h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
                    uint4 u;
                    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(o + j) = u;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < N) o[j] = __float2half_rn(v[j]);
            }
            break;
        }
        case EPI_F32: {
            float* o = reinterpret_cast<float*>(e.out) + static_cast<size_t>(row) * e.ldo + col0;
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < N) o[j] = v[j];
            }
            break;
        }
        case EPI_RESID: {  // out = resid + alpha * (acc + bias); masked rows contribute 0
            float alpha = e.alpha;
            if (e.row_lens != nullptr) {
                const int b = row / e.T, t = row - b * e.T;
                if (t >= e.row_lens[b]) alpha = 0.0f;
            }
            const float* r = e.resid + static_cast<size_t>(row) * e.ldo + col0;
            float* o = reinterpret_cast<float*>(e.out) + static_cast<size_t>(row) * e.ldo + col0;
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(r + j);
                    *reinterpret_cast<float4*>(o + j) = make_float4(fmaf(alpha, v[j], x.x), fmaf(alpha, v[j + 1], x.y),
                                                                    fmaf(alpha, v[j + 2], x.z), fmaf(alpha, v[j + 3], x.w));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < N) o[j] = fmaf(alpha, v[j], r[j]);
            }
            break;
        }
        case EPI_GLU: {  // weight rows pre-interleaved [16 values | 16 gates] per 32 columns
            float* o = reinterpret_cast<float*>(e.out) + static_cast<size_t>(row) * e.ldo + (col0 >> 1);
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(o + j) =
                    make_float4(v[j] * sigmoid_f(v[j + 16]), v[j + 1] * sigmoid_f(v[j + 17]),
                                v[j + 2] * sigmoid_f(v[j + 18]), v[j + 3] * sigmoid_f(v[j + 19]));
            break;
        }
        case EPI_ROPE: {  // columns = per-head [q(dh) | k(dh) | v(dh)], dh % 32 == 0
            const int dh = e.head_dim;
            const int within = col0 % (3 * dh);
            const int sect = within / dh;  // 0 q, 1 k, 2 v
            if (sect < 2) {
                const int t = row % e.T;
                const int p0 = (within - sect * dh) >> 1;
                const float* cs = e.rope_cos + static_cast<size_t>(t) * (dh >> 1) + p0;
                const float* sn = e.rope_sin + static_cast<size_t>(t) * (dh >> 1) + p0;
                const float sc = sect == 0 ? e.alpha : 1.0f;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float c = __ldg(cs + (j >> 1)), s = __ldg(sn + (j >> 1));
                    const float x0 = v[j], x1 = v[j + 1];
                    v[j] = (x0 * c - x1 * s) * sc;
                    v[j + 1] = (x1 * c + x0 * s) * sc;
                }
            }
            __half* o = reinterpret_cast<__half*>(e.out) + static_cast<size_t>(row) * e.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                __half2 h0 = __floats2half2_rn(v[j], v[j + 1]), h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
                __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]), h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(o + j) = u;
            }
            break;
        }
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, (BN <= 128 ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmEpilogue epi, int M, int N, int K) {
    using S = GemmSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * GEMM_BM;
    const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_ptr, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], S::STAGE_BYTES);
                uint8_t* a_dst = smem + s * S::STAGE_BYTES;
                tma_load_2d(a_dst, &tmap_a, &full_bar[s], kb * GEMM_BK, m0);
                tma_load_2d(a_dst + S::A_BYTES, &tmap_b, &full_bar[s], kb * GEMM_BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(GEMM_BM, BN, 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
                const uint64_t da = make_kmajor_sw128_desc(a_addr);
                const uint64_t db = make_kmajor_sw128_desc(a_addr + S::A_BYTES);
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k)  // +32 B per UMMA_K=16 halfs -> +2 in (addr>>4)
                    umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t acc[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, acc);
            tmem_ld_wait();
            epilogue_chunk(epi, acc, row, n0 + c * 32, M, N);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BN);
}

template <int BN, int STAGES>
static int launch_gemm(const void* A, int lda, const void* W, int ldw, const GemmEpilogue& epi, int M, int N, int K,
                       cudaStream_t stream) {
    using S = GemmSmem<BN, STAGES>;
    CUtensorMap ta, tb;
    int rc = make_tmap_2d_f16(&ta, A, M, K, lda, GEMM_BM, GEMM_BK);
    if (rc) return rc;
    rc = make_tmap_2d_f16(&tb, W, N, K, ldw, BN, GEMM_BK);
    if (rc) return rc;
    auto kern = gemm_tc_kernel<BN, STAGES>;
    SBK_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    dim3 grid(ceil_div(N, BN), ceil_div(M, GEMM_BM));
    GemmProfile* prof = gemm_profile();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof->enabled) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, stream);
    }
    kern<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(ta, tb, epi, M, N, K);
    if (prof->enabled) {
        cudaEventRecord(e1, stream);
        prof->ev.push_back(e0);
        prof->ev.push_back(e1);
        prof->flops.push_back(2.0 * M * N * K);
    }
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

int gemm_f16(const void* A, int lda, const void* W, int ldw, const GemmEpilogue& epi, int M, int N, int K,
             cudaStream_t stream) {
    SBK_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_f16: empty problem M=%d N=%d K=%d", M, N, K);
    SBK_REQUIRE((K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0, "gemm_f16: K/lda/ldw must be multiples of 8");
    SBK_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                "gemm_f16: operands must be 16-byte aligned");
    if (epi.mode == EPI_GLU || epi.mode == EPI_ROPE)
        SBK_REQUIRE(N % 32 == 0, "gemm_f16: GLU/RoPE epilogues need N %% 32 == 0");
    if (epi.mode == EPI_ROPE) SBK_REQUIRE(epi.head_dim % 32 == 0, "gemm_f16: RoPE epilogue needs head_dim %% 32 == 0");
    return launch_gemm<128, 3>(A, lda, W, ldw, epi, M, N, K, stream);
}

}  // namespace sbk
