// KV-cached Transformer decoder step + greedy bookkeeping for S2STransformer{Greedy,Beam}Searcher.
//
// The reference (decoders/seq2seq.py:360-367,1929-1934 -> TransformerASR.decode
// lobes/models/transformer/TransformerASR.py:426-473 -> Transformer.py:751-834,915-963) re-embeds and
// re-runs all decoder layers over the WHOLE prefix every step and re-projects the encoder memory to K/V in
// every layer every step. Here: cross-attention K/V are projected once per utterance (one tcgen05 GEMM
// over all layers), self-attention K/V are appended to a cache, and one step touches only the newest token.
// Numerically the step computes exactly the reference's last-position output.
//
// All step kernels read the current step index from device memory so that a single captured CUDA graph
// can be replayed for every step.
#include <algorithm>

#include "common.cuh"
#include "sbk_internal.h"

namespace sbk {

__device__ __forceinline__ void mma16816_d(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Programmatic dependent launch: every decode-step kernel lets its successor start launching immediately and
// waits for its predecessor's memory only right before it touches activations, so launch latency and the
// weight prefetch of kernel N+1 overlap the tail of kernel N.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static bool g_use_pdl = true;
void set_pdl(bool on) { g_use_pdl = on; }

template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// --------------------------------------------------------------------------- skinny GEMM (weight streaming)
// y[n_rows, N] = epi( A[n_rows, K] x W[N, K]^T (fp16) + bias ).  n_rows is the number of live hypotheses
// (32..320): the cost is streaming W once plus a chain of dependent L2 round trips, so one CTA owns 8 output
// columns x 32 rows, its 8 warps split K, and every warp issues all the loads of a chunk of UNR k-steps
// before the first mma.  A is either fp16 in global memory, or (LN variant) LayerNorm(x fp32) computed by the
// CTA itself into shared memory -- this fuses the decoder's pre-norms (Transformer.py:788-827) into the
// projection that consumes them.  Deterministic in-CTA split-K reduction through shared memory.
constexpr int SK_ROWS = 32;
constexpr int SK_WARPS = 8;

// UNR: k-steps whose loads are issued back to back; NT: 8-column tiles per CTA; VPL: LayerNorm-fused variant
// when > 0, with K == 128 * VPL (each lane holds VPL float4 of each of its 4 rows).
template <int UNR, int NT, int VPL>
__global__ void __launch_bounds__(SK_WARPS * 32, (VPL <= 4 ? 2 : 1)) skinny_gemm_kernel(const SkinnyArgs a) {
    constexpr bool LN = VPL > 0;
    __shared__ float red[SK_WARPS][SK_ROWS][8 * NT + 1];
    extern __shared__ __align__(16) __half a_sm[];  // LN: [32][K + 8]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const int n0 = blockIdx.x * 8 * NT;
    const int row0 = blockIdx.y * SK_ROWS;
    const int rows = min(SK_ROWS, a.n_rows - row0);
    const int k_per_warp = ((a.K / 16 + SK_WARPS - 1) / SK_WARPS) * 16;
    const int k_begin = warp * k_per_warp, k_end = min(a.K, k_begin + k_per_warp);
    pdl_trigger();

    float acc[2][NT][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0f;
    const __half* wrow[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j)
        wrow[j] = a.W + static_cast<size_t>(min(n0 + 8 * j + g, a.N - 1)) * a.ldw + 2 * c;  // clamp: N tail discarded
    // weights do not depend on the previous kernel: fetch the first chunk before waiting on it
    uint32_t bf[UNR][NT][2];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const int k = k_begin + 16 * u;
        const bool ok = k < k_end;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            bf[u][j][0] = ok ? __ldg(reinterpret_cast<const uint32_t*>(wrow[j] + k)) : 0u;
            bf[u][j][1] = ok ? __ldg(reinterpret_cast<const uint32_t*>(wrow[j] + k + 8)) : 0u;
        }
    }
    pdl_wait();

    const int astr = LN ? a.K + 8 : a.lda;
    const __half* abase = LN ? a_sm : a.A;
    if constexpr (LN) {
        // LayerNorm of this CTA's 32 rows: warp w owns rows w, w+8, w+16, w+24; all 4*VPL loads are issued before
        // the first reduction (fp32 statistics, two-pass), result fp16 in shared memory.
        float4 v[4][VPL];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float* xr = a.X + static_cast<size_t>(min(row0 + warp + 8 * r, a.n_rows - 1)) * a.K;
#pragma unroll
            for (int i = 0; i < VPL; ++i) v[r][i] = *reinterpret_cast<const float4*>(xr + 4 * (lane + 32 * i));
        }
        float4 gm[VPL], bt[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            gm[i] = __ldg(reinterpret_cast<const float4*>(a.ln_g + 4 * (lane + 32 * i)));
            bt[i] = __ldg(reinterpret_cast<const float4*>(a.ln_b + 4 * (lane + 32 * i)));
        }
        float s[4], q[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            s[r] = 0.0f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) s[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < 4; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
        const float inv_k = 1.0f / static_cast<float>(a.K);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            s[r] *= inv_k;  // mean
            q[r] = 0.0f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                v[r][i].x -= s[r]; v[r][i].y -= s[r]; v[r][i].z -= s[r]; v[r][i].w -= s[r];
                q[r] += (v[r][i].x * v[r][i].x + v[r][i].y * v[r][i].y) + (v[r][i].z * v[r][i].z + v[r][i].w * v[r][i].w);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < 4; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float rstd = rsqrtf(q[r] * inv_k + a.ln_eps);
            __half* dst = a_sm + (warp + 8 * r) * astr;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                __half2 h0 = __floats2half2_rn(v[r][i].x * rstd * gm[i].x + bt[i].x, v[r][i].y * rstd * gm[i].y + bt[i].y);
                __half2 h1 = __floats2half2_rn(v[r][i].z * rstd * gm[i].z + bt[i].z, v[r][i].w * rstd * gm[i].w + bt[i].w);
                uint2 u;
                u.x = *reinterpret_cast<uint32_t*>(&h0);
                u.y = *reinterpret_cast<uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(dst + 4 * (lane + 32 * i)) = u;
            }
        }
        __syncthreads();
    }
    const __half* arow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = LN ? (i * 8 + g) : min(row0 + i * 8 + g, a.n_rows - 1);
        arow[i] = abase + static_cast<size_t>(rr) * astr + 2 * c;
    }
    for (int k0 = k_begin; k0 < k_end; k0 += 16 * UNR) {
        uint32_t af[UNR][2][4];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int k = k0 + 16 * u;
            const bool ok = k < k_end;
            if (k0 != k_begin) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    bf[u][j][0] = ok ? __ldg(reinterpret_cast<const uint32_t*>(wrow[j] + k)) : 0u;
                    bf[u][j][1] = ok ? __ldg(reinterpret_cast<const uint32_t*>(wrow[j] + k + 8)) : 0u;
                }
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                af[u][mt][0] = ok ? *reinterpret_cast<const uint32_t*>(arow[2 * mt] + k) : 0u;
                af[u][mt][1] = ok ? *reinterpret_cast<const uint32_t*>(arow[2 * mt + 1] + k) : 0u;
                af[u][mt][2] = ok ? *reinterpret_cast<const uint32_t*>(arow[2 * mt] + k + 8) : 0u;
                af[u][mt][3] = ok ? *reinterpret_cast<const uint32_t*>(arow[2 * mt + 1] + k + 8) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                mma16816_d(acc[0][j], af[u][0], bf[u][j][0], bf[u][j][1]);
                mma16816_d(acc[1][j], af[u][1], bf[u][j][0], bf[u][j][1]);
            }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            red[warp][mt * 16 + g][8 * j + 2 * c] = acc[mt][j][0];
            red[warp][mt * 16 + g][8 * j + 2 * c + 1] = acc[mt][j][1];
            red[warp][mt * 16 + g + 8][8 * j + 2 * c] = acc[mt][j][2];
            red[warp][mt * 16 + g + 8][8 * j + 2 * c + 1] = acc[mt][j][3];
        }
    __syncthreads();
    const int step = a.step_ptr ? *a.step_ptr : 0;
#pragma unroll
    for (int it = 0; it < NT; ++it) {
        const int idx = threadIdx.x + it * SK_WARPS * 32;  // 32 rows x (8*NT) columns
        const int r = idx / (8 * NT), j = idx - r * (8 * NT);
        const int col = n0 + j;
        const int row = row0 + r;
        if (r < rows && col < a.N) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) v += red[w][r][j];
            if (a.bias) v += __ldg(a.bias + col);
            switch (a.epi) {
                case SK_F16: reinterpret_cast<__half*>(a.out)[static_cast<size_t>(row) * a.ldo + col] = __float2half_rn(v); break;
                case SK_F16_RELU:
                    reinterpret_cast<__half*>(a.out)[static_cast<size_t>(row) * a.ldo + col] = __float2half_rn(fmaxf(v, 0.0f));
                    break;
                case SK_F16_GELU:
                    reinterpret_cast<__half*>(a.out)[static_cast<size_t>(row) * a.ldo + col] = __float2half_rn(gelu_erf_f(v));
                    break;
                case SK_F32: reinterpret_cast<float*>(a.out)[static_cast<size_t>(row) * a.ldo + col] = v; break;
                case SK_RESID: reinterpret_cast<float*>(a.out)[static_cast<size_t>(row) * a.ldo + col] += v; break;
                case SK_QKV_CACHE: {
                    if (col < a.d) {
                        reinterpret_cast<__half*>(a.out)[static_cast<size_t>(row) * a.ldo + col] = __float2half_rn(v * a.q_scale);
                    } else if (col < 2 * a.d) {
                        a.kcache[(static_cast<size_t>(row) * a.S_max + step) * a.d + (col - a.d)] = __float2half_rn(v);
                    } else {
                        a.vcache[(static_cast<size_t>(row) * a.S_max + step) * a.d + (col - 2 * a.d)] = __float2half_rn(v);
                    }
                    break;
                }
            }
        }
    }
}

int skinny_gemm(const SkinnyArgs& a, cudaStream_t stream) {
    SBK_REQUIRE(a.K % 16 == 0 && a.lda % 2 == 0 && a.ldw % 2 == 0, "skinny_gemm: K %% 16 required (K=%d)", a.K);
    if (a.n_rows == 0) return SBK_OK;
    cudaError_t e;
    const int ry = ceil_div(a.n_rows, SK_ROWS);
    if (a.X != nullptr) {
        const size_t smem = static_cast<size_t>(SK_ROWS) * (a.K + 8) * 2;
        dim3 grid(ceil_div(a.N, 16), ry);
        static bool attr_done = false;  // static + dynamic shared memory exceeds the 48 KB default
        if (!attr_done) {
            cudaFuncSetAttribute(skinny_gemm_kernel<4, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
            cudaFuncSetAttribute(skinny_gemm_kernel<2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
            cudaFuncSetAttribute(skinny_gemm_kernel<4, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
            cudaFuncSetAttribute(skinny_gemm_kernel<8, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
            attr_done = true;
        }
        if (a.K == 512) e = launch_k(skinny_gemm_kernel<4, 2, 4>, grid, dim3(SK_WARPS * 32), smem, stream, a);
        else if (a.K == 256) e = launch_k(skinny_gemm_kernel<2, 2, 2>, grid, dim3(SK_WARPS * 32), smem, stream, a);
        else if (a.K == 768) e = launch_k(skinny_gemm_kernel<4, 2, 6>, grid, dim3(SK_WARPS * 32), smem, stream, a);
        else if (a.K == 1024) e = launch_k(skinny_gemm_kernel<8, 2, 8>, grid, dim3(SK_WARPS * 32), smem, stream, a);
        else {
            set_error("skinny_gemm(LN): d_model=%d not built (256/512/768/1024)", a.K);
            return SBK_ERR_UNSUPPORTED;
        }
    } else if (a.K <= 1024) {
        e = launch_k(skinny_gemm_kernel<4, 1, 0>, dim3(ceil_div(a.N, 8), ry), dim3(SK_WARPS * 32), 0, stream, a);
    } else {
        e = launch_k(skinny_gemm_kernel<8, 1, 0>, dim3(ceil_div(a.N, 8), ry), dim3(SK_WARPS * 32), 0, stream, a);
    }
    SBK_CUDA_CHECK(e);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

// --------------------------------------------------------------------------- decode-time attention (1 query / row)
// One CTA per (row, head), 4 warps split the keys (flash-decoding style): each warp scores its key slice with
// lanes parallel over keys, then accumulates p.V with 4 lane-groups over keys x 8 lanes over 16-byte dim chunks;
// the 4 partial (max, sum, out) triples are merged through shared memory.
// Self-attention: keys = cache positions [0, step]; cross-attention: keys = encoder frames [0, enc_len[utt]).
// (nn.MultiheadAttention semantics, scale 1/sqrt(d_h) already folded into q.)  head_dim == 64.
constexpr int DA_WARPS = 4;

__global__ void __launch_bounds__(DA_WARPS * 32) dec_attention_kernel(const DecAttnArgs a) {
    extern __shared__ float da_smem[];            // [max_keys] scores
    __shared__ float part_o[DA_WARPS][64];
    __shared__ float part_m[DA_WARPS], part_l[DA_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x, h = blockIdx.y;
    const int blk = r / a.rows_per_block;
    pdl_trigger();
    pdl_wait();
    int n_keys;
    if (a.n_keys_ptr) n_keys = *a.n_keys_ptr + 1;
    else n_keys = a.enc_len ? min(a.enc_len[blk], a.n_keys_fixed) : a.n_keys_fixed;
    const int per = (n_keys + DA_WARPS - 1) / DA_WARPS;
    const int kb = warp * per, ke = min(n_keys, kb + per);
    const __half* q = a.q + static_cast<size_t>(r) * a.ldq + h * 64;
    const __half* kbase = a.kbase + static_cast<size_t>(blk) * a.row_stride + h * 64;
    const __half* vbase = a.vbase + static_cast<size_t>(blk) * a.row_stride + h * 64;
    // full query vector in registers (every lane)
    float qf[64];
#pragma unroll
    for (int e = 0; e < 64; e += 8) {
        const uint4 qv = *reinterpret_cast<const uint4*>(q + e);
        const __half2* q2 = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(q2[t]);
            qf[e + 2 * t] = f.x;
            qf[e + 2 * t + 1] = f.y;
        }
    }
    float mx = -INFINITY;
    for (int j = kb + lane; j < ke; j += 32) {
        const __half* kr = kbase + static_cast<size_t>(j) * a.key_stride;
        uint4 kv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) kv[e] = *reinterpret_cast<const uint4*>(kr + e * 8);
        float dot = 0.0f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const __half2* k2 = reinterpret_cast<const __half2*>(&kv[e]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 kf = __half22float2(k2[t]);
                dot = fmaf(kf.x, qf[e * 8 + 2 * t], dot);
                dot = fmaf(kf.y, qf[e * 8 + 2 * t + 1], dot);
            }
        }
        da_smem[j] = dot;
        mx = fmaxf(mx, dot);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = kb + lane; j < ke; j += 32) {
        const float p = __expf(da_smem[j] - mx);
        da_smem[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    // p.V : lane group gq handles keys kb+gq, kb+gq+4, ...; lane%8 owns dims [8*(lane%8), +8)
    const int gq = lane >> 3, dl = (lane & 7) * 8;
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.0f;
    int j = kb + gq;
    for (; j + 12 < ke; j += 16) {  // 4 keys in flight per lane
        uint4 vv[4];
        float p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            vv[u] = *reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(j + 4 * u) * a.key_stride + dl);
            p[u] = da_smem[j + 4 * u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const __half2* v2 = reinterpret_cast<const __half2*>(&vv[u]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 vf = __half22float2(v2[t]);
                o[2 * t] = fmaf(p[u], vf.x, o[2 * t]);
                o[2 * t + 1] = fmaf(p[u], vf.y, o[2 * t + 1]);
            }
        }
    }
    for (; j < ke; j += 4) {
        const uint4 vv = *reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(j) * a.key_stride + dl);
        const float p = da_smem[j];
        const __half2* v2 = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 vf = __half22float2(v2[t]);
            o[2 * t] = fmaf(p, vf.x, o[2 * t]);
            o[2 * t + 1] = fmaf(p, vf.y, o[2 * t + 1]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
        o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
    }
    if (lane < 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) part_o[warp][dl + e] = o[e];
    }
    if (lane == 0) { part_m[warp] = mx; part_l[warp] = sum; }
    __syncthreads();
    if (threadIdx.x < 64) {
        float M = part_m[0];
#pragma unroll
        for (int w = 1; w < DA_WARPS; ++w) M = fmaxf(M, part_m[w]);
        float num = 0.0f, den = 0.0f;
#pragma unroll
        for (int w = 0; w < DA_WARPS; ++w) {
            const float sc = part_m[w] == -INFINITY ? 0.0f : __expf(part_m[w] - M);
            num += part_o[w][threadIdx.x] * sc;
            den += part_l[w] * sc;
        }
        a.out[static_cast<size_t>(r) * a.ldo + h * 64 + threadIdx.x] = __float2half_rn(num / den);
    }
}

int dec_attention(const DecAttnArgs& a, int n_rows, int max_keys, cudaStream_t stream) {
    SBK_REQUIRE(a.dh == 64, "dec_attention: head_dim=%d not built (64 only)", a.dh);
    if (n_rows == 0) return SBK_OK;
    const size_t smem = static_cast<size_t>(max_keys) * sizeof(float);
    SBK_REQUIRE(smem <= 40 * 1024, "dec_attention: too many keys (%d)", max_keys);
    DecAttnArgs b = a;
    b.n_keys_fixed = max_keys;
    SBK_CUDA_CHECK(launch_k(dec_attention_kernel, dim3(n_rows, a.H), dim3(DA_WARPS * 32), smem, stream, b));
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}


// --------------------------------------------------------------------------- greedy step bookkeeping
// decoders/seq2seq.py:226-257: argmax, fp32 log_softmax, has_ended |= (tok == eos); ended rows get
// log_probs = -inf (=> prediction eos, score 0 after :259-263) and keep feeding eos.
// One CTA per row. Writes tokens[r][step+1], pred[r][step], score[r][step], optional log-prob row, then the
// NEXT step's decoder input x[r] = emb[tok] * sqrt(d) + pe[step+1] (Transformer.py:966-995, :252-303) and
// advances this row's step counter (every kernel of the next step reads step[0] after this kernel is done).
__global__ void __launch_bounds__(256)
greedy_select_kernel(const float* __restrict__ logits, int V, int* __restrict__ step_arr, int eos, int* tokens,
                     int tok_stride, int* has_ended, int* ended_count, int* pred, float* score, int out_stride,
                     float* log_probs /* [n, L, V] or null */, int L, const float* __restrict__ emb,
                     const float* __restrict__ pe, int d, float sqrt_d, float* __restrict__ x_next) {
    __shared__ float s_val[8];
    __shared__ int s_idx[8];
    __shared__ float s_sum[8];
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x, step = step_arr[r];
    const float* lg = logits + static_cast<size_t>(r) * V;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const float v = lg[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    best = s_val[0]; bi = s_idx[0];
    for (int w = 1; w < 8; ++w)
        if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
    float sum = 0.0f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) sum += expf(lg[i] - best);
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.0f;
    for (int w = 0; w < 8; ++w) sum += s_sum[w];
    const float lse = best + logf(sum);
    const int was_ended = has_ended[r];
    const int ended = was_ended | (bi == eos ? 1 : 0);
    if (log_probs) {
        float* lp = log_probs + (static_cast<size_t>(r) * L + step) * V;
        for (int i = threadIdx.x; i < V; i += blockDim.x) lp[i] = ended ? -INFINITY : lg[i] - lse;
    }
    const int tok = ended ? eos : bi;
    __syncthreads();  // everyone has read has_ended[r] / step_arr[r] before thread 0 updates them
    if (threadIdx.x == 0) {
        tokens[static_cast<size_t>(r) * tok_stride + step + 1] = tok;
        pred[static_cast<size_t>(r) * out_stride + step] = tok;
        score[static_cast<size_t>(r) * out_stride + step] = ended ? 0.0f : best - lse;
        if (ended && !was_ended) {
            has_ended[r] = 1;
            atomicAdd(ended_count, 1);
        }
        step_arr[r] = step + 1;
    }
    const float* e = emb + static_cast<size_t>(tok) * d;
    const float* p = pe + static_cast<size_t>(step + 1) * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x_next[static_cast<size_t>(r) * d + i] = e[i] * sqrt_d + p[i];
}

// tokens[r][0] = bos, step[r] = 0, x[r] = emb[bos] * sqrt(d) + pe[0]
__global__ void greedy_reset_kernel(int* tokens, int tok_stride, int bos, int* step_arr, int* has_ended,
                                    int* ended_count, const float* __restrict__ emb, const float* __restrict__ pe, int d,
                                    float sqrt_d, float* __restrict__ x) {
    const int r = blockIdx.x;
    if (threadIdx.x == 0) {
        tokens[static_cast<size_t>(r) * tok_stride] = bos;
        has_ended[r] = 0;
        step_arr[r] = 0;
        if (r == 0) *ended_count = 0;
    }
    const float* e = emb + static_cast<size_t>(bos) * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x[static_cast<size_t>(r) * d + i] = e[i] * sqrt_d + pe[i];
}

int greedy_reset(int* tokens, int tok_stride, int n_rows, int bos, int* step_arr, int* has_ended, int* ended_count,
                 const float* emb, const float* pe, int d, float* x, cudaStream_t stream) {
    if (n_rows == 0) return SBK_OK;
    greedy_reset_kernel<<<n_rows, 128, 0, stream>>>(tokens, tok_stride, bos, step_arr, has_ended, ended_count, emb, pe, d,
                                                   sqrtf(static_cast<float>(d)), x);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

int greedy_select(const float* logits, int n_rows, int V, int* step_arr, int eos, int* tokens, int tok_stride,
                  int* has_ended, int* ended_count, int* pred, float* score, int out_stride, float* log_probs, int L,
                  const float* emb, const float* pe, int d, float* x_next, cudaStream_t stream) {
    if (n_rows == 0) return SBK_OK;
    SBK_CUDA_CHECK(launch_k(greedy_select_kernel, dim3(n_rows), dim3(256), 0, stream, logits, V, step_arr, eos, tokens,
                            tok_stride, has_ended, ended_count, pred, score, out_stride, log_probs, L, emb, pe, d,
                            sqrtf(static_cast<float>(d)), x_next));
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

}  // namespace sbk
