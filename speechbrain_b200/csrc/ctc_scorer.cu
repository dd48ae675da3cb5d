// CTC prefix scorer for joint CTC/attention beam search (full-vocabulary scoring), sm_100a.
//
// Reference: speechbrain/decoders/ctc.py:46-295 (CTCPrefixScore.__init__/forward_step/permute_mem) driven by
// CTCScorer (decoders/scorer.py:183-249) as a *full* scorer of ScorerBuilder.score (:1221-1268), ctc_window_size = 0.
//
// The reference materialises the forward variables r (T, 2, n_bh, V) for EVERY candidate token at every step (1.3 GB
// at B=32, beam 4) and then gathers the `beam` survivors (permute_mem).  Here the prefix probability psi of all
// n_bh x V candidates is computed with the recurrence held in registers (one thread per (hypothesis, token), frames
// streamed from the masked CTC log-posteriors x[B, T, V] -- coalesced over the token axis), nothing of r is stored; after
// the beam kernel has picked the survivors, `ctc_update` re-runs the recurrence for just those n_bh (parent, token)
// pairs and writes their forward variables (T x 2 per hypothesis) for the next step.  HBM-bound: x is read once per
// hypothesis row per step (L2-resident across the beams of an utterance).
//
// State per hypothesis row (ping-pong by step parity):  rsum[t] = logsumexp(r_nb[t], r_b[t]),  rb[t] = r_b[t],
// psi_prev = psi of the prefix itself.
#include "common.cuh"
#include "sbk_internal.h"

namespace sbk {

namespace {

constexpr float CTC_NEG = -1e20f;  // CTCPrefixScore.minus_inf (ctc.py:54)

// log(exp(a) + exp(b)), accurate version (state update) and fast version (full-vocabulary scoring)
__device__ __forceinline__ float logaddexp_acc(float a, float b) {
    const float m = fmaxf(a, b);
    return m + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ float logaddexp_fast(float a, float b) {
    const float m = fmaxf(a, b);
    return m + __logf(1.0f + __expf(-fabsf(a - b)));
}

// In place: x[b, t, :] = log_softmax(x[b, t, :]); frames t >= enc_len[b]: minus_inf everywhere, 0 at index 0
// (ctc.py:59-62 hard-codes channel 0 there); xb[b, t] = x[b, t, blank].
__global__ void __launch_bounds__(256)
ctc_logsoftmax_mask_kernel(float* __restrict__ x, float* __restrict__ xb, const int* __restrict__ enc_len, int T, int V,
                           int blank) {
    __shared__ float s_red[8];
    const int row = blockIdx.x, b = row / T, t = row - b * T, tid = threadIdx.x;
    float* xr = x + static_cast<size_t>(row) * V;
    if (t >= enc_len[b]) {
        for (int j = tid; j < V; j += 256) xr[j] = (j == 0) ? 0.0f : CTC_NEG;
        if (tid == 0) xb[row] = (blank == 0) ? 0.0f : CTC_NEG;
        return;
    }
    float mx = -INFINITY;
    for (int j = tid; j < V; j += 256) mx = fmaxf(mx, xr[j]);
    mx = warp_max(mx);
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_red[w]);
    __syncthreads();
    float sm = 0.0f;
    for (int j = tid; j < V; j += 256) sm += expf(xr[j] - mx);
    sm = warp_sum(sm);
    if ((tid & 31) == 0) s_red[tid >> 5] = sm;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < 8; ++w) tot += s_red[w];
    const float lse = mx + logf(tot);
    for (int j = tid; j < V; j += 256) {
        const float v = xr[j] - lse;
        xr[j] = v;
        if (j == blank) xb[row] = v;
    }
}

// states = None (ctc.py:112-126): r_nb = minus_inf, r_b[t] = cumsum_t x[t, blank]; psi_prev = 0.
__global__ void ctc_init_kernel(const float* __restrict__ xb, int T, int beam, float* __restrict__ rsum,
                                float* __restrict__ rb, float* __restrict__ psi_prev) {
    extern __shared__ float s_cum[];
    const int row = blockIdx.x, b = row / beam;
    if (threadIdx.x == 0) {
        float acc = 0.0f;
        for (int t = 0; t < T; ++t) { acc += xb[static_cast<size_t>(b) * T + t]; s_cum[t] = acc; }
        psi_prev[row] = 0.0f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        rb[static_cast<size_t>(row) * T + t] = s_cum[t];
        rsum[static_cast<size_t>(row) * T + t] = logaddexp_acc(CTC_NEG, s_cum[t]);
    }
}

// The step index comes from DEVICE memory (the beam search's per-row step counters) so that one captured CUDA graph of a
// whole search step can be replayed for every step; the ping-pong halves of the state follow the step's parity.
struct CtcArgs {
    const float* x; const float* xb;         // [B, T, V], [B, T]
    float* rsum_base; float* rb_base;        // [2][n_bh, T]: half (step & 1) holds the prefixes being extended
    float* psi_base;                         // [2][n_bh]
    const int* enc_len;                      // [B]
    const int* hist_tok; const int* hist_pred;  // beam history [steps, n_bh]
    const int* step_ptr; int step_adj;       // step = step_ptr[row] + step_adj (the update runs after the counters advanced)
    int n_bh, bos, T, V, beam, blank, eos;
    float weight; float* out; int accumulate;   // score kernel: out[n_bh, V] (+)= weight * (psi - psi_prev)
};

// forward_step (ctc.py:80-249), candidates = None: thread (row, c) runs Alg.2 of Watanabe et al. over the frames.
constexpr int CTC_THREADS = 128;
__global__ void __launch_bounds__(CTC_THREADS) ctc_score_kernel(const CtcArgs a) {
    extern __shared__ float smem[];
    float* s_rsum = smem;
    float* s_rb = smem + a.T;
    float* s_xb = smem + 2 * a.T;
    const int row = blockIdx.y, b = row / a.beam, T = a.T, V = a.V;
    const int step = a.step_ptr[row] + a.step_adj;
    const size_t half = static_cast<size_t>(step & 1);
    const float* rsum_in = a.rsum_base + half * a.n_bh * T;
    const float* rb_in = a.rb_base + half * a.n_bh * T;
    const float* psi_prev = a.psi_base + half * a.n_bh;
    for (int t = threadIdx.x; t < T; t += CTC_THREADS) {
        s_rsum[t] = rsum_in[static_cast<size_t>(row) * T + t];
        s_rb[t] = rb_in[static_cast<size_t>(row) * T + t];
        s_xb[t] = a.xb[static_cast<size_t>(b) * T + t];
    }
    __syncthreads();
    const int c = blockIdx.x * CTC_THREADS + threadIdx.x;
    if (c >= V) return;
    const int last_char = step == 0 ? a.bos : a.hist_tok[static_cast<size_t>(step - 1) * a.n_bh + row];
    const float* xc = a.x + static_cast<size_t>(b) * T * V + c;
    float psi;
    if (c == a.blank && a.eos != a.blank) {
        psi = CTC_NEG;
    } else if (c == a.eos) {
        psi = s_rsum[a.enc_len[b] - 1];                                 // Alg.2-3
    } else {
        const float* phi = (c == last_char) ? s_rb : s_rsum;            // Alg.2-10
        float r_nb, r_b = CTC_NEG;
        int start;
        if (step == 0) { r_nb = xc[0]; start = 1; }                     // Alg.2-6
        else { r_nb = CTC_NEG; start = step; }
        float pm = r_nb, ps = 1.0f;                                     // running logsumexp, seeded with psi_init
#pragma unroll 4
        for (int t = start; t < T; ++t) {
            const float ph = phi[t - 1];
            const float xn = xc[static_cast<size_t>(t) * V];
            const float nb = logaddexp_fast(r_nb, ph) + xn;             // Alg.2-11
            const float bl = logaddexp_fast(r_nb, r_b) + s_xb[t];       // Alg.2-12
            const float term = ph + xn;                                 // Alg.2-13
            if (term > pm) { ps = ps * __expf(pm - term) + 1.0f; pm = term; }
            else ps += __expf(term - pm);
            r_nb = nb; r_b = bl;
        }
        psi = pm + __logf(ps);
    }
    const float sc = a.weight * (psi - psi_prev[row]);
    float* o = a.out + static_cast<size_t>(row) * V + c;
    *o = a.accumulate ? *o + sc : sc;
}

// permute_mem (ctc.py:251-295) without the (T, 2, n_bh, V) tensor: re-run the recurrence for the chosen (parent, token)
// of every new hypothesis and keep its forward variables.
__global__ void __launch_bounds__(128) ctc_update_kernel(const CtcArgs a) {
    extern __shared__ float smem[];
    const int T = a.T, V = a.V;
    float* s_phi = smem;
    float* s_xn = smem + T;
    float* s_xb = smem + 2 * T;
    float* s_nb = smem + 3 * T;
    float* s_bl = smem + 4 * T;
    const int row = blockIdx.x, b = row / a.beam;
    const int step = a.step_ptr[row] + a.step_adj;
    const size_t half = static_cast<size_t>(step & 1), other = half ^ 1;
    const float* rsum_in = a.rsum_base + half * a.n_bh * T;
    const float* rb_in = a.rb_base + half * a.n_bh * T;
    float* rsum_out = a.rsum_base + other * a.n_bh * T;
    float* rb_out = a.rb_base + other * a.n_bh * T;
    float* psi_out = a.psi_base + other * a.n_bh;
    const size_t h = static_cast<size_t>(step) * a.n_bh + row;
    const int tok = a.hist_tok[h], prow = a.hist_pred[h];
    const int last_char = step == 0 ? a.bos : a.hist_tok[static_cast<size_t>(step - 1) * a.n_bh + prow];
    const float* phi = (tok == last_char) ? rb_in : rsum_in;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        s_phi[t] = phi[static_cast<size_t>(prow) * T + t];
        s_xn[t] = a.x[(static_cast<size_t>(b) * T + t) * V + tok];
        s_xb[t] = a.xb[static_cast<size_t>(b) * T + t];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float r_nb, r_b = CTC_NEG;
        int start;
        if (step == 0) { r_nb = s_xn[0]; start = 1; }
        else { r_nb = CTC_NEG; start = step; }
        for (int t = 0; t < start - 1 && t < T; ++t) { s_nb[t] = CTC_NEG; s_bl[t] = CTC_NEG; }
        if (start - 1 < T) { s_nb[start - 1] = r_nb; s_bl[start - 1] = CTC_NEG; }
        float pm = r_nb, ps = 1.0f;
        for (int t = start; t < T; ++t) {
            const float ph = s_phi[t - 1], xn = s_xn[t];
            const float nb = logaddexp_acc(r_nb, ph) + xn;
            const float bl = logaddexp_acc(r_nb, r_b) + s_xb[t];
            const float term = ph + xn;
            if (term > pm) { ps = ps * expf(pm - term) + 1.0f; pm = term; }
            else ps += expf(term - pm);
            r_nb = nb; r_b = bl;
            s_nb[t] = nb; s_bl[t] = bl;
        }
        float psi = pm + logf(ps);
        if (tok == a.eos) psi = rsum_in[static_cast<size_t>(prow) * T + a.enc_len[b] - 1];
        if (tok == a.blank && a.eos != a.blank) psi = CTC_NEG;
        psi_out[row] = psi;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        rb_out[static_cast<size_t>(row) * T + t] = s_bl[t];
        rsum_out[static_cast<size_t>(row) * T + t] = logaddexp_acc(s_nb[t], s_bl[t]);
    }
}

}  // namespace

// --------------------------------------------------------------------------- CTC greedy (EncoderASR, decoders/ctc.py:335-378)
// One CTA per frame: optional in-place log_softmax of the row (what the recipe's `log_softmax` module does after ctc_lin) and
// its arg-max (first index on ties, like torch.max).
__global__ void __launch_bounds__(256)
rows_logsoftmax_argmax_kernel(float* __restrict__ x, int V, int do_logsoftmax, int* __restrict__ idx) {
    __shared__ float s_val[8];
    __shared__ int s_idx[8];
    const int row = blockIdx.x, tid = threadIdx.x;
    float* xr = x + static_cast<size_t>(row) * V;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = tid; j < V; j += 256) {
        const float v = xr[j];
        if (v > best || (v == best && j < bi)) { best = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
    __syncthreads();
    best = s_val[0]; bi = s_idx[0];
    for (int w = 1; w < 8; ++w)
        if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
    if (idx != nullptr && tid == 0) idx[row] = bi;
    if (!do_logsoftmax) return;
    __syncthreads();
    float sm = 0.0f;
    for (int j = tid; j < V; j += 256) sm += expf(xr[j] - best);
    sm = warp_sum(sm);
    if ((tid & 31) == 0) s_val[tid >> 5] = sm;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < 8; ++w) tot += s_val[w];
    const float lse = best + logf(tot);
    for (int j = tid; j < V; j += 256) xr[j] -= lse;
}

int rows_logsoftmax_argmax(float* x, int rows, int V, bool do_logsoftmax, int* idx, cudaStream_t stream) {
    if (rows == 0) return SBK_OK;
    rows_logsoftmax_argmax_kernel<<<rows, 256, 0, stream>>>(x, V, do_logsoftmax ? 1 : 0, idx);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

int ctc_prefix_reset(float* x, float* xb, const int* enc_len, int B, int T, int V, int blank, int beam, float* rsum, float* rb,
                     float* psi_prev, cudaStream_t stream) {
    SBK_REQUIRE(T >= 1 && T * 5 * 4 <= 200 * 1024, "ctc scorer: T=%d out of range", T);
    ctc_logsoftmax_mask_kernel<<<B * T, 256, 0, stream>>>(x, xb, enc_len, T, V, blank);
    SBK_LAUNCH_CHECK();
    ctc_init_kernel<<<B * beam, 128, T * 4, stream>>>(xb, T, beam, rsum, rb, psi_prev);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

static CtcArgs make_args(const CtcStep& p, int step_adj) {
    CtcArgs a;
    a.x = p.x; a.xb = p.xb; a.rsum_base = p.rsum_base; a.rb_base = p.rb_base; a.psi_base = p.psi_base; a.enc_len = p.enc_len;
    a.hist_tok = p.hist_tok; a.hist_pred = p.hist_pred; a.step_ptr = p.step_ptr; a.step_adj = step_adj;
    a.n_bh = p.n_bh; a.bos = p.bos; a.T = p.T; a.V = p.V;
    a.beam = p.beam; a.blank = p.blank; a.eos = p.eos; a.weight = p.weight; a.out = p.out; a.accumulate = p.accumulate;
    return a;
}

int ctc_prefix_score(const CtcStep& p, cudaStream_t stream) {
    const CtcArgs a = make_args(p, 0);   // runs before the beam kernel advances the step counters
    static bool attr = false;
    if (!attr) {
        SBK_CUDA_CHECK(cudaFuncSetAttribute(ctc_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SBK_CUDA_CHECK(cudaFuncSetAttribute(ctc_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    ctc_score_kernel<<<dim3(ceil_div(p.V, CTC_THREADS), p.n_bh), CTC_THREADS, 3 * p.T * 4, stream>>>(a);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

int ctc_prefix_update(const CtcStep& p, cudaStream_t stream) {
    const CtcArgs a = make_args(p, -1);  // runs after it
    ctc_update_kernel<<<p.n_bh, 128, 5 * p.T * 4, stream>>>(a);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

}  // namespace sbk
