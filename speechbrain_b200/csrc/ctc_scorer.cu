// CTC prefix scorer for joint CTC/attention beam search (full-vocabulary scoring), sm_100a.
//
// Reference: speechbrain/decoders/ctc.py:46-295 (CTCPrefixScore.__init__/forward_step/permute_mem) driven by
// CTCScorer (decoders/scorer.py:183-249) as a *full* scorer of ScorerBuilder.score (:1221-1268), ctc_window_size = 0.
//
// The reference materialises the forward variables r (T, 2, n_bh, V) for EVERY candidate token at every step (1.3 GB
// at B=32, beam 4) and then gathers the `beam` survivors (permute_mem).  The score it returns, though, only needs
//   psi(h, c) = logsumexp_t( phi_h[t-1] + x[t, c] )        phi_h = the PARENT's forward variables (Alg.2-10/13),
// i.e. a log-semiring product [n_bh, T] x [T, V] per utterance.  Here it is evaluated in the linear domain: the
// posteriors are exponentiated once per utterance (xlin = exp(x), next to x), each hypothesis' phi is shifted by its
// own maximum and exponentiated once per step into shared memory, and thread (utterance, token) accumulates
// sum_t A_h[t] * xlin[t, c] for all the `beam` hypotheses of its utterance at once -- no transcendental in the inner
// loop, xlin read once per utterance per step (coalesced over the token axis) instead of once per hypothesis.  A sum
// that underflows (every term more than ~e^-69 below the row maximum) is redone for that one (hypothesis, token) in
// the log domain, so the result never depends on the fp32 exponent range.  Nothing of r is stored; after the beam
// kernel has picked the survivors, `ctc_update` runs the recurrence for just those n_bh (parent, token) pairs and
// writes their forward variables (T x 2 per hypothesis) for the next step.
//
// State per hypothesis row (ping-pong by step parity):  rsum[t] = logsumexp(r_nb[t], r_b[t]),  rb[t] = r_b[t],
// psi_prev = psi of the prefix itself.
#include "common.cuh"
#include "sbk_internal.h"

namespace sbk {

namespace {

constexpr float CTC_NEG = -1e20f;  // CTCPrefixScore.minus_inf (ctc.py:54)

// log(exp(a) + exp(b))
__device__ __forceinline__ float logaddexp_acc(float a, float b) {
    const float m = fmaxf(a, b);
    return m + log1pf(expf(-fabsf(a - b)));
}

// In place: x[b, t, :] = log_softmax(x[b, t, :]); frames t >= enc_len[b]: minus_inf everywhere, 0 at index 0
// (ctc.py:59-62 hard-codes channel 0 there); xb[b, t] = x[b, t, blank].
__global__ void __launch_bounds__(256)
ctc_logsoftmax_mask_kernel(float* __restrict__ x, float* __restrict__ xlin, float* __restrict__ xb,
                           const int* __restrict__ enc_len, int T, int V, int blank) {
    __shared__ float s_red[8];
    const int row = blockIdx.x, b = row / T, t = row - b * T, tid = threadIdx.x;
    float* xr = x + static_cast<size_t>(row) * V;
    float* xl = xlin + static_cast<size_t>(row) * V;
    if (t >= enc_len[b]) {
        for (int j = tid; j < V; j += 256) { xr[j] = (j == 0) ? 0.0f : CTC_NEG; xl[j] = (j == 0) ? 1.0f : 0.0f; }
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
        xl[j] = expf(v);
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
    const float* xlin;                       // exp(x)
    float* rsum_base; float* rb_base;        // [2][n_bh, T]: half (step & 1) holds the prefixes being extended
    float* psi_base;                         // [2][n_bh]
    const int* enc_len;                      // [B]
    const int* hist_tok; const int* hist_pred;  // beam history [steps, n_bh]
    const int* step_ptr; int step_adj;       // step = step_ptr[row] + step_adj (the update runs after the counters advanced)
    int n_bh, bos, T, V, beam, blank, eos;
    float weight; float* out; int accumulate;   // score kernel: out[n_bh, V] (+)= weight * (psi - psi_prev)
};

// forward_step (ctc.py:80-249), candidates = None, for the R hypotheses blockIdx.y * R .. + R of one utterance (R divides
// the beam width) and 128 tokens.  Shared memory: A[2][R][Tp] -- variant 0 from rsum (Alg.2-10, c != last token), variant 1
// from r_b (c == last token); A[.][h][t] = exp(phi_h[t-1] - M_h) for t >= max(step, 1), and at step 0 the seed
// psi_init = x[0, c] (Alg.2-6) is the extra term A[.][h][0] = exp(-M_h) of the same sum.
// Threads: 128 tokens x CTC_TSPLIT frame groups (group g takes the 4-frame chunks g, g + CTC_TSPLIT, ...): with one thread
// per token the kernel is a chain of dependent memory round trips at a quarter of the SM's warp slots; the partial sums of the
// frame groups meet in shared memory.
constexpr int CTC_THREADS = 128;
constexpr int CTC_TSPLIT = 4;
template <int R>
__global__ void __launch_bounds__(CTC_THREADS * CTC_TSPLIT) ctc_score_kernel(const CtcArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_M[2][R];
    __shared__ int s_last[R];
    constexpr int NT = CTC_THREADS * CTC_TSPLIT;
    const int T = a.T, V = a.V, Tp = (T + 3) & ~3;
    float* s_part = smem + 2 * R * Tp;   // [CTC_TSPLIT - 1][R][CTC_THREADS] partial sums of the frame groups 1..
    const int row0 = blockIdx.y * R, b = row0 / a.beam;
    const int step = a.step_ptr[row0] + a.step_adj;
    const size_t half = static_cast<size_t>(step & 1);
    const float* rsum_in = a.rsum_base + half * a.n_bh * T;
    const float* rb_in = a.rb_base + half * a.n_bh * T;
    const float* psi_prev = a.psi_base + half * a.n_bh;
    const int t_lo = step == 0 ? 0 : step;   // first term of the sum
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // per (variant, hypothesis): maximum of the exponents, one warp each
    for (int i = warp; i < 2 * R; i += NT / 32) {
        const int var = i / R, h = i - var * R;
        const float* phi = (var ? rb_in : rsum_in) + static_cast<size_t>(row0 + h) * T;
        float mx = step == 0 ? 0.0f : -INFINITY;
        for (int t = max(t_lo, 1) + lane; t < T; t += 32) mx = fmaxf(mx, phi[t - 1]);
        mx = warp_max(mx);
        if (lane == 0) s_M[var][h] = mx;
    }
    if (tid < R) s_last[tid] = step == 0 ? a.bos : a.hist_tok[static_cast<size_t>(step - 1) * a.n_bh + row0 + tid];
    __syncthreads();
    for (int i = tid; i < 2 * R * Tp; i += NT) {
        const int vh = i / Tp, t = i - vh * Tp, var = vh / R, h = vh - var * R;
        float v = 0.0f;
        if (t < T && t >= t_lo) {
            const float M = s_M[var][h];
            v = t == 0 ? __expf(-M) : __expf((var ? rb_in : rsum_in)[static_cast<size_t>(row0 + h) * T + t - 1] - M);
        }
        smem[i] = v;
    }
    __syncthreads();
    const int tx = tid & (CTC_THREADS - 1), grp = tid / CTC_THREADS;
    const int c = blockIdx.x * CTC_THREADS + tx;
    const bool live = c < V;
    int base[R];
    float acc[R];
#pragma unroll
    for (int h = 0; h < R; ++h) { base[h] = ((c == s_last[h]) ? R * Tp : 0) + h * Tp; acc[h] = 0.0f; }
    if (live) {
        const float* xc = a.xlin + static_cast<size_t>(b) * T * V + c;
        // (the next chunk's posteriors are requested before the current ones are consumed)
        float xn[4];
        const int t_first = (t_lo & ~3) + 4 * grp;
#pragma unroll
        for (int u = 0; u < 4; ++u) xn[u] = (t_first + u < T) ? xc[static_cast<size_t>(t_first + u) * V] : 0.0f;
        for (int t = t_first; t < T; t += 4 * CTC_TSPLIT) {
            float xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                xv[u] = xn[u];
                const int tn = t + 4 * CTC_TSPLIT + u;
                xn[u] = (tn < T) ? xc[static_cast<size_t>(tn) * V] : 0.0f;
            }
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const float4 av = *reinterpret_cast<const float4*>(smem + base[h] + t);
                acc[h] = fmaf(av.x, xv[0], acc[h]); acc[h] = fmaf(av.y, xv[1], acc[h]);
                acc[h] = fmaf(av.z, xv[2], acc[h]); acc[h] = fmaf(av.w, xv[3], acc[h]);
            }
        }
    }
    if (grp > 0) {
#pragma unroll
        for (int h = 0; h < R; ++h) s_part[((grp - 1) * R + h) * CTC_THREADS + tx] = acc[h];
    }
    __syncthreads();
    if (grp > 0 || !live) return;
#pragma unroll
    for (int h = 0; h < R; ++h) {
#pragma unroll
        for (int g = 0; g < CTC_TSPLIT - 1; ++g) acc[h] += s_part[(g * R + h) * CTC_THREADS + tx];
        const int row = row0 + h;
        float psi;
        if (c == a.blank && a.eos != a.blank) {
            psi = CTC_NEG;
        } else if (c == a.eos) {
            psi = rsum_in[static_cast<size_t>(row) * T + a.enc_len[b] - 1];   // Alg.2-3
        } else if (acc[h] > 1e-30f) {
            psi = s_M[c == s_last[h] ? 1 : 0][h] + __logf(acc[h]);
        } else {   // out of the linear range: the same sum in the log domain (rare; keeps the result range-independent)
            const float* phi = ((c == s_last[h]) ? rb_in : rsum_in) + static_cast<size_t>(row) * T;
            const float* xg = a.x + static_cast<size_t>(b) * T * V + c;
            float pm, ps = 1.0f;
            int start;
            if (step == 0) { pm = xg[0]; start = 1; } else { pm = CTC_NEG; start = step; }
            for (int t = start; t < T; ++t) {
                const float term = phi[t - 1] + xg[static_cast<size_t>(t) * V];
                if (term > pm) { ps = ps * __expf(pm - term) + 1.0f; pm = term; }
                else ps += __expf(term - pm);
            }
            psi = pm + __logf(ps);
        }
        const float sc = a.weight * (psi - psi_prev[row]);
        float* o = a.out + static_cast<size_t>(row) * V + c;
        *o = a.accumulate ? *o + sc : sc;
    }
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
    const int step = a.step_ptr[row] + a.step_adj;   // the step whose survivors are being updated
    if (step < 0) return;                            // enqueued at the head of every search step: nothing to do before the first
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

int ctc_prefix_reset(float* x, float* xlin, float* xb, const int* enc_len, int B, int T, int V, int blank, int beam, float* rsum,
                     float* rb, float* psi_prev, cudaStream_t stream) {
    SBK_REQUIRE(T >= 1 && T * 5 * 4 <= 200 * 1024, "ctc scorer: T=%d out of range", T);
    ctc_logsoftmax_mask_kernel<<<B * T, 256, 0, stream>>>(x, xlin, xb, enc_len, T, V, blank);
    SBK_LAUNCH_CHECK();
    ctc_init_kernel<<<B * beam, 128, T * 4, stream>>>(xb, T, beam, rsum, rb, psi_prev);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

static CtcArgs make_args(const CtcStep& p, int step_adj) {
    CtcArgs a;
    a.x = p.x; a.xlin = p.xlin; a.xb = p.xb; a.rsum_base = p.rsum_base; a.rb_base = p.rb_base; a.psi_base = p.psi_base; a.enc_len = p.enc_len;
    a.hist_tok = p.hist_tok; a.hist_pred = p.hist_pred; a.step_ptr = p.step_ptr; a.step_adj = step_adj;
    a.n_bh = p.n_bh; a.bos = p.bos; a.T = p.T; a.V = p.V;
    a.beam = p.beam; a.blank = p.blank; a.eos = p.eos; a.weight = p.weight; a.out = p.out; a.accumulate = p.accumulate;
    return a;
}

template <int R>
static int launch_score(const CtcArgs& a, cudaStream_t stream) {
    static bool attr = false;
    if (!attr) {
        SBK_CUDA_CHECK(cudaFuncSetAttribute(ctc_score_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    const int Tp = (a.T + 3) & ~3;
    ctc_score_kernel<R><<<dim3(ceil_div(a.V, CTC_THREADS), a.n_bh / R), CTC_THREADS * CTC_TSPLIT,
                          ((size_t)2 * R * Tp + (size_t)(CTC_TSPLIT - 1) * R * CTC_THREADS) * 4, stream>>>(a);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

int ctc_prefix_score(const CtcStep& p, cudaStream_t stream) {
    const CtcArgs a = make_args(p, 0);   // runs before the beam kernel advances the step counters
    static bool attr = false;
    if (!attr) {
        SBK_CUDA_CHECK(cudaFuncSetAttribute(ctc_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    // hypotheses per CTA: the widest compiled group that divides the beam and whose exponent table fits shared memory
    const int Tp = (p.T + 3) & ~3;
    const int fit = (200 * 1024) / ((2 * Tp + (CTC_TSPLIT - 1) * CTC_THREADS) * 4);
    static const int widths[] = {16, 12, 11, 10, 8, 6, 5, 4, 3, 2, 1};
    int R = 1;
    for (int w : widths) if (p.beam % w == 0 && w <= fit) { R = w; break; }
    switch (R) {
        case 16: return launch_score<16>(a, stream);
        case 12: return launch_score<12>(a, stream);
        case 11: return launch_score<11>(a, stream);
        case 10: return launch_score<10>(a, stream);
        case 8: return launch_score<8>(a, stream);
        case 6: return launch_score<6>(a, stream);
        case 5: return launch_score<5>(a, stream);
        case 4: return launch_score<4>(a, stream);
        case 3: return launch_score<3>(a, stream);
        case 2: return launch_score<2>(a, stream);
        default: return launch_score<1>(a, stream);
    }
}

int ctc_prefix_update(const CtcStep& p, cudaStream_t stream) {
    const CtcArgs a = make_args(p, -1);  // runs after it (at the head of the next search step)
    ctc_update_kernel<<<p.n_bh, 128, 5 * p.T * 4, stream>>>(a);
    SBK_LAUNCH_CHECK();
    return SBK_OK;
}

}  // namespace sbk
