// Internal (C++) interfaces between the kernels' host launchers and the C-ABI layer.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace sbk {

enum GemmEpiMode { EPI_F16 = 0, EPI_F32 = 1, EPI_RESID = 2, EPI_GLU = 3, EPI_ROPE = 4, EPI_QKV_CACHE = 5 };
enum GemmAct { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_RELU = 3, ACT_SILU_FAST = 4 /* tanh.approx form; for EPI_GLU: fast gate sigmoid */ };

struct GemmEpilogue {
    int mode = EPI_F16;
    int act = ACT_NONE;
    const float* bias = nullptr;   // [N] or null
    void* out = nullptr;           // fp16 or fp32, row stride ldo (elements)
    int ldo = 0;
    const float* resid = nullptr;  // EPI_RESID: fp32 [M, ldo]
    float alpha = 1.0f;            // EPI_RESID scale; EPI_ROPE: scale applied to q
    const int* row_lens = nullptr; // EPI_RESID: rows (b, t >= row_lens[b]) get alpha = 0
    int T = 1;                     // frames per utterance (row = b*T + t)
    const float* rope_cos = nullptr;  // EPI_ROPE: [T, head_dim/2]
    const float* rope_sin = nullptr;
    int head_dim = 64;
    // EPI_QKV_CACHE (decoder self-attention in_proj, columns [q | k | v] of width qkv_d): q -> out (fp16, ldo), k / v ->
    // cache[(row * S_max + step_ptr[row]) * qkv_d + col]
    __half* kcache = nullptr; __half* vcache = nullptr; const int* step_ptr = nullptr; int S_max = 0; int qkv_d = 0;
    // EPI_F16 (2-CTA kernel) scatter of the cross-attention [K | V] projection (N = 2 * kv_heads * 64, row = utt * T + t) to
    // part[K|V][utt][head][t][64]; kv_part_stride = elements between the K part and the V part.  0 = plain row-major store.
    int kv_heads = 0; size_t kv_part_stride = 0;
};

// out = epilogue(A[M,K] fp16 x W[N,K]^T fp16), tcgen05 tensor cores. gemm_tc.cu
int gemm_f16(const void* A, int lda, const void* W, int ldw, const GemmEpilogue& epi, int M, int N, int K,
             cudaStream_t stream);
// Small-M variant for the decode steps of several batches (M = live hypotheses, 64..512): 128 x 32/64 tiles, deep TMA
// ring; any epilogue mode incl. EPI_QKV_CACHE. gemm_tc.cu
int gemm_f16_small(const void* A, int lda, const void* W, int ldw, const GemmEpilogue& epi, int M, int N, int K,
                   cudaStream_t stream);
// 2-CTA (cta_group::2) persistent variant, N % 256 == 0. gemm_tc2.cu
int gemm_f16_2cta(const void* A, int lda, const void* W, int ldw, const GemmEpilogue& epi, int M, int N, int K,
                  cudaStream_t stream);


const char* last_error();
void launch_count_begin_capture();
long long launch_count_end_capture();
void launch_count_add(long long n);
long long launch_count();

// Optional live timing of every tcgen05 GEMM launch (CUDA events on the launching stream).
struct GemmProfile {
    bool enabled = false;
    std::vector<cudaEvent_t> ev;   // start/stop pairs
    std::vector<double> flops;     // 2*M*N*K per launch
    std::vector<int> shape;        // M, N, K, epilogue mode per launch
};
GemmProfile* gemm_profile();

// ---- fbank.cu
struct Fbank;
int fbank_create(Fbank** out, int n_fft, int hop, int n_mels, const float* window_host, const float* mel_matrix_host,
                 float amin, float top_db);
void fbank_destroy(Fbank* fb);
int fbank_num_frames(const Fbank* fb, int L);
int fbank_forward(const Fbank* fb, const float* wav, int B, int L, float* out, int* utt_max, const float* mean,
                  const float* stdv, float eps, cudaStream_t stream);
int global_norm_forward(const float* x, float* out, int B, int T, int F, const float* mean, const float* stdv,
                        float eps, cudaStream_t stream);
int sentence_norm_forward(const float* x, float* out, const float* rel_len, int B, int T, int F, int std_norm,
                          int avoid_padding_norm, float eps, cudaStream_t stream);

// ---- frontend.cu
int cnn_frontend_forward(const float* feats, int B, int T0, int F0, const float* w1, const float* b1, const float* g1,
                         const float* be1, int C1, const __half* w2p, const float* b2, const float* g2,
                         const float* be2, int C2, __half* act1_h, float* act1_f, __half* out_h, float* out_f,
                         cudaStream_t stream);

// ---- encoder_ops.cu
int layernorm_rows(const float* x, void* out, bool out_half, const float* gamma, const float* beta, int M, int D,
                   float eps, bool act_silu, cudaStream_t stream);
// y = LN_a(x) (fp32, stored when y_out != null), z = LN_b(y) -> z_out (fp16 or fp32): two chained LayerNorms in one pass
int layernorm2_rows(const float* x, float* y_out, void* z_out, bool z_half, const float* ga, const float* ba, float eps_a,
                    const float* gb, const float* bb, float eps_b, int M, int D, cudaStream_t stream);
int cast_f32_f16(const float* in, __half* out, size_t n, cudaStream_t stream);
// chunk > 0: Dynamic Chunk Convolution (inputs past the end of the output frame's chunk are zero)
int dwconv_ln_swish(const float* glu, int B, int T, int D, int K, const float* wdw, const float* bdw,
                    const float* gamma, const float* beta, float eps, __half* out, cudaStream_t stream, int chunk = 0);
int encoder_attention(const __half* qkv, int ld, int B, int T, int H, int head_dim, const int* lens, bool relpos,
                      const float* pos_u, const float* pos_v, const __half* P, int ldp, float scale, __half* out,
                      int ldo, cudaStream_t stream, int chunk = 0, int left_chunks = -1);

// ---- decoder.cu
enum SkinnyEpi { SK_F16 = 0, SK_F16_GELU = 1, SK_F32 = 2, SK_RESID = 3, SK_QKV_CACHE = 4, SK_F16_RELU = 5 };
struct SkinnyArgs {
    const __half* A; int lda;
    const __half* W; int ldw;
    const float* bias;
    int n_rows, N, K, epi;
    void* out; int ldo;              // SK_F16/F32: out ; SK_RESID: fp32 x (in place) ; SK_QKV_CACHE: q buffer fp16 [n, d]
    __half* kcache; __half* vcache;  // SK_QKV_CACHE: [n_rows, S_max, d]
    const int* step_ptr; int S_max; int d; float q_scale;
    // LayerNorm-fused variant: A = LayerNorm(X fp32 [n_rows, K]) computed in-kernel (X != nullptr)
    const float* X; const float* ln_g; const float* ln_b; float ln_eps;
};
void set_pdl(bool on);
int skinny_gemm(const SkinnyArgs& a, cudaStream_t stream);
struct DecAttnArgs {
    const __half* q; int ldq;
    const __half* kbase; const __half* vbase;
    size_t row_stride;
    int key_stride;
    int head_stride = 0;  // elements between heads inside a row block; 0 = dh (heads side by side in one key row)
    int rows_per_block;
    const int* n_keys_ptr;
    const int* enc_len;
    int n_keys_fixed;
    int H, dh;
    __half* out; int ldo;
    const int* lineage; int lin_stride;  // beam search: [2][n_rows][lin_stride] cache-row table (null for greedy)
    const int* tok_cache; int pad_tok;   // LM pad mask: keys whose token (tok_cache[phys_row][pos]) == pad_tok are masked
};
int dec_attention(const DecAttnArgs& a, int n_rows, int max_keys, cudaStream_t stream);
int make_tmap_kv_f16(CUtensorMap* out, const void* base, int n_utt, int T, int H, uint64_t key_stride_elems,
                     uint64_t utt_stride_elems, int box_T);
struct BeamLm {  // TransformerLM scorer state the beam step feeds (all null/0 when there is no LM)
    const float* emb = nullptr; const float* pe = nullptr; int d = 0;
    float* x = nullptr; __half* x16 = nullptr; int* tok_cache = nullptr;
};
struct BeamStepArgs {
    const float* logits; int V; int beam; int S_max;
    float* seq_scores; int* lineage; int* step_arr; int* finished; int* n_full;
    int* hist_tok; int* hist_pred; float* hist_score; float* hist_lp;
    float temperature, eos_threshold, minus_inf;
    int min_steps, eos, use_eos_threshold, length_norm;
    const float* emb; const float* pe; int d; float* x_next;
    const float* add_scores;  // [n_bh, V] pre-weighted scorer scores or null
    BeamLm lm;
    float attn_weight = 1.0f;  // 1 - ctc_weight (seq2seq.py:803-804, _attn_weight_step)
    int blank = -1;            // CTC blank index, blocked in the log-probs (scorer.py:1248-1250); -1 = no CTC scorer
    float add_const = 0.0f;    // LengthScorer: weight * 1 added to every token (scorer.py:1043-1071)
    const float* add_row = nullptr;  // CoverageScorer: [n_bh] weighted score added to every token of a hypothesis
    float* scratch = nullptr;        // [n_bh * 33] floats: per-row candidates between the two kernels of a step (beam <= 16)
};
// CoverageScorer (decoders/scorer.py:788-955) on the last decoder layer's head-averaged cross-attention
struct CoverageStep {
    const __half* q; int ldq; const __half* kbase; size_t utt_stride; int key_stride; const int* enc_len;
    int head_stride = 0;  // 0 = 64 (heads side by side in a key row)
    int rows_per_utt, T, H; float* cov_base; const int* hist_pred; const int* step_ptr; int n_bh;
    float threshold, weight; float* out;
};
int coverage_score(const CoverageStep& p, cudaStream_t stream);
// CTC prefix scorer (ctc_scorer.cu)
struct CtcStep {
    const float* x; const float* xlin; const float* xb; const int* enc_len;   // xlin = exp(x)
    float* rsum_base; float* rb_base; float* psi_base;  // [2][n_bh, T], [2][n_bh, T], [2][n_bh]: ping-pong by step parity
    float* tab; float* tabM;                            // score-kernel operand tables [n_bh * 2 * (T + 3)], [n_bh * 2]
    const int* hist_tok; const int* hist_pred;
    const int* step_ptr;                                // device step counters [n_bh] (same value in every row)
    int n_bh, bos, T, V, beam, blank, eos;
    float weight; float* out; int accumulate;
};
int ctc_prefix_reset(float* x, float* xlin, float* xb, const int* enc_len, int B, int T, int V, int blank, int beam, float* rsum,
                     float* rb, float* psi_prev, float* tab, float* tabM, cudaStream_t stream);
int ctc_prefix_score(const CtcStep& p, cudaStream_t stream);
// x [rows, V] fp32: optional in-place log_softmax per row, arg-max per row -> idx (may be null)
int rows_logsoftmax_argmax(float* x, int rows, int V, bool do_logsoftmax, int* idx, cudaStream_t stream);
int ctc_prefix_update(const CtcStep& p, cudaStream_t stream);
int beam_reset(int n_bh, int beam, int S_max, int bos, int* step_arr, float* seq_scores, int* lineage, int* finished,
               int* n_full, const float* emb, const float* pe, int d, float* x, const BeamLm* lm, cudaStream_t stream);
int layernorm_dual(float* x, __half* x16, const float* gamma, const float* beta, int M, int D, float eps, bool write_f32,
                   cudaStream_t stream);
int weighted_log_softmax(const float* logits, float* out, int rows, int V, float temperature, float weight,
                         cudaStream_t stream);
int beam_step(const BeamStepArgs& p, int B, cudaStream_t stream);
int greedy_reset(int* tokens, int tok_stride, int n_rows, int bos, int* step_arr, int* has_ended, int* ended_count,
                 const float* emb, const float* pe, int d, float* x, cudaStream_t stream);
int greedy_select(const float* logits, int n_rows, int V, int* step_arr, int eos, int* tokens, int tok_stride,
                  int* has_ended, int* ended_count, int* pred, float* score, int out_stride, float* log_probs, int L,
                  const float* emb, const float* pe, int d, float* x_next, cudaStream_t stream);

}  // namespace sbk
