// Data model of the fused frame-step kernel: tile constants, the phase descriptor, device-side state and launch parameters.
// Part of the ar_engine.cu translation unit (include order: ar_program, ar_gemv, ar_attention, ar_sampler,
// the persistent kernel in ar_engine.cu, ar_prefill).
#pragma once

namespace {

constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int HD = 128;          // head_dim (required)
constexpr int MAXB = Q3_MAX_BATCH;
constexpr int MAXCOLS = 32;      // columns per pass (batch rows or prefill tokens)
constexpr int MAXSPLIT = 16;
constexpr int RMAX = 2;          // max GQA group size (q heads per kv head)
constexpr int PCOL = 20;         // padded row count of a partial column (bank-conflict-free)
constexpr int XS_COL_BYTES = 4096 + 64;     // one staged column: K=2048 bf16 (+64 B skew)
constexpr int XS_BYTES = 32 * XS_COL_BYTES;  // staged activations at NT=4: 32 cols
// Shared memory is sized per batch class (NT n8-tiles): a small request leaves most of the 228 KB as L1, which is
// what absorbs register spills / ABI stack traffic (with a 216 KB request every spill is an L2 round trip).
constexpr int ATT_SMEM = (2 * 2 * 128 + 32 * 2 * 130) * 4;    // attention: qs (<= 2 queries) + per-half-warp partials
constexpr int SAMPLER_SMEM = (2 * 4096 + 64 + 256) * 4;
__host__ __device__ constexpr int xs_bytes_nt(int nt) { return nt * 8 * XS_COL_BYTES; }
__host__ __device__ constexpr int part_bytes_nt(int nt) { return 16 * 2 * nt * 8 * 20 * 4; }
__host__ __device__ constexpr int smem_bytes_nt(int nt) {
  return (xs_bytes_nt(nt) + part_bytes_nt(nt) > ATT_SMEM ? xs_bytes_nt(nt) + part_bytes_nt(nt) : ATT_SMEM) + 1024;
}
constexpr int MAXV = 4096;       // max vocab handled by the sampler

enum PhaseType { PH_GEMV = 0, PH_ATTN = 1, PH_SAMPLE = 2 };
enum Epi { EPI_STORE = 0, EPI_BIAS = 1, EPI_RESID = 2, EPI_SWIGLU = 3, EPI_LOGITS = 4 };
enum NcMode { NC_B = 0, NC_2B = 1 };
enum SeqMode { SEQ_CP = 0, SEQ_DECODE = 1 };

struct Phase {
  int type, epi, ncmode, stack;
  // ---- GEMV
  const uint4* w;      // packed weights
  int n_tiles, kb;     // rows/16, K/32
  int tq, tr;          // n_tiles = tq*grid + tr: CTA c owns tq (+1 if c < tr) consecutive tiles
  const bf16* src;     // [nc][src_ld]
  int src_ld;
  const bf16* norm_w;  // RMSNorm weight applied while staging (nullable)
  float eps;
  void* dst;
  int dst_ld;
  const bf16* bias;
  bf16* save_normed;   // optional copy of the normed input (past_hidden), ld = K
  // ---- ATTN
  int layer, seqmode, nq, ctx_end;
  const bf16* qn;
  const bf16* kn;
  // ---- SAMPLE
  int group;           // 0 = talker codebook-0; j>=1 = code predictor codebook j
  int pad_;
};

struct StackDev {
  int hidden, layers, nh, nkv, inter, vocab;
  float eps;
  bf16 *h, *qkv, *attn, *act;  // activations [cols][...]
  bf16 *kc, *vc;               // KV cache [seq][layer][nkv][cap][128]
  int cap;
  const bf16 *rope_cos, *rope_sin;  // [cap][64]
  float* logits;               // [MAXB][vocab]
};

struct DevState {
  unsigned int bar_count;
  int error;
  int B;
  int step;            // frames whose 16 codes are complete
  int len0[MAXB];
  int finished[MAXB];
  int n_valid[MAXB];
  int n_gen[MAXB];
  int c0[MAXB];
  int trailing_len[MAXB];
  int cur[MAXB][Q3_NUM_GROUPS_MAX];
  unsigned int split_cnt[MAXB * 16];
};

struct KParams {
  const Phase* prog;
  int n_phases;
  int mode;        // 0 = one pass over the program (prefill chunk / prefill head), 1 = frame loop
  int max_iters;
  DevState* st;
  StackDev talker, cp;
  int G, eos, has_proj;
  int B;                    // sequences in this request (constant per launch)
  int len0[MAXB];           // prompt lengths
  int trailing_len[MAXB];
  q3_sampling sp;
  // sampler / embed resources
  const bf16* emb_t;        // talker codec_embedding [V][H]
  const bf16* emb_cp;       // cp codec_embedding [G-1][Vc][H]
  bf16* x_cp;               // CP input [2][B][H]
  const bf16* cp_next;      // rows fed to passes >= 1: projected embedding table [(G-1)*Vc][Hc] or emb_cp itself
  bf16* cp_next_dst;        // where they go: cp.h (table / Identity projection) or x_cp (projection phase follows)
  int cp_next_w;            // row width of cp_next / cp_next_dst
  bf16* past_hidden;        // [B][H]
  const bf16* trailing;     // [B][stride][H]
  int trailing_stride;
  const bf16* tts_pad;      // [H]
  unsigned char* seen;      // [B][V]
  int* codes_out;           // [B][codes_stride][G]
  int codes_stride;
  float* split_buf;         // [MAXB*nkv*MAXSPLIT][RMAX][130]
  // debug hooks
  const int* forced;
  int n_forced;
  float* dbg_tlogits;
  float* dbg_clogits;
  int dbg_skip;             // ablation bits (tools/ablate_phase.py): 1 stage, 2 main loop, 4 epilogue, 8 preload, 16 whole body
  unsigned long long* prof;  // [n_phases][grid][8] globaltimer ns: [0] phase end, [1] barrier passed, [2..5] inner marks, [6] start
};

}  // namespace
