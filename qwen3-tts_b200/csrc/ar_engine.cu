// AR hot path of Qwen3-TTS on B200 (sm_100a): talker + 15-pass code predictor + logits processing + sampling
// + next-embed as ONE persistent cooperative kernel per group of frame-steps.
//
// Replaces (reference, paths relative to /root/reference/qwen_tts/core/models/modeling_qwen3_tts.py):
//   decode step  :1669-1744   code predictor :1250-1312 / :1671-1680   layers :1393-1424, :985-1012
//   attention    :761-805, :916-958, :634-657   RMSNorm :605-610   RoPE :660-724, :858-882   MLP :853-855
//   sampling     HF processors configured at :2044-2066 (semantics restated in oracle/sampler.py)
//
// Design:
//   * the step is a "phase program": GEMV / attention / sample phases separated by a software grid barrier;
//     one CTA per SM, all CTAs walk the same program (cooperative launch guarantees co-residency).
//   * GEMV phases stream weights exactly once per step, straight from HBM/L2 into mma.sync A-fragments
//     (weights are re-packed at load into 8x32 bf16 sub-tiles = 512 B contiguous per warp load; a K
//     permutation shared by A and B lets every lane use one 16-byte vector load per 8 k-values), batch
//     columns ride in the N dimension (n8 tiles), fp32 accumulate, cross-warp split-K reduced in smem.
//   * the next phase's weight slice is pulled into L2 with cp.async.bulk.prefetch.L2 before a CTA waits on
//     the barrier, so HBM keeps streaming across the dependency bubbles.
//   * rounding points mirror the PyTorch bf16 path (linear outputs, RMSNorm, RoPE, residual adds are
//     rounded to bf16 exactly where the reference rounds).
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "../../include/qwen3tts_b200.h"

#include <cooperative_groups.h>
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <vector>

thread_local std::string g_q3_err;
int q3_set_err(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_q3_err = buf;
  return 1;
}

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

// ------------------------------------------------------------------------------------------------
// grid barrier (monotonic counter; arrive = release, wait = acquire)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(DevState* st, unsigned int& epoch) {
  __syncthreads();  // every thread's global writes happen-before thread 0's release (bar.sync is cumulative)
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&st->bar_count) : "memory");
    long long t0 = clock64();
    unsigned int v;
    while (true) {
      // RELAXED poll on purpose: ld.acquire.gpu compiles to LDG.STRONG + CCTL.IVALL, i.e. it invalidates the whole
      // L1 on every poll iteration (measured: ~55 invalidations per barrier), which evicts the stack / spill lines
      // of all 16 warps and makes every phase start cold.  Correctness does not need the invalidation: every
      // cross-CTA read in this kernel is an L2 load (ld.global.cg), the writers released at gpu scope before
      // their arrival became visible, and the GPU does not speculate loads past this loop + bar.sync.
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&st->bar_count) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if (clock64() - t0 > 8000000000LL) {  // ~4 s: never hang the box
        st->error = 77;
        __threadfence();
        __trap();
      }
    }
  }
  __syncthreads();
}

// fine-grained profiling marks (thread 0 of CTA 0 only, first frame of a profiled launch)
__shared__ unsigned long long* g_prof_row;
#define PROF_MARK(k)                                                                   \
  do {                                                                                 \
    if (threadIdx.x == 0 && g_prof_row) {                                              \
      unsigned long long _t;                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                          \
      g_prof_row[k] = _t;                                                              \
    }                                                                                  \
  } while (0)

// NOTE ON CODE SIZE: the frame program walks ~560 phases per frame-step, alternating between the three phase
// bodies below.  Their combined hot code must stay inside the SM's ~32 KB instruction cache, otherwise every
// phase re-fetches its instructions from L2 (measured: ~4 us of pure fetch stall per phase with 150 KB of code).
// Hence: loops are rolled (#pragma unroll 1) wherever latency is not at stake, bulk data goes through shared
// memory instead of unrolled register arrays, and there is no 64-bit division on the device.

// ------------------------------------------------------------------------------------------------
// block-wide helpers (NTHREADS threads)
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ float block_reduce(float v, float* red, int op /*0 max, 1 sum*/) {
  if (op == 0) v = warp_max(v); else v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[threadIdx.x & (NWARPS - 1)];
#pragma unroll
  for (int o = NWARPS / 2; o > 0; o >>= 1) {
    const float n = __shfl_xor_sync(0xffffffffu, r, o);
    r = op == 0 ? fmaxf(r, n) : r + n;
  }
  return r;
}
__device__ __noinline__ int block_min_int(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = red[threadIdx.x & (NWARPS - 1)];
#pragma unroll
  for (int o = NWARPS / 2; o > 0; o >>= 1) r = min(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}

// ------------------------------------------------------------------------------------------------
// GEMV phase:  dst[col][row] = epi( sum_k W[row][k] * x[col][k] )
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int xs_stride_bytes(int K) { return ((K * 2 + 127) / 128) * 128 + 64; }

__device__ __forceinline__ int phase_nc(int ncmode, const KParams& P) {
  const int B = P.B;
  return ncmode == NC_B ? B : 2 * B;
}

// balanced contiguous split of a phase's row tiles over the CTAs (tq/tr precomputed on the host)
__device__ __forceinline__ void cta_tiles(int tq, int tr, int& t0, int& ntc) {
  const int c = blockIdx.x;
  t0 = c * tq + min(c, tr);
  ntc = tq + (c < tr ? 1 : 0);
}

__device__ __forceinline__ void prefetch_phase_weights(const Phase& ph) {
  if (ph.type != PH_GEMV) return;
  int t0, ntc;
  cta_tiles(ph.tq, ph.tr, t0, ntc);
  const unsigned int bytes = (unsigned int)ntc * (unsigned int)ph.kb * 1024u;
  const unsigned int off = threadIdx.x * 32768u;
  if (off < bytes) {
    const char* base = reinterpret_cast<const char*>(ph.w) + (size_t)t0 * ph.kb * 1024;
    l2_prefetch_bulk(base + off, min(32768u, bytes - off));
  }
}

__device__ __forceinline__ uint32_t norm_pair(uint32_t x2, uint32_t w2, float inv) {
  __nv_bfloat162 t = __floats2bfloat162_rn(bf16lo(x2) * inv, bf16hi(x2) * inv);
  __nv_bfloat162 r = __hmul2(t, *reinterpret_cast<const __nv_bfloat162*>(&w2));
  return *reinterpret_cast<uint32_t*>(&r);
}

// stage x (optionally RMS-normed) into smem as bf16 [col][K] (rows skewed by 64 B).  One warp per column; a lane
// issues up to 8 independent 16-byte loads (K <= 2048 per pass) before touching the data; the RMSNorm runs on the
// registers (sum of squares -> warp reduce -> scale) and the result is written to smem once.  The norm weights
// were prefetched into nw_s one phase ahead (see the main loop), so they cost no global round trip here.
__device__ __forceinline__ void stage_columns(const bf16* __restrict__ src, int src_ld, bool normed, float eps,
                                              bf16* __restrict__ save, int K, int nc, char* __restrict__ xs, int xstride,
                                              const uint4* __restrict__ nw_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = K >> 3;
#pragma unroll 1
  for (int col = warp; col < nc; col += NWARPS) {
    const uint4* xr = reinterpret_cast<const uint4*>(src + (size_t)col * src_ld);
    uint4* drow = reinterpret_cast<uint4*>(xs + (size_t)col * xstride);
#pragma unroll 1
    for (int vb = 0; vb < nv; vb += 256) {  // warp-uniform trip count (warp_sum below); one pass when K <= 2048
      const int v0 = vb + lane;
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (v0 + 32 * i < nv) v[i] = ldcg16(xr + v0 + 32 * i);
      if (normed) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (v0 + 32 * i < nv) {
            float f;
            f = bf16lo(v[i].x); ss += f * f; f = bf16hi(v[i].x); ss += f * f;
            f = bf16lo(v[i].y); ss += f * f; f = bf16hi(v[i].y); ss += f * f;
            f = bf16lo(v[i].z); ss += f * f; f = bf16hi(v[i].z); ss += f * f;
            f = bf16lo(v[i].w); ss += f * f; f = bf16hi(v[i].w); ss += f * f;
          }
        }
        ss = warp_sum(ss);
        PROF_MARK(7);
        const float inv = rsqrtf(ss / (float)K + eps);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (v0 + 32 * i < nv) {
            // bf16(bf16(x*inv) * w): fp32 scale, one packed RN conversion, then a packed bf16 multiply (HMUL2.BF16
            // rounds the exact product to nearest-even = the reference's bf16 x bf16 -> bf16 multiply)
            const uint4 w = nw_s[v0 + 32 * i];
            uint4 o;
            o.x = norm_pair(v[i].x, w.x, inv);
            o.y = norm_pair(v[i].y, w.y, inv);
            o.z = norm_pair(v[i].z, w.z, inv);
            o.w = norm_pair(v[i].w, w.w, inv);
            v[i] = o;
            if (save) reinterpret_cast<uint4*>(save + (size_t)col * K)[v0 + 32 * i] = o;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (v0 + 32 * i < nv) drow[v0 + 32 * i] = v[i];
    }
  }
}

constexpr int DEPTH = (NTHREADS <= 256) ? 8 : 4;  // k32-blocks (2 x 16 B per lane each) kept in flight per warp

__device__ __forceinline__ void gemv_preload(uint4 (&a)[DEPTH][2], const uint4* __restrict__ wp, int nk) {
#pragma unroll
  for (int i = 0; i < DEPTH; ++i)
    if (i < nk) { a[i][0] = ldg_stream(wp + i * 64); a[i][1] = ldg_stream(wp + i * 64 + 32); }
}

template <int NT, bool STAGED>
__device__ __forceinline__ void gemv_segment(float (&acc)[NT][4], uint4 (&a)[DEPTH][2], const uint4* __restrict__ wp, int nk, int kb0,
                                             const char* __restrict__ xs, int xstride, const bf16* __restrict__ src,
                                             int src_ld, int nc, int g, int t) {
  // rolling register pipeline: DEPTH k32-blocks (2 x 16 B per lane each) always in flight; a slot is refilled the
  // moment it has been copied out, so no fragment is ever held twice
#pragma unroll 1
  for (int k0 = 0; k0 < nk; k0 += DEPTH) {
#pragma unroll
    for (int i = 0; i < DEPTH; ++i) {
      if (k0 + i < nk) {
        const int kb = kb0 + k0 + i;
        const uint4 r = a[i][0], s = a[i][1];
        if (k0 + i + DEPTH < nk) {
          a[i][0] = ldg_stream(wp + (k0 + i + DEPTH) * 64);
          a[i][1] = ldg_stream(wp + (k0 + i + DEPTH) * 64 + 32);
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          uint4 b;
          const int col = n * 8 + g;
          if (STAGED) {
            b = *reinterpret_cast<const uint4*>(xs + (size_t)col * xstride + kb * 64 + t * 16);
          } else {
            b = (col < nc) ? ldcg16(src + (size_t)col * src_ld + kb * 32 + t * 8) : make_uint4(0, 0, 0, 0);
          }
          mma_bf16_16816(acc[n], r.x, s.x, r.y, s.y, b.x, b.y);
          mma_bf16_16816(acc[n], r.z, s.z, r.w, s.w, b.z, b.w);
        }
      }
    }
  }
}

template <int NT>
__device__ __noinline__ void gemv_phase(const Phase& ph, const KParams& P, unsigned char* smem, const uint4* nw_s) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // descriptor fields -> registers once
  const uint4* const wbase = ph.w;
  const int KB = ph.kb, K = KB * 32, epi = ph.epi;
  const bf16* const src = ph.src;
  const int src_ld = ph.src_ld, dst_ld = ph.dst_ld;
  const bf16* const norm_w = ph.norm_w;
  void* const dst = ph.dst;
  const bf16* const bias = ph.bias;
  const int nc = phase_nc(ph.ncmode, P);
  int t0, ntc;
  cta_tiles(ph.tq, ph.tr, t0, ntc);
  const int skip = P.dbg_skip;
  if (skip & 16) return;

  char* xs = reinterpret_cast<char*>(smem);
  float* part = reinterpret_cast<float*>(smem + xs_bytes_nt(NT));  // [NWARPS][2][NT*8][PCOL]
  const int xstride = xs_stride_bytes(K);
  const bool staged = (norm_w != nullptr) || (xstride * (NT * 8) <= xs_bytes_nt(NT));
  bf16* const save = (ph.save_normed != nullptr && blockIdx.x == 0) ? ph.save_normed : nullptr;
  // first weight fragments of this warp go in flight BEFORE the activations are staged (they do not depend on x)
  uint4 afr[DEPTH][2];
  const int TB0 = min(NWARPS, ntc);
  const int upw0 = (TB0 * KB + NWARPS - 1) / NWARPS;
  const bool have0 = ntc > 0 && warp * upw0 < TB0 * KB && !(skip & (8 | 2));
  if (have0) {
    const int u = warp * upw0, tl = u / KB, kb0 = u - tl * KB;
    gemv_preload(afr, wbase + ((size_t)(t0 + tl) * KB + kb0) * 64 + lane, min(KB - kb0, min(TB0 * KB, u + upw0) - u));
  }
  PROF_MARK(2);
  if (staged && (ntc > 0 || save) && !(skip & 1)) stage_columns(src, src_ld, norm_w != nullptr, ph.eps, save, K, nc, xs, xstride, nw_s);
  __syncthreads();
  PROF_MARK(3);
  if (ntc <= 0) return;

#pragma unroll 1
  for (int tb0 = 0; tb0 < ntc; tb0 += NWARPS) {
    const int TB = min(NWARPS, ntc - tb0);
    const int units = TB * KB;
    const int upw = (units + NWARPS - 1) / NWARPS;
    const int u1 = min(units, (warp + 1) * upw);
    int seg = 0;
#pragma unroll 1
    for (int u = warp * upw; u < u1;) {
      const int tl = u / KB;
      const int kb0 = u - tl * KB;
      const int nk = min(KB - kb0, u1 - u);
      float acc[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      const uint4* wp = wbase + ((size_t)(t0 + tb0 + tl) * KB + kb0) * 64 + lane;
      if (!(skip & 2)) {
      if (!(tb0 == 0 && seg == 0) || (skip & 8)) gemv_preload(afr, wp, nk);  // the very first segment was preloaded above
      if (staged) gemv_segment<NT, true>(acc, afr, wp, nk, kb0, xs, xstride, src, src_ld, nc, g, t);
      else gemv_segment<NT, false>(acc, afr, wp, nk, kb0, xs, xstride, src, src_ld, nc, g, t);
      }
      // spill partial sums: part[warp][seg][col][row]
      float* pp = part + ((warp * 2 + seg) * (NT * 8)) * PCOL;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int col = n * 8 + 2 * t;
        pp[(col)*PCOL + g] = acc[n][0];
        pp[(col + 1) * PCOL + g] = acc[n][1];
        pp[(col)*PCOL + g + 8] = acc[n][2];
        pp[(col + 1) * PCOL + g + 8] = acc[n][3];
      }
      ++seg;
      u += nk;
    }
    __syncthreads();
    PROF_MARK(4);
    // ---- cross-warp reduce + epilogue, one element per thread-iteration (independent global round trips)
    const bool swiglu = epi == EPI_SWIGLU;
    const int rsh = swiglu ? 3 : 4;  // rows per tile: 8 (gate/up pairs) or 16
    const int nelem = (skip & 4) ? 0 : (TB << rsh) * nc;
#pragma unroll 1
    for (int e = tid; e < nelem; e += NTHREADS) {
      const int r = e & ((1 << rsh) - 1);
      const int q = e >> rsh;
      const int col = q / TB, tl = q - col * TB;
      const int wf = (tl * KB) / upw, wl = ((tl + 1) * KB - 1) / upw;
      const int tile = t0 + tb0 + tl;
      const int row = tile * 16 + r;
      float resid = 0.f;
      if (epi == EPI_RESID) resid = bf2f(ldcg_bf16(reinterpret_cast<bf16*>(dst) + (size_t)col * dst_ld + row));  // in flight
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
      for (int w = wf; w <= wl; ++w) {
        const int sg = tl - (w * upw) / KB;
        const float* pp = part + ((w * 2 + sg) * (NT * 8) + col) * PCOL;
        s0 += pp[r];
        if (swiglu) s1 += pp[r + 8];
      }
      if (swiglu) {
        // rows 0-7 = gate, 8-15 = up of the same 8 intermediate channels (:853-855, bf16 rounding points)
        const float gt = rbf(s0), up = rbf(s1);
        const float sl = rbf(gt / (1.f + __expf(-gt)));
        reinterpret_cast<bf16*>(dst)[(size_t)col * dst_ld + tile * 8 + r] = f2bf(sl * up);
      } else if (epi == EPI_LOGITS) {  // bf16 linear output, then .float() (HF _sample)
        reinterpret_cast<float*>(dst)[(size_t)col * dst_ld + row] = rbf(s0);
      } else {
        if (epi == EPI_BIAS) s0 += bf2f(bias[row]);
        else if (epi == EPI_RESID) s0 = resid + rbf(s0);
        reinterpret_cast<bf16*>(dst)[(size_t)col * dst_ld + row] = f2bf(s0);
      }
    }
    __syncthreads();
    PROF_MARK(5);
  }
}

// ------------------------------------------------------------------------------------------------
// attention phase (per (sequence, kv head, split) unit): q/k RMSNorm + RoPE, KV append, single-query GQA
// ------------------------------------------------------------------------------------------------
// one warp normalises + rotates one 128-vector; lane owns dims {l, l+32, l+64, l+96}
__device__ __noinline__ void norm_rope_vec(const bf16* src, const bf16* nw, float eps, const bf16* cosr, const bf16* sinr,
                                           float* out_f32, bf16* out_bf16) {
  const int lane = threadIdx.x & 31;
  float x[4], w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = bf2f(ldcg_bf16(src + lane + 32 * i));
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = bf2f(nw[lane + 32 * i]);
  // rotate_half pairs: (l, l+64) and (l+32, l+96); cos/sin tables are [64] (emb = cat(freqs, freqs))
  const float c0 = bf2f(cosr[lane]), s0 = bf2f(sinr[lane]);
  const float c1 = bf2f(cosr[lane + 32]), s1 = bf2f(sinr[lane + 32]);
  float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)HD + eps);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = rbf(rbf(x[i] * inv) * w[i]);
  float o[4];
  o[0] = rbf(rbf(x[0] * c0) + rbf(-x[2] * s0));
  o[2] = rbf(rbf(x[2] * c0) + rbf(x[0] * s0));
  o[1] = rbf(rbf(x[1] * c1) + rbf(-x[3] * s1));
  o[3] = rbf(rbf(x[3] * c1) + rbf(x[1] * s1));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (out_f32) out_f32[lane + 32 * i] = o[i];
    else out_bf16[lane + 32 * i] = f2bf(o[i]);
  }
}

__device__ __noinline__ void attn_phase(const Phase& ph, const KParams& P, unsigned char* smem, int frame) {
  const StackDev& S = ph.stack == 0 ? P.talker : P.cp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nh = S.nh, nkv = S.nkv, layers = S.layers, cap = S.cap;
  const int R = nh / nkv;  // <= RMAX
  const int B = P.B;
  const int qkv_ld = (nh + 2 * nkv) * HD;
  const int seqmode = ph.seqmode, layer = ph.layer;
  const int nseq = B;
  const float eps = S.eps;
  const bf16 *qn = ph.qn, *kn = ph.kn;

  int nsplit = 1;
  if (seqmode == SEQ_DECODE) {
    int cmax = 0;
    for (int b = 0; b < B; ++b) cmax = max(cmax, P.len0[b]);
    cmax += frame + 1;
    const int byctx = (cmax + 127) >> 7;
    const int bygrid = (int)gridDim.x / (B * nkv);
    nsplit = max(1, min(min(byctx, bygrid), MAXSPLIT));
  }
  const int units = nseq * nkv * nsplit;

  float* qs = reinterpret_cast<float*>(smem);                 // [nq<=2][RMAX][128]
  float* red = qs + 2 * RMAX * HD;                            // [32 halfwarps][RMAX][130]
  __shared__ int s_ticket;

#pragma unroll 1
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int sp = unit % nsplit;
    const int kvh = (unit / nsplit) % nkv;
    const int si = unit / (nsplit * nkv);
    int seq, q0, nq, qstride, ctx_end;
    if (seqmode == SEQ_CP) { seq = si; q0 = si; nq = ph.nq; qstride = B; ctx_end = ph.ctx_end; }
    else { seq = si; q0 = si; nq = 1; qstride = 0; ctx_end = P.len0[si] + frame + 1; }
    const int SL = (ctx_end + nsplit - 1) / nsplit;
    const int s0 = sp * SL, s1 = min(ctx_end, s0 + SL);
    bf16* kc = S.kc + (((size_t)seq * layers + layer) * nkv + kvh) * (size_t)cap * HD;
    bf16* vc = S.vc + (((size_t)seq * layers + layer) * nkv + kvh) * (size_t)cap * HD;

    // ---- per query token: q heads -> smem (fp32), k (norm+rope) and v -> cache (owner split only)
    const int nvec = nq * (R + 2);
#pragma unroll 1
    for (int v = warp; v < nvec; v += NWARPS) {
      const int j = v / (R + 2), which = v - j * (R + 2);
      const int col = q0 + j * qstride;
      const int pos = ctx_end - nq + j;
      const bool owner = (pos >= s0 && pos < s1);
      const bf16* base = S.qkv + (size_t)col * qkv_ld;
      const bf16* cosr = S.rope_cos + (size_t)pos * 64;
      const bf16* sinr = S.rope_sin + (size_t)pos * 64;
      if (which < R) {
        norm_rope_vec(base + (kvh * R + which) * HD, qn, eps, cosr, sinr, qs + (j * RMAX + which) * HD, nullptr);
      } else if (owner) {
        if (which == R) {
          norm_rope_vec(base + (nh + kvh) * HD, kn, eps, cosr, sinr, nullptr, kc + (size_t)pos * HD);
        } else {
          const bf16* vsrc = base + (nh + nkv + kvh) * HD;
#pragma unroll
          for (int i = 0; i < 4; ++i) vc[(size_t)pos * HD + lane + 32 * i] = ldcg_bf16(vsrc + lane + 32 * i);
        }
      }
    }
    __threadfence_block();
    __syncthreads();
    PROF_MARK(2);

    const float scale = rsqrtf((float)HD);
    const int hw = warp * 2 + (lane >> 4), l16 = lane & 15;
    const int rr_ = tid >> 7, dd = tid & (HD - 1);
#pragma unroll 1
    for (int j = 0; j < nq; ++j) {
      const int col = q0 + j * qstride;
      const int pos = ctx_end - nq + j;
      const int e1 = min(s1, pos + 1);
      float q[RMAX][8];
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) q[r][i] = (r < R) ? qs[(j * RMAX + r) * HD + l16 * 8 + i] : 0.f;
      float m[RMAX], l[RMAX], o[RMAX][8];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[r][i] = 0.f;
      }
      // warp-uniform trip count (full-mask shuffles below); each half-warp handles 2 tokens per iteration so
      // that 4 independent 16-byte loads are in flight before the dependent softmax update
#pragma unroll 1
      for (int tb = s0 + warp * 4; tb < e1; tb += NWARPS * 4) {
        const int tk0 = tb + (lane >> 4), tk1 = tk0 + 2;
        uint4 kv[2], vv[2];
        kv[0] = kv[1] = vv[0] = vv[1] = make_uint4(0, 0, 0, 0);
        if (tk0 < e1) { kv[0] = ldcg16(kc + (size_t)tk0 * HD + l16 * 8); vv[0] = ldcg16(vc + (size_t)tk0 * HD + l16 * 8); }
        if (tk1 < e1) { kv[1] = ldcg16(kc + (size_t)tk1 * HD + l16 * 8); vv[1] = ldcg16(vc + (size_t)tk1 * HD + l16 * 8); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {  // fully unrolled: register arrays must keep compile-time indices (no local memory)
          const bool valid = (u == 0 ? tk0 : tk1) < e1;
          const uint4 kk = kv[u], v4 = vv[u];
          const float kf[8] = {bf16lo(kk.x), bf16hi(kk.x), bf16lo(kk.y), bf16hi(kk.y),
                               bf16lo(kk.z), bf16hi(kk.z), bf16lo(kk.w), bf16hi(kk.w)};
          const float vf[8] = {bf16lo(v4.x), bf16hi(v4.x), bf16lo(v4.y), bf16hi(v4.y),
                               bf16lo(v4.z), bf16hi(v4.z), bf16lo(v4.w), bf16hi(v4.w)};
#pragma unroll
          for (int r = 0; r < RMAX; ++r) {
            if (r < R) {
              float d = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) d += q[r][i] * kf[i];
              d += __shfl_xor_sync(0xffffffffu, d, 8);
              d += __shfl_xor_sync(0xffffffffu, d, 4);
              d += __shfl_xor_sync(0xffffffffu, d, 2);
              d += __shfl_xor_sync(0xffffffffu, d, 1);
              if (valid) {
                d *= scale;
                const float mn = fmaxf(m[r], d);
                const float corr = __expf(m[r] - mn);  // exp(-inf)=0 on the first token
                const float p = __expf(d - mn);
                l[r] = l[r] * corr + p;
#pragma unroll
                for (int i = 0; i < 8; ++i) o[r][i] = o[r][i] * corr + p * vf[i];
                m[r] = mn;
              }
            }
          }
        }
      }
      // ---- combine the 32 half-warps
      PROF_MARK(3);
      __syncthreads();
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R) {
          float* rr = red + ((size_t)hw * RMAX + r) * 130;
          if (l16 == 0) { rr[0] = m[r]; rr[1] = l[r]; }
#pragma unroll
          for (int i = 0; i < 8; ++i) rr[2 + l16 * 8 + i] = o[r][i];
        }
      }
      __syncthreads();
      float M = -INFINITY, L = 0.f, O = 0.f;
      // token t of this split maps to half-warp ((t>>2)<<1) | (t&1): only the first nhw half-warps hold data
      const int ntok = max(e1 - s0, 0);
      const int nhw = min(2 * NWARPS, ((ntok + 3) >> 2) << 1);
      if (rr_ < R) {
#pragma unroll 2
        for (int h2 = 0; h2 < nhw; ++h2) M = fmaxf(M, red[((size_t)h2 * RMAX + rr_) * 130]);
#pragma unroll 2
        for (int h2 = 0; h2 < nhw; ++h2) {
          const float* rp = red + ((size_t)h2 * RMAX + rr_) * 130;
          const float wgt = (rp[0] == -INFINITY) ? 0.f : __expf(rp[0] - M);
          L += rp[1] * wgt;
          O += rp[2 + dd] * wgt;
        }
      }
      PROF_MARK(4);
      bf16* outp = S.attn + (size_t)col * (nh * HD) + (kvh * R + rr_) * HD + dd;
      if (nsplit == 1) {
        if (rr_ < R) *outp = f2bf(O / L);
      } else {
        // cross-CTA split combine: publish (M,L,O) and let the last arriver finish (deterministic order)
        float* sb0 = P.split_buf + (((size_t)(seq * nkv + kvh) * MAXSPLIT) * RMAX) * 130;
        float* sb = sb0 + ((size_t)sp * RMAX) * 130;
        if (rr_ < R) {
          if (dd == 0) { sb[rr_ * 130] = M; sb[rr_ * 130 + 1] = L; }
          sb[rr_ * 130 + 2 + dd] = O;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_ticket = (int)atomicAdd(&P.st->split_cnt[seq * nkv + kvh], 1u);
        __syncthreads();
        if (s_ticket == nsplit - 1) {
          __threadfence();
          if (rr_ < R) {
            float M2 = -INFINITY, L2 = 0.f, O2 = 0.f;
#pragma unroll 1
            for (int s2 = 0; s2 < nsplit; ++s2) M2 = fmaxf(M2, ldcgf(sb0 + ((size_t)s2 * RMAX + rr_) * 130));
#pragma unroll 1
            for (int s2 = 0; s2 < nsplit; ++s2) {
              const float* rp = sb0 + ((size_t)s2 * RMAX + rr_) * 130;
              const float mm = ldcgf(rp);
              const float wgt = (mm == -INFINITY) ? 0.f : __expf(mm - M2);
              L2 += ldcgf(rp + 1) * wgt;
              O2 += ldcgf(rp + 2 + dd) * wgt;
            }
            *outp = f2bf(O2 / L2);
          }
          if (tid == 0) P.st->split_cnt[seq * nkv + kvh] = 0;
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// sample phase: HF logits processors + argmax / inverse-CDF sampling + next-embed (one CTA per row)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int fkey(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// k-th largest of sv[0..V) (block-wide radix select over 4x8 bits); returns the threshold value
__device__ __noinline__ float kth_largest(const float* sv, int V, int k, unsigned int* hist, int* sh) {
  unsigned int prefix = 0, mask = 0;
#pragma unroll 1
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = pass * 8;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll 1
    for (int i = threadIdx.x; i < V; i += NTHREADS) {
      const unsigned int key = fkey(sv[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    // suffix counts over the 256 bins (8 warps)
    unsigned int cnt = 0, incl = 0;
    if (threadIdx.x < 256) {
      cnt = hist[threadIdx.x];
      incl = cnt;
      const int ln = threadIdx.x & 31;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned int n = __shfl_down_sync(0xffffffffu, incl, o);
        if (ln + o < 32) incl += n;
      }
      if (ln == 0) sh[threadIdx.x >> 5] = (int)incl;  // warp totals
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      unsigned int above = 0;
#pragma unroll 1
      for (int w = (threadIdx.x >> 5) + 1; w < 8; ++w) above += (unsigned int)sh[w];
      incl += above;                       // elements with digit >= d
      const unsigned int excl = incl - cnt;  // elements with digit > d
      if ((int)excl < k && k <= (int)incl) { sh[8] = threadIdx.x; sh[9] = k - (int)excl; }
    }
    __syncthreads();
    prefix |= ((unsigned int)sh[8]) << shift;
    mask |= 255u << shift;
    k = sh[9];
    __syncthreads();
  }
  const unsigned int u = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix;
  return __uint_as_float(u);
}

__device__ __noinline__ void sample_phase(const Phase& ph, const KParams& P, unsigned char* smem, int frame, bool in_prefill) {
  const int b = blockIdx.x;
  DevState* st = P.st;
  const int B = P.B;
  if (b >= B) return;
  const int tid = threadIdx.x;
  const int group = ph.group;
  const bool talker = group == 0;
  const StackDev& S = talker ? P.talker : P.cp;
  const int V = S.vocab;
  const int H = P.talker.hidden;
  float* sv = reinterpret_cast<float*>(smem);       // [MAXV]
  float* pv = sv + MAXV;                            // [MAXV]
  float* red = pv + MAXV;                           // [32]
  int* ired = reinterpret_cast<int*>(red + 32);     // [32]
  unsigned int* hist = reinterpret_cast<unsigned int*>(ired + 32);  // [256]
  // frame index of the token being sampled
  const int fidx = talker ? (in_prefill ? 0 : frame + 1) : frame;
  const bool do_sample = talker ? P.sp.do_sample : P.sp.subtalker_dosample;
  const float temperature = talker ? P.sp.temperature : P.sp.subtalker_temperature;
  const int top_k = talker ? P.sp.top_k : P.sp.subtalker_top_k;
  const float top_p = talker ? P.sp.top_p : P.sp.subtalker_top_p;
  const float* lg = S.logits + (size_t)b * V;
  const int n_gen_b = talker ? ldcgi(&st->n_gen[b]) : 0;

  // raw logits -> smem with all loads of a thread in flight, then a rolled processing pass
  {
    float lreg[MAXV / NTHREADS];
#pragma unroll
    for (int r = 0; r < MAXV / NTHREADS; ++r) { const int i = tid + r * NTHREADS; if (i < V) lreg[r] = ldcgf(lg + i); }
#pragma unroll
    for (int r = 0; r < MAXV / NTHREADS; ++r) { const int i = tid + r * NTHREADS; if (i < V) sv[i] = lreg[r]; }
  }
  float* dbg = talker ? (P.dbg_tlogits ? P.dbg_tlogits + ((size_t)fidx * B + b) * V : nullptr)
                      : (P.dbg_clogits ? P.dbg_clogits + (((size_t)frame * (P.G - 1) + (group - 1)) * B + b) * V : nullptr);
  const float rp = P.sp.repetition_penalty;
  const float inv_t = (do_sample && temperature != 1.0f) ? temperature : 1.0f;
#pragma unroll 1
  for (int i = tid; i < V; i += NTHREADS) {
    float s = sv[i];
    if (dbg) dbg[i] = s;
    if (talker) {
      // 1. repetition penalty over generated codebook-0 tokens
      if (rp != 1.0f && __ldcg(P.seen + (size_t)b * V + i)) s = s < 0.f ? s * rp : s / rp;
      // 2. min_new_tokens (and the fixed-horizon benchmark switch)
      if (i == P.eos && (n_gen_b < P.sp.min_new_tokens || P.sp.suppress_eos)) s = -INFINITY;
      // 3. suppress [V-1024, V) \ {eos}
      if (i >= V - 1024 && i != P.eos) s = -INFINITY;
    }
    if (inv_t != 1.0f) s = s / inv_t;
    sv[i] = s;
  }
  __syncthreads();

  int tok;
  if (!do_sample) {
    float mx = -INFINITY;
#pragma unroll 1
    for (int i = tid; i < V; i += NTHREADS) mx = fmaxf(mx, sv[i]);
    mx = block_reduce(mx, red, 0);
    int idx = 0x7fffffff;
#pragma unroll 1
    for (int i = tid; i < V; i += NTHREADS)
      if (sv[i] == mx) idx = min(idx, i);
    tok = block_min_int(idx, ired);
  } else {
    if (top_k > 0 && top_k < V) {
      const float thr = kth_largest(sv, V, top_k, hist, ired);
#pragma unroll 1
      for (int i = tid; i < V; i += NTHREADS)
        if (sv[i] < thr) sv[i] = -INFINITY;
      __syncthreads();
    }
    float mx = -INFINITY;
#pragma unroll 1
    for (int i = tid; i < V; i += NTHREADS) mx = fmaxf(mx, sv[i]);
    mx = block_reduce(mx, red, 0);
    if (top_p < 1.0f) {
      // ascending-order cumulative softmax <= 1-p is removed, highest kept (O(n^2) over the kept set)
      float tot = 0.f;
#pragma unroll 1
      for (int i = tid; i < V; i += NTHREADS) { const float p = __expf(sv[i] - mx); pv[i] = p; tot += p; }
      tot = block_reduce(tot, red, 1);
      __syncthreads();
      unsigned int rm_mask = 0;  // removal flags stay in registers until every thread has finished reading pv/sv
      int slot = 0;
#pragma unroll 1
      for (int i = tid; i < V; i += NTHREADS, ++slot) {
        const float si = sv[i];
        bool rm = false;
        if (si != -INFINITY) {
          float cum = 0.f;
          bool is_top = true;
#pragma unroll 1
          for (int j = 0; j < V; ++j) {
            const float sj = sv[j];
            if (sj == -INFINITY) continue;
            if (sj < si || (sj == si && j <= i)) cum += pv[j];
            if (sj > si || (sj == si && j > i)) is_top = false;
          }
          rm = (cum / tot <= 1.0f - top_p) && !is_top;
        }
        if (rm) rm_mask |= 1u << slot;
      }
      __syncthreads();
      slot = 0;
#pragma unroll 1
      for (int i = tid; i < V; i += NTHREADS, ++slot)
        if (rm_mask & (1u << slot)) sv[i] = -INFINITY;
      __syncthreads();
    }
    // softmax + inverse CDF in token-id order (blocked mapping for the scan)
    const int E = (V + NTHREADS - 1) / NTHREADS;
    const int i0 = tid * E, i1 = min(V, i0 + E);
    float loc = 0.f;
#pragma unroll 1
    for (int i = i0; i < i1; ++i) { const float p = __expf(sv[i] - mx); pv[i] = p; loc += p; }
    // block inclusive scan of loc
    float incl = loc;
    const int ln = tid & 31, wp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, incl, o);
      if (ln >= o) incl += n;
    }
    __syncthreads();
    if (ln == 31) red[wp] = incl;
    __syncthreads();
    float base = 0.f, total = 0.f;
#pragma unroll 1
    for (int w = 0; w < NWARPS; ++w) { if (w < wp) base += red[w]; total += red[w]; }
    const float excl = base + incl - loc;
    const float u = philox_uniform(P.sp.seed, (uint32_t)b, (uint32_t)fidx, (uint32_t)group);
    const float target = u * total;
    int cand = 0x7fffffff, lastpos = -1;
    float run = excl;
#pragma unroll 1
    for (int i = i0; i < i1; ++i) {
      run += pv[i];
      if (pv[i] > 0.f) { lastpos = i; if (run > target && cand == 0x7fffffff) cand = i; }
    }
    cand = block_min_int(cand, ired);
    if (cand == 0x7fffffff) cand = -block_min_int(-lastpos, ired);
    tok = cand;
  }
  // teacher forcing (tests)
  if (P.forced && fidx < P.n_forced) {
    const int f = P.forced[((size_t)b * P.n_forced + fidx) * P.G + group];
    if (f >= 0) tok = f;
  }

  if (talker) {
    const int was_finished = ldcgi(&st->finished[b]);
    if (was_finished) tok = P.eos;  // HF pads finished rows with pad_token_id (= eos)
    __syncthreads();
    if (tid == 0) {
      if (!was_finished) {
        if (tok == P.eos) { st->finished[b] = 1; st->n_valid[b] = fidx; }
        else { P.seen[(size_t)b * V + tok] = 1; }
        st->n_gen[b] = n_gen_b + 1;
      }
      st->c0[b] = tok;
      st->cur[b][0] = tok;
    }
    // CP input for the next frame: token 0 = past_hidden (already saved by the head phase), token 1 = E0[c0]
    bf16* x1 = P.x_cp + ((size_t)B + b) * H;
    bf16* x0 = P.x_cp + (size_t)b * H;
    const bf16* e = P.emb_t + (size_t)tok * H;
    const bf16* ph_ = P.past_hidden + (size_t)b * H;
#pragma unroll 1
    for (int i = tid * 8; i < H; i += NTHREADS * 8) {
      *reinterpret_cast<uint4*>(x1 + i) = *reinterpret_cast<const uint4*>(e + i);
      *reinterpret_cast<uint4*>(x0 + i) = ldcg16(ph_ + i);
    }
  } else {
    const int j = group;  // codebook index 1..G-1
    if (tid == 0) {
      st->cur[b][j] = tok;
      if (P.codes_out && frame < P.codes_stride) {
        int* row = P.codes_out + ((size_t)b * P.codes_stride + frame) * P.G;
        row[j] = tok;
        if (j == 1) row[0] = ldcgi(&st->cur[b][0]);
      }
    }
    __syncthreads();
    const int Vc = P.cp.vocab;
    if (j < P.G - 1) {
      // input of the next pass: codec_embedding[j-1](c_j)  (:1281)
      // (with the projection table the row is small_to_mtp_projection(embedding) already, :1283, and lands in cp.h)
      const int Wn = P.cp_next_w;
      const bf16* e = P.cp_next + ((size_t)(j - 1) * Vc + tok) * Wn;
      bf16* x = P.cp_next_dst + (size_t)b * Wn;
#pragma unroll 1
      for (int i = tid * 8; i < Wn; i += NTHREADS * 8)
        *reinterpret_cast<uint4*>(x + i) = *reinterpret_cast<const uint4*>(e + i);
    } else {
      // next talker input: sum of the 16 codebook embeddings (fp32 sum, one bf16 rounding) + text (:1682-1692)
      const bf16* txt = (frame < P.trailing_len[b])
                            ? P.trailing + ((size_t)b * P.trailing_stride + frame) * H
                            : P.tts_pad;
      int* codes = reinterpret_cast<int*>(hist);  // smem scratch: the 16 codes of this frame
      if (tid < P.G) codes[tid] = (tid == j) ? tok : ldcgi(&st->cur[b][tid]);
      __syncthreads();
#pragma unroll 1
      for (int i = tid; i < H; i += NTHREADS) {
        float s = bf2f(P.emb_t[(size_t)codes[0] * H + i]);
#pragma unroll 4
        for (int g2 = 1; g2 < P.G; ++g2) s += bf2f(P.emb_cp[((size_t)(g2 - 1) * Vc + codes[g2]) * H + i]);
        P.talker.h[(size_t)b * H + i] = f2bf(rbf(s) + bf2f(txt[i]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NTHREADS, 1) q3_program_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Phase s_ph[2];
  __shared__ uint4 s_nw[2][256];  // RMSNorm weights of the current / next GEMV phase (<= 2048 bf16)
  DevState* st = P.st;
  unsigned int epoch = 0;  // host resets bar_count to 0 before every launch
  if (threadIdx.x < (int)(sizeof(Phase) / 4))
    reinterpret_cast<uint32_t*>(&s_ph[0])[threadIdx.x] = reinterpret_cast<const uint32_t*>(P.prog)[threadIdx.x];
  __syncthreads();
  if (s_ph[0].type == PH_GEMV && s_ph[0].norm_w != nullptr && (int)threadIdx.x < s_ph[0].kb * 4)
    s_nw[0][threadIdx.x] = reinterpret_cast<const uint4*>(s_ph[0].norm_w)[threadIdx.x];
  __syncthreads();
  const int step_base = st->step;
  int iters_done = 0;
  int slot = 0;
  const int niter = P.mode == 1 ? P.max_iters : 1;
#pragma unroll 1
  for (int it = 0; it < niter; ++it) {
    const int frame = step_base + it;
    if (P.mode == 1) {
      bool all = true;
      for (int b = 0; b < P.B; ++b) all = all && (ldcgi(&st->finished[b]) != 0);
      if (all) break;
    }
#pragma unroll 1
    for (int pi = 0; pi < P.n_phases; ++pi) {
      // descriptor of this phase sits in s_ph[slot] (fetched one phase ahead).  The load of the NEXT descriptor is
      // issued now into a register and only stored to smem after the body, so its L2 round trip is hidden.
      int nx = pi + 1;
      if (nx >= P.n_phases) nx = (P.mode == 1) ? 0 : -1;
      const bool dhave = nx >= 0 && threadIdx.x < (int)(sizeof(Phase) / 4);
      uint32_t dreg = 0;
      if (dhave) dreg = reinterpret_cast<const uint32_t*>(P.prog + nx)[threadIdx.x];
      const Phase& ph = s_ph[slot];
      if (threadIdx.x == 0) {
        g_prof_row = (P.prof && it == 0) ? P.prof + ((size_t)pi * gridDim.x + blockIdx.x) * 8 : nullptr;
        PROF_MARK(6);
      }
      const int type = ph.type;
      if (type == PH_GEMV) gemv_phase<NT>(ph, P, smem, s_nw[slot]);
      else if (type == PH_ATTN) attn_phase(ph, P, smem, frame);
      else sample_phase(ph, P, smem, frame, P.mode == 0);
      if (dhave) reinterpret_cast<uint32_t*>(&s_ph[slot ^ 1])[threadIdx.x] = dreg;
      __syncthreads();
      // pull the next GEMV's weight slice toward L2 while the barrier drains, and fetch its norm weights (the
      // load is issued before the barrier, the smem store happens after it: zero exposed latency)
      uint4 nwv = make_uint4(0, 0, 0, 0);
      bool nw_have = false;
      if (nx >= 0) {
        const Phase& nph = s_ph[slot ^ 1];
        if (!(P.dbg_skip & 32)) prefetch_phase_weights(nph);
        if (!(P.dbg_skip & 64) && nph.type == PH_GEMV && nph.norm_w != nullptr && (int)threadIdx.x < nph.kb * 4) {
          nwv = reinterpret_cast<const uint4*>(nph.norm_w)[threadIdx.x];
          nw_have = true;
        }
      }
      PROF_MARK(0);
      grid_barrier(st, epoch);
      slot ^= 1;
      if (nw_have) s_nw[slot][threadIdx.x] = nwv;  // visible to the next phase after its first __syncthreads... see below
      __syncthreads();
      PROF_MARK(1);
    }
    ++iters_done;
  }
  if (P.mode == 1 && blockIdx.x == 0 && threadIdx.x == 0) st->step = step_base + iters_done;
}

// ------------------------------------------------------------------------------------------------
// PREFILL on tensor cores: talker linears are tcgen05 tap-GEMMs (gemm_sm100.cu) over all prompt tokens at once
// (M = sum of prompt lengths); the row-wise pieces around them are the small kernels below.
// ------------------------------------------------------------------------------------------------
__global__ void pf_rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y, int rows, int C,
                                  float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * C);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  float ss = 0.f;
  for (int i = lane; i < C / 8; i += 32) {
    const uint4 v = xr[i];
    float f;
    f = bf16lo(v.x); ss += f * f; f = bf16hi(v.x); ss += f * f;
    f = bf16lo(v.y); ss += f * f; f = bf16hi(v.y); ss += f * f;
    f = bf16lo(v.z); ss += f * f; f = bf16hi(v.z); ss += f * f;
    f = bf16lo(v.w); ss += f * f; f = bf16hi(v.w); ss += f * f;
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)C + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * C);
  for (int i = lane; i < C / 8; i += 32) {
    const uint4 v = xr[i], wv = wr[i];
    uint4 o;
    o.x = pack_bf16(rbf(bf16lo(v.x) * inv) * bf16lo(wv.x), rbf(bf16hi(v.x) * inv) * bf16hi(wv.x));
    o.y = pack_bf16(rbf(bf16lo(v.y) * inv) * bf16lo(wv.y), rbf(bf16hi(v.y) * inv) * bf16hi(wv.y));
    o.z = pack_bf16(rbf(bf16lo(v.z) * inv) * bf16lo(wv.z), rbf(bf16hi(v.z) * inv) * bf16hi(wv.z));
    o.w = pack_bf16(rbf(bf16lo(v.w) * inv) * bf16lo(wv.w), rbf(bf16hi(v.w) * inv) * bf16hi(wv.w));
    yr[i] = o;
  }
}

// per (token, head-vector): q heads RMSNorm+RoPE in place; k head -> K cache (normed, roped); v head -> V cache
__global__ void pf_qkv_post_kernel(bf16* __restrict__ qkv, int ntok, const int* __restrict__ tok_seq, const int* __restrict__ tok_pos,
                                   int nh, int nkv, const bf16* __restrict__ qn, const bf16* __restrict__ kn, float eps,
                                   const bf16* __restrict__ cosT, const bf16* __restrict__ sinT, bf16* __restrict__ kc,
                                   bf16* __restrict__ vc, int layer, int layers, int cap) {
  const int nvec = nh + 2 * nkv;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= ntok * nvec) return;
  const int lane = threadIdx.x & 31;
  const int tok = wid / nvec, v = wid - tok * nvec;
  const int seq = tok_seq[tok], pos = tok_pos[tok];
  bf16* src = qkv + (size_t)tok * (size_t)(nvec * HD) + (size_t)v * HD;
  const bf16* cosr = cosT + (size_t)pos * 64;
  const bf16* sinr = sinT + (size_t)pos * 64;
  if (v < nh) {
    norm_rope_vec(src, qn, eps, cosr, sinr, nullptr, src);
  } else if (v < nh + nkv) {
    const int kvh = v - nh;
    norm_rope_vec(src, kn, eps, cosr, sinr, nullptr, kc + ((((size_t)seq * layers + layer) * nkv + kvh) * cap + pos) * HD);
  } else {
    const int kvh = v - nh - nkv;
    bf16* d = vc + ((((size_t)seq * layers + layer) * nkv + kvh) * cap + pos) * HD;
#pragma unroll
    for (int i = 0; i < 4; ++i) d[lane + 32 * i] = src[lane + 32 * i];
  }
}

// causal attention over the KV cache, one warp per (token, q head); keys in blocks of 32 with an online softmax
__global__ void pf_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ attn, int ntok, const int* __restrict__ tok_seq,
                                    const int* __restrict__ tok_pos, int nh, int nkv, const bf16* __restrict__ kc,
                                    const bf16* __restrict__ vc, int layer, int layers, int cap) {
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= ntok * nh) return;
  const int lane = threadIdx.x & 31;
  const int tok = wid / nh, h = wid - tok * nh;
  const int seq = tok_seq[tok], pos = tok_pos[tok];
  const int kvh = h / (nh / nkv);
  const bf16* q = qkv + (size_t)tok * (size_t)((nh + 2 * nkv) * HD) + (size_t)h * HD;
  const bf16* K = kc + (((size_t)seq * layers + layer) * nkv + kvh) * (size_t)cap * HD;
  const bf16* V = vc + (((size_t)seq * layers + layer) * nkv + kvh) * (size_t)cap * HD;
  const float scale = rsqrtf((float)HD);
  float m = -INFINITY, l = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
  for (int kb = 0; kb <= pos; kb += 32) {
    const int key = kb + lane;
    float sc = -INFINITY;
    if (key <= pos) {
      const uint4* kr = reinterpret_cast<const uint4*>(K + (size_t)key * HD);
      const uint4* qr = reinterpret_cast<const uint4*>(q);
      float d = 0.f;
#pragma unroll 4
      for (int i = 0; i < HD / 8; ++i) {
        const uint4 a = qr[i], b = kr[i];
        d += bf16lo(a.x) * bf16lo(b.x) + bf16hi(a.x) * bf16hi(b.x) + bf16lo(a.y) * bf16lo(b.y) + bf16hi(a.y) * bf16hi(b.y) +
             bf16lo(a.z) * bf16lo(b.z) + bf16hi(a.z) * bf16hi(b.z) + bf16lo(a.w) * bf16lo(b.w) + bf16hi(a.w) * bf16hi(b.w);
      }
      sc = d * scale;
    }
    const float mn = fmaxf(m, warp_max(sc));
    const float corr = __expf(m - mn);
    const float p = (key <= pos) ? __expf(sc - mn) : 0.f;
    l = l * corr + warp_sum(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] *= corr;
    const int nk = min(32, pos - kb + 1);
    for (int j = 0; j < nk; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const uint2 vv = *reinterpret_cast<const uint2*>(V + (size_t)(kb + j) * HD + lane * 4);
      o[0] += pj * bf16lo(vv.x); o[1] += pj * bf16hi(vv.x); o[2] += pj * bf16lo(vv.y); o[3] += pj * bf16hi(vv.y);
    }
    m = mn;
  }
  const float inv = 1.f / l;
  uint2 r;
  r.x = pack_bf16(o[0] * inv, o[1] * inv);
  r.y = pack_bf16(o[2] * inv, o[3] * inv);
  *reinterpret_cast<uint2*>(attn + (size_t)tok * (size_t)(nh * HD) + (size_t)h * HD + lane * 4) = r;
}

// ------------------------------------------------------------------------------------------------
// weight packing: row-major [N][K] bf16 -> stream of (16 rows x 32 k) 1 KB blocks, each two 8x32 halves
// ------------------------------------------------------------------------------------------------
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = bf2f(src[i]);
}

__global__ void pack_weight_kernel(const bf16* __restrict__ src, uint4* __restrict__ dst, int N, int K) {
  // one thread per 16-byte chunk of the destination
  const size_t total = (size_t)N * K / 8;
  const int KB = K / 32;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t blk = i / 64;         // (tile, kb)
    const int within = (int)(i % 64);  // half j (32 chunks each), then g*4 + t
    const int j = within / 32, g = (within % 32) / 4, t = within % 4;
    const size_t tile = blk / KB;
    const int kb = (int)(blk % KB);
    const size_t row = tile * 16 + j * 8 + g;
    dst[i] = *reinterpret_cast<const uint4*>(src + row * K + kb * 32 + t * 8);
  }
}

}  // namespace

// =================================================================================================
// host side
// =================================================================================================
struct PackedW {
  uint4* w = nullptr;
  int n = 0, k = 0;
};

struct q3_engine {
  q3_engine_cfg cfg;
  int sm_count = 0;
  cudaStream_t copy_stream = nullptr;
  std::map<std::string, PackedW> packed;   // GEMV weights
  std::map<std::string, bf16*> gemm_w;      // plain [N][K] copies of the talker layer weights for the tcgen05 prefill GEMMs
  bf16 *pf_x = nullptr, *pf_xn = nullptr, *pf_qkv = nullptr, *pf_attn = nullptr, *pf_act = nullptr;
  int *pf_seq = nullptr, *pf_pos = nullptr;
  int pf_cap = 0;
  std::map<std::string, bf16*> plain;      // norms, biases, embeddings, rope tables
  std::map<std::string, std::pair<int64_t, int64_t>> plain_shape;
  std::vector<void*> allocs;
  // device buffers
  DevState* st = nullptr;
  StackDev talker{}, cp{};
  bf16 *x_cp = nullptr, *past_hidden = nullptr, *h_last = nullptr;
  bf16 *trailing = nullptr, *tts_pad = nullptr;
  int trailing_cap = 0;
  size_t trailing_alloc = 0;
  int max_len0 = 0, frames_issued = 0;
  int len0[MAXB] = {0}, trailing_len[MAXB] = {0};
  unsigned char* seen = nullptr;
  float* split_buf = nullptr;
  Phase* prog_dev = nullptr;
  int prog_cap = 0;
  // programs (host copies) for the current batch size
  int prog_B = -1;
  int dbg_skip = 0;
  bool use_proj_tab = true;   // Q3_CP_PROJ_TAB=0 keeps the projection GEMV in passes >= 1 (A/B knob)
  bf16* proj_tab = nullptr;   // small_to_mtp_projection(cp.codec_embedding) [(G-1)*Vc][Hc], built once at finalize
  std::vector<Phase> prog_layers, prog_head, prog_frame;
  int off_layers = 0, off_head = 0, off_frame = 0;
  q3_sampling sp{};
  bool finalized = false;
  int B = 0;
  int codes_stride = 0;
  const int* forced = nullptr;
  int n_forced = 0;
  float *dbg_t = nullptr, *dbg_c = nullptr;
  unsigned long long* prof = nullptr;
  size_t smem_bytes = 0;
  double w_talker_bytes = 0, w_cp_unique_bytes = 0, w_cp_stream_bytes = 0;

  template <typename T>
  int alloc(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return q3_set_err("cudaMalloc(%zu B) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    cudaMemset(q, 0, count * sizeof(T));
    allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return 0;
  }
};

static int stack_alloc(q3_engine* e, StackDev& S, const q3_stack_cfg& c, int nseq, int cap, int cols) {
  S.hidden = c.hidden_size; S.layers = c.num_layers; S.nh = c.num_heads; S.nkv = c.num_kv_heads;
  S.inter = c.intermediate_size; S.vocab = c.vocab_size; S.eps = c.rms_eps; S.cap = cap;
  if (e->alloc(&S.h, (size_t)cols * c.hidden_size)) return 1;
  if (e->alloc(&S.qkv, (size_t)cols * (c.num_heads + 2 * c.num_kv_heads) * HD)) return 1;
  if (e->alloc(&S.attn, (size_t)cols * c.num_heads * HD)) return 1;
  if (e->alloc(&S.act, (size_t)cols * c.intermediate_size)) return 1;
  const size_t kv = (size_t)nseq * c.num_layers * c.num_kv_heads * cap * HD;
  if (e->alloc(&S.kc, kv)) return 1;
  if (e->alloc(&S.vc, kv)) return 1;
  if (e->alloc(&S.logits, (size_t)MAXB * c.vocab_size)) return 1;
  return 0;
}

extern "C" int q3_abi_version(void) { return Q3_ABI_VERSION; }
extern "C" const char* q3_last_error(void) { return g_q3_err.c_str(); }

extern "C" int q3_engine_create(const q3_engine_cfg* cfg, q3_engine** out) {
  Q3_REQUIRE(cfg && out, "null argument");
  Q3_REQUIRE(cfg->talker.head_dim == HD && cfg->cp.head_dim == HD, "head_dim must be 128");
  Q3_REQUIRE(cfg->max_batch >= 1 && cfg->max_batch <= MAXB, "max_batch must be in [1,%d]", MAXB);
  Q3_REQUIRE(cfg->num_code_groups >= 2 && cfg->num_code_groups <= Q3_NUM_GROUPS_MAX, "bad num_code_groups");
  Q3_REQUIRE(cfg->talker.vocab_size <= MAXV && cfg->cp.vocab_size <= MAXV, "vocab > %d unsupported", MAXV);
  Q3_REQUIRE(cfg->talker.vocab_size >= 1024, "talker vocab must be >= 1024 (suppress range)");
  for (const q3_stack_cfg* s : {&cfg->talker, &cfg->cp}) {
    Q3_REQUIRE(s->hidden_size % 32 == 0 && s->intermediate_size % 32 == 0 && s->vocab_size % 16 == 0,
               "hidden/intermediate must be multiples of 32 and vocab of 16");
    Q3_REQUIRE(s->num_heads % s->num_kv_heads == 0 && s->num_heads / s->num_kv_heads <= RMAX, "GQA group (heads / kv_heads) must be <= 2");
    Q3_REQUIRE(xs_stride_bytes(s->hidden_size) * MAXCOLS <= XS_BYTES && xs_stride_bytes(s->num_heads * HD) * MAXCOLS <= XS_BYTES,
               "hidden_size / num_heads*head_dim above 2048 unsupported");
  }
  Q3_REQUIRE(cfg->has_cp_projection || cfg->talker.hidden_size == cfg->cp.hidden_size,
             "Identity projection requires equal hidden sizes");
  Q3_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  Q3_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  Q3_REQUIRE(prop.major == 10, "this library is built for sm_100a (B200); device is sm_%d%d", prop.major, prop.minor);
  q3_engine* e = new q3_engine();
  e->cfg = *cfg;
  e->sm_count = prop.multiProcessorCount;
  if (const char* g = getenv("Q3_GRID")) {  // experiment knob: persistent grid smaller than the SM count
    const int v = atoi(g);
    if (v >= 8 && v <= e->sm_count) e->sm_count = v;
  }
  if (const char* f = getenv("Q3_CP_PROJ_TAB")) e->use_proj_tab = atoi(f) != 0;
  Q3_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  if (gemm_init()) return 1;
  if (e->alloc(&e->st, 1)) return 1;
  if (stack_alloc(e, e->talker, cfg->talker, cfg->max_batch, cfg->max_ctx, MAXCOLS)) return 1;
  if (stack_alloc(e, e->cp, cfg->cp, cfg->max_batch, 32, MAXCOLS)) return 1;
  const int H = cfg->talker.hidden_size;
  if (e->alloc(&e->x_cp, (size_t)2 * MAXB * H)) return 1;
  if (e->alloc(&e->past_hidden, (size_t)MAXB * H)) return 1;
  if (e->alloc(&e->h_last, (size_t)MAXB * H)) return 1;
  if (e->alloc(&e->tts_pad, (size_t)H)) return 1;
  if (e->alloc(&e->seen, (size_t)MAXB * cfg->talker.vocab_size)) return 1;
  if (e->alloc(&e->split_buf, (size_t)MAXB * cfg->talker.num_kv_heads * MAXSPLIT * RMAX * 130)) return 1;
  e->smem_bytes = smem_bytes_nt(4);
  static_assert(SAMPLER_SMEM <= ATT_SMEM, "sampler must fit the minimum shared-memory request");
  Q3_REQUIRE(e->smem_bytes <= (size_t)prop.sharedMemPerBlockOptin, "not enough shared memory per block");
  *out = e;
  return 0;
}

extern "C" void q3_engine_destroy(q3_engine* e) {
  if (!e) return;
  for (void* p : e->allocs) cudaFree(p);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  delete e;
}

static bool is_gemv_weight(const std::string& n) {
  auto ends = [&](const char* s) { size_t l = strlen(s); return n.size() >= l && n.compare(n.size() - l, l, s) == 0; };
  if (ends(".qkv") || ends(".gate_up") || ends(".o") || ends(".down")) return true;
  if (n == "talker.codec_head" || n == "cp.proj") return true;
  if (n.rfind("cp.lm_head.", 0) == 0) return true;
  return false;
}

extern "C" int q3_engine_load_tensor(q3_engine* e, const char* name, const void* dev, int64_t rows, int64_t cols) {
  Q3_REQUIRE(e && name && dev, "null argument");
  Q3_CUDA(cudaSetDevice(e->cfg.device));
  std::string n(name);
  if (is_gemv_weight(n)) {
    Q3_REQUIRE(rows % 16 == 0 && cols % 32 == 0, "%s: GEMV weight must be [16a][32b], got [%lld][%lld]", name,
               (long long)rows, (long long)cols);
    PackedW pw;
    pw.n = (int)rows; pw.k = (int)cols;
    if (e->alloc(&pw.w, (size_t)rows * cols / 8)) return 1;
    pack_weight_kernel<<<1024, 256, 0, e->copy_stream>>>(reinterpret_cast<const bf16*>(dev), pw.w, (int)rows, (int)cols);
    Q3_CUDA(cudaGetLastError());
    Q3_CUDA(cudaStreamSynchronize(e->copy_stream));
    e->packed[n] = pw;
    if (n.rfind("talker.layers.", 0) == 0 || (n == "cp.proj" && cols % 64 == 0)) {  // tcgen05 GEMM operand: [N][K] is already K-major
      Q3_REQUIRE(cols % 64 == 0, "%s: K must be a multiple of 64 for the prefill GEMM", name);
      bf16* g = nullptr;
      if (e->alloc(&g, (size_t)rows * cols)) return 1;
      Q3_CUDA(cudaMemcpyAsync(g, dev, (size_t)rows * cols * sizeof(bf16), cudaMemcpyDeviceToDevice, e->copy_stream));
      Q3_CUDA(cudaStreamSynchronize(e->copy_stream));
      e->gemm_w[n] = g;
    }
  } else {
    bf16* p = nullptr;
    if (e->alloc(&p, (size_t)rows * cols)) return 1;
    Q3_CUDA(cudaMemcpyAsync(p, dev, (size_t)rows * cols * sizeof(bf16), cudaMemcpyDeviceToDevice, e->copy_stream));
    Q3_CUDA(cudaStreamSynchronize(e->copy_stream));
    e->plain[n] = p;
    e->plain_shape[n] = {rows, cols};
  }
  return 0;
}

static int need_packed(q3_engine* e, const std::string& n, int rows, int cols, PackedW* out) {
  auto it = e->packed.find(n);
  Q3_REQUIRE(it != e->packed.end(), "missing tensor %s", n.c_str());
  Q3_REQUIRE(it->second.n == rows && it->second.k == cols, "tensor %s has shape [%d][%d], expected [%d][%d]", n.c_str(),
             it->second.n, it->second.k, rows, cols);
  *out = it->second;
  return 0;
}
static int need_plain(q3_engine* e, const std::string& n, int64_t count, const bf16** out) {
  auto it = e->plain.find(n);
  Q3_REQUIRE(it != e->plain.end(), "missing tensor %s", n.c_str());
  auto sh = e->plain_shape[n];
  Q3_REQUIRE(sh.first * sh.second == count, "tensor %s has %lld elements, expected %lld", n.c_str(),
             (long long)(sh.first * sh.second), (long long)count);
  *out = it->second;
  return 0;
}

// ---- program construction ------------------------------------------------------------------------
static Phase gemv(const PackedW& w, const bf16* src, int src_ld, const bf16* norm_w, float eps, void* dst, int dst_ld,
                  int epi, int ncmode, const bf16* bias = nullptr, bf16* save_normed = nullptr) {
  Phase p{};
  p.type = PH_GEMV; p.epi = epi; p.ncmode = ncmode;
  p.w = w.w; p.n_tiles = w.n / 16; p.kb = w.k / 32;
  p.src = src; p.src_ld = src_ld; p.norm_w = norm_w; p.eps = eps; p.dst = dst; p.dst_ld = dst_ld; p.bias = bias;
  p.save_normed = save_normed;
  return p;
}

static int add_layers(q3_engine* e, std::vector<Phase>& prog, const char* pfx, StackDev& S, bf16* hbuf, int ncmode,
                      int seqmode, int nq, int ctx_end) {
  const int Hh = S.hidden, QKV = (S.nh + 2 * S.nkv) * HD;
  for (int l = 0; l < S.layers; ++l) {
    std::string p = std::string(pfx) + ".layers." + std::to_string(l);
    PackedW wqkv, wo, wgu, wd;
    const bf16 *ln1, *ln2, *qn, *kn;
    if (need_packed(e, p + ".qkv", QKV, Hh, &wqkv) || need_packed(e, p + ".o", Hh, S.nh * HD, &wo) ||
        need_packed(e, p + ".gate_up", 2 * S.inter, Hh, &wgu) || need_packed(e, p + ".down", Hh, S.inter, &wd) ||
        need_plain(e, p + ".ln1", Hh, &ln1) || need_plain(e, p + ".ln2", Hh, &ln2) ||
        need_plain(e, p + ".q_norm", HD, &qn) || need_plain(e, p + ".k_norm", HD, &kn))
      return 1;
    prog.push_back(gemv(wqkv, hbuf, Hh, ln1, S.eps, S.qkv, QKV, EPI_STORE, ncmode));
    Phase a{};
    a.type = PH_ATTN; a.stack = (&S == &e->talker) ? 0 : 1; a.layer = l; a.seqmode = seqmode; a.nq = nq; a.ctx_end = ctx_end;
    a.qn = qn; a.kn = kn;
    prog.push_back(a);
    prog.push_back(gemv(wo, S.attn, S.nh * HD, nullptr, 0.f, hbuf, Hh, EPI_RESID, ncmode));
    prog.push_back(gemv(wgu, hbuf, Hh, ln2, S.eps, S.act, S.inter, EPI_SWIGLU, ncmode));
    prog.push_back(gemv(wd, S.act, S.inter, nullptr, 0.f, hbuf, Hh, EPI_RESID, ncmode));
  }
  return 0;
}

static int build_programs(q3_engine* e, int B) {
  if (e->prog_B == B) return 0;
  const q3_engine_cfg& c = e->cfg;
  StackDev &T = e->talker, &C = e->cp;
  const int H = T.hidden, Hc = C.hidden, G = c.num_code_groups;
  const bf16 *tnorm, *cnorm;
  PackedW whead;
  if (need_plain(e, "talker.norm", H, &tnorm) || need_plain(e, "cp.norm", Hc, &cnorm) ||
      need_packed(e, "talker.codec_head", T.vocab, H, &whead))
    return 1;
  PackedW wproj{};
  const bf16* bproj = nullptr;
  if (c.has_cp_projection) {
    if (need_packed(e, "cp.proj", Hc, H, &wproj) || need_plain(e, "cp.proj_bias", Hc, &bproj)) return 1;
  }
  // ---- prefill-chunk program: talker layers over the chunk's columns
  e->prog_layers.clear();  // (prefill runs on the tcgen05 GEMM path, see q3_prefill)
  // ---- prefill-head program: final norm + codec_head on the last token of each row, sample frame-0 codebook-0
  e->prog_head.clear();
  {
    Phase hd = gemv(whead, e->h_last, H, tnorm, T.eps, T.logits, T.vocab, EPI_LOGITS, NC_B, nullptr, e->past_hidden);
    e->prog_head.push_back(hd);
    Phase s{}; s.type = PH_SAMPLE; s.group = 0;
    e->prog_head.push_back(s);
  }
  // ---- frame program
  std::vector<Phase>& F = e->prog_frame;
  F.clear();
  const bool joint = (2 * B <= MAXCOLS);
  // The CP input x lives token-major as [2][B][H]: block 0 = past_hidden (pass 0) or the previous codebook's
  // embedding (passes >= 1), block 1 = E0[c0] (pass 0 only).  With a projection (1.7B) x is e->x_cp and a
  // bias-GEMV writes the projected rows into cp.h; with the Identity projection (0.6B, H == Hc) x IS cp.h
  // (e->x_cp_ptr() aliases it) and the layers run in place on the block they need.
  bf16* xin = c.has_cp_projection ? e->x_cp : C.h;
  auto cp_pass = [&](int ncmode, int nq, int ctx_end, int block, bool from_table = false) -> bf16* {
    bf16* hb = C.h;
    if (from_table) {
      // the previous sample phase already wrote the projected embedding rows into cp.h
    } else if (c.has_cp_projection) {
      F.push_back(gemv(wproj, xin + (size_t)block * B * H, H, nullptr, 0.f, C.h, Hc, EPI_BIAS, ncmode, bproj));
    } else {
      hb = C.h + (size_t)block * B * Hc;
    }
    if (add_layers(e, F, "cp", C, hb, ncmode, SEQ_CP, nq, ctx_end)) return nullptr;
    return hb;
  };
  for (int j = 0; j < G - 1; ++j) {
    PackedW wlm;
    if (need_packed(e, "cp.lm_head." + std::to_string(j), C.vocab, Hc, &wlm)) return 1;
    bf16* hb;
    if (j == 0) {
      if (joint) {
        if (!(hb = cp_pass(NC_2B, 2, 2, 0))) return 1;
        hb += (size_t)B * Hc;  // logits come from the second token's rows
      } else {
        if (!cp_pass(NC_B, 1, 1, 0)) return 1;
        if (!(hb = cp_pass(NC_B, 1, 2, 1))) return 1;
      }
    } else {
      if (!(hb = cp_pass(NC_B, 1, j + 2, 0, e->proj_tab != nullptr))) return 1;
    }
    F.push_back(gemv(wlm, hb, Hc, cnorm, C.eps, C.logits, C.vocab, EPI_LOGITS, NC_B));
    Phase s{}; s.type = PH_SAMPLE; s.group = j + 1;
    F.push_back(s);
  }
  if (add_layers(e, F, "talker", T, T.h, NC_B, SEQ_DECODE, 1, 0)) return 1;
  F.push_back(gemv(whead, T.h, H, tnorm, T.eps, T.logits, T.vocab, EPI_LOGITS, NC_B, nullptr, e->past_hidden));
  {
    Phase s{}; s.type = PH_SAMPLE; s.group = 0;
    F.push_back(s);
  }
  for (std::vector<Phase>* pv : {&e->prog_layers, &e->prog_head, &e->prog_frame})
    for (Phase& p : *pv)
      if (p.type == PH_GEMV) { p.tq = p.n_tiles / e->sm_count; p.tr = p.n_tiles % e->sm_count; }
  // upload
  const size_t total = e->prog_layers.size() + e->prog_head.size() + F.size();
  if ((int)total > e->prog_cap) {
    if (e->alloc(&e->prog_dev, total + 64)) return 1;
    e->prog_cap = (int)total + 64;
  }
  e->off_layers = 0;
  e->off_head = (int)e->prog_layers.size();
  e->off_frame = e->off_head + (int)e->prog_head.size();
  Q3_CUDA(cudaMemcpy(e->prog_dev + e->off_layers, e->prog_layers.data(), e->prog_layers.size() * sizeof(Phase), cudaMemcpyHostToDevice));
  Q3_CUDA(cudaMemcpy(e->prog_dev + e->off_head, e->prog_head.data(), e->prog_head.size() * sizeof(Phase), cudaMemcpyHostToDevice));
  Q3_CUDA(cudaMemcpy(e->prog_dev + e->off_frame, F.data(), F.size() * sizeof(Phase), cudaMemcpyHostToDevice));
  e->prog_B = B;
  return 0;
}

extern "C" int q3_engine_finalize(q3_engine* e) {
  Q3_REQUIRE(e, "null engine");
  Q3_CUDA(cudaSetDevice(e->cfg.device));
  const bf16* p;
  const q3_engine_cfg& c = e->cfg;
  if (need_plain(e, "talker.codec_embedding", (int64_t)c.talker.vocab_size * c.talker.hidden_size, &p)) return 1;
  if (need_plain(e, "talker.rope_cos", (int64_t)c.max_ctx * 64, &p)) return 1;
  e->talker.rope_cos = p;
  if (need_plain(e, "talker.rope_sin", (int64_t)c.max_ctx * 64, &p)) return 1;
  e->talker.rope_sin = p;
  if (need_plain(e, "cp.rope_cos", (int64_t)32 * 64, &p)) return 1;
  e->cp.rope_cos = p;
  if (need_plain(e, "cp.rope_sin", (int64_t)32 * 64, &p)) return 1;
  e->cp.rope_sin = p;
  if (need_plain(e, "cp.codec_embedding",
                 (int64_t)(c.num_code_groups - 1) * c.cp.vocab_size * c.talker.hidden_size, &p))
    return 1;
  // projection table: passes >= 1 of the code predictor feed small_to_mtp_projection(codec_embedding[j-1](c_j)) (:1281-1283),
  // a function of the sampled id only -> one tcgen05 GEMM over all (G-1)*Vc embedding rows here, no GEMV phase per pass
  if (c.has_cp_projection && e->use_proj_tab && e->gemm_w.count("cp.proj") && !e->proj_tab) {
    const int H = c.talker.hidden_size, Hc = c.cp.hidden_size;
    const int rows = (c.num_code_groups - 1) * c.cp.vocab_size;
    const bf16* bproj;
    if (need_plain(e, "cp.proj_bias", Hc, &bproj)) return 1;
    float* bias_f = nullptr;
    if (e->alloc(&bias_f, (size_t)Hc) || e->alloc(&e->proj_tab, (size_t)rows * Hc)) return 1;
    bf16_to_f32_kernel<<<(Hc + 255) / 256, 256, 0, e->copy_stream>>>(bproj, bias_f, Hc);
    GemmEpilogue ep{};
    ep.bias = bias_f; ep.cmod = Hc; ep.out_raw = e->proj_tab;
    GemmPlan plan;
    const int zero = 0;
    if (gemm_make_plan(&plan, p, 1, rows, H, H, (int64_t)rows * H, e->gemm_w["cp.proj"], Hc, H, 1, &zero,
                       gemm_pick_bn(Hc, (rows + 127) / 128, 1), ep))
      return 1;
    if (gemm_launch(plan, e->copy_stream)) return 1;
    Q3_CUDA(cudaStreamSynchronize(e->copy_stream));
  }
  if (build_programs(e, 1)) return 1;
  // kernel attributes
  Q3_CUDA(cudaFuncSetAttribute(q3_program_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_nt(1)));
  Q3_CUDA(cudaFuncSetAttribute(q3_program_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_nt(2)));
  Q3_CUDA(cudaFuncSetAttribute(q3_program_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_nt(4)));
  // algorithmic bytes
  auto lw = [](const q3_stack_cfg& s) {
    return (double)s.hidden_size * (s.num_heads * HD) * 2 + 2.0 * s.hidden_size * (s.num_kv_heads * HD) +
           3.0 * s.hidden_size * s.intermediate_size + 2.0 * s.hidden_size + 2.0 * HD;
  };
  const double wt = c.talker.num_layers * lw(c.talker) + c.talker.hidden_size + (double)c.talker.vocab_size * c.talker.hidden_size;
  const double proj = c.has_cp_projection ? ((double)c.talker.hidden_size * c.cp.hidden_size + c.cp.hidden_size) : 0.0;
  const double heads = (double)(c.num_code_groups - 1) * c.cp.hidden_size * c.cp.vocab_size;
  const double cpl = c.cp.num_layers * lw(c.cp) + c.cp.hidden_size;
  e->w_talker_bytes = 2.0 * wt;
  e->w_cp_unique_bytes = 2.0 * (cpl + proj + heads);
  e->w_cp_stream_bytes = 2.0 * ((c.num_code_groups - 1) * (cpl + proj) + heads);
  e->finalized = true;
  return 0;
}

static int launch_program(q3_engine* e, int off, int n, int mode, int max_iters, int nt,
                          int* codes_dev, cudaStream_t stream) {
  KParams P{};
  P.prog = e->prog_dev + off; P.n_phases = n; P.mode = mode; P.max_iters = max_iters; P.st = e->st;
  P.talker = e->talker; P.cp = e->cp; P.G = e->cfg.num_code_groups; P.eos = e->cfg.codec_eos_token_id;
  P.has_proj = e->cfg.has_cp_projection; P.sp = e->sp;
  P.B = e->B;
  for (int b = 0; b < MAXB; ++b) { P.len0[b] = e->len0[b]; P.trailing_len[b] = e->trailing_len[b]; }
  P.emb_t = e->plain["talker.codec_embedding"]; P.emb_cp = e->plain["cp.codec_embedding"];
  P.x_cp = e->cfg.has_cp_projection ? e->x_cp : e->cp.h; P.past_hidden = e->past_hidden; P.trailing = e->trailing; P.trailing_stride = e->trailing_cap;
  if (e->proj_tab) { P.cp_next = e->proj_tab; P.cp_next_dst = e->cp.h; P.cp_next_w = e->cfg.cp.hidden_size; }
  else { P.cp_next = P.emb_cp; P.cp_next_dst = P.x_cp; P.cp_next_w = e->cfg.talker.hidden_size; }
  P.tts_pad = e->tts_pad; P.seen = e->seen; P.codes_out = codes_dev; P.codes_stride = e->codes_stride;
  P.split_buf = e->split_buf; P.forced = e->forced; P.n_forced = e->n_forced; P.dbg_tlogits = e->dbg_t; P.dbg_clogits = e->dbg_c; P.prof = (mode == 1) ? e->prof : nullptr; P.dbg_skip = e->dbg_skip;
  Q3_CUDA(cudaMemsetAsync(&e->st->bar_count, 0, sizeof(unsigned int), stream));
  void* args[] = {&P};
  const void* fn = nt == 1 ? (const void*)q3_program_kernel<1> : nt == 2 ? (const void*)q3_program_kernel<2>
                                                                        : (const void*)q3_program_kernel<4>;
  Q3_CUDA(cudaLaunchCooperativeKernel(fn, dim3(e->sm_count), dim3(NTHREADS), args, (size_t)smem_bytes_nt(nt), stream));
  return 0;
}

extern "C" int q3_prefill(q3_engine* e, int32_t B, const void* embeds_dev, const int32_t* lens_host,
                          const void* trailing_dev, const int32_t* trailing_lens_host, int32_t trailing_stride,
                          const void* tts_pad_dev, const q3_sampling* sp, void* stream_) {
  Q3_REQUIRE(e && e->finalized, "engine not finalized");
  Q3_REQUIRE(B >= 1 && B <= e->cfg.max_batch, "batch %d out of range", B);
  Q3_REQUIRE(embeds_dev && lens_host && tts_pad_dev && sp, "null argument");
  Q3_CUDA(cudaSetDevice(e->cfg.device));
  cudaStream_t stream = (cudaStream_t)stream_;
  if (build_programs(e, B)) return 1;
  e->B = B; e->sp = *sp;
  const int H = e->cfg.talker.hidden_size;
  // per-request state reset
  DevState hs;
  memset(&hs, 0, sizeof(hs));
  hs.B = B;
  for (int b = 0; b < B; ++b) {
    Q3_REQUIRE(lens_host[b] >= 1 && lens_host[b] < e->cfg.max_ctx, "prompt length %d out of range", lens_host[b]);
    hs.len0[b] = lens_host[b];
    e->len0[b] = lens_host[b];
    e->max_len0 = b == 0 ? lens_host[b] : std::max(e->max_len0, lens_host[b]);
    hs.trailing_len[b] = trailing_lens_host ? trailing_lens_host[b] : 0;
    e->trailing_len[b] = hs.trailing_len[b];
    Q3_REQUIRE(hs.trailing_len[b] <= trailing_stride, "trailing length exceeds stride");
  }
  Q3_CUDA(cudaMemcpyAsync(e->st, &hs, sizeof(hs), cudaMemcpyHostToDevice, stream));
  Q3_CUDA(cudaMemsetAsync(e->seen, 0, (size_t)MAXB * e->cfg.talker.vocab_size, stream));
  Q3_CUDA(cudaMemcpyAsync(e->tts_pad, tts_pad_dev, (size_t)H * 2, cudaMemcpyDeviceToDevice, stream));
  if (trailing_stride > 0 && trailing_dev) {
    const size_t need = (size_t)MAXB * trailing_stride * H;
    if (need > e->trailing_alloc) {  // (re)allocate the engine-owned copy of trailing_text_hidden
      if (e->alloc(&e->trailing, need)) return 1;
      e->trailing_alloc = need;
    }
    Q3_CUDA(cudaMemcpyAsync(e->trailing, trailing_dev, (size_t)B * trailing_stride * H * 2, cudaMemcpyDeviceToDevice, stream));
  }
  e->trailing_cap = trailing_stride;
  e->codes_stride = 0;
  e->frames_issued = 0;
  // ---- prefill on tensor cores: all prompt tokens of all rows at once (packed [ntok][H], rows back to back)
  int ntok = 0;
  std::vector<int> hseq, hpos, last(B);
  for (int b = 0; b < B; ++b) {
    for (int p = 0; p < lens_host[b]; ++p) { hseq.push_back(b); hpos.push_back(p); }
    ntok += lens_host[b];
    last[b] = ntok - 1;
  }
  const q3_stack_cfg& tc = e->cfg.talker;
  const int nh = tc.num_heads, nkv = tc.num_kv_heads, I = tc.intermediate_size, QKV = (nh + 2 * nkv) * HD;
  if (ntok > e->pf_cap) {
    const int cap = std::max(ntok, 1024);
    if (e->alloc(&e->pf_x, (size_t)cap * H) || e->alloc(&e->pf_xn, (size_t)cap * H) || e->alloc(&e->pf_qkv, (size_t)cap * QKV) ||
        e->alloc(&e->pf_attn, (size_t)cap * nh * HD) || e->alloc(&e->pf_act, (size_t)cap * I) || e->alloc(&e->pf_seq, (size_t)cap) ||
        e->alloc(&e->pf_pos, (size_t)cap))
      return 1;
    e->pf_cap = cap;
  }
  Q3_CUDA(cudaMemcpyAsync(e->pf_seq, hseq.data(), (size_t)ntok * sizeof(int), cudaMemcpyHostToDevice, stream));
  Q3_CUDA(cudaMemcpyAsync(e->pf_pos, hpos.data(), (size_t)ntok * sizeof(int), cudaMemcpyHostToDevice, stream));
  Q3_CUDA(cudaStreamSynchronize(stream));  // hseq/hpos are stack-owned pageable buffers
  Q3_CUDA(cudaMemcpyAsync(e->pf_x, embeds_dev, (size_t)ntok * H * 2, cudaMemcpyDeviceToDevice, stream));
  const int mt = (ntok + 127) / 128;
  const int zero = 0;
  auto gemm = [&](const bf16* a, int K, const std::string& wname, int N, GemmEpilogue ep) -> int {
    auto it = e->gemm_w.find(wname);
    Q3_REQUIRE(it != e->gemm_w.end(), "missing prefill GEMM weight %s", wname.c_str());
    ep.cmod = N;
    GemmPlan plan;
    if (gemm_make_plan(&plan, a, 1, ntok, K, K, (int64_t)ntok * K, it->second, N, K, 1, &zero, gemm_pick_bn(N, mt, 1), ep)) return 1;
    return gemm_launch(plan, stream);
  };
  const int rows_per_blk = 8;
  for (int l = 0; l < tc.num_layers; ++l) {
    const std::string p = "talker.layers." + std::to_string(l);
    const bf16 *ln1 = e->plain[p + ".ln1"], *ln2 = e->plain[p + ".ln2"], *qn = e->plain[p + ".q_norm"], *kn = e->plain[p + ".k_norm"];
    Q3_REQUIRE(ln1 && ln2 && qn && kn, "missing norm weights of %s", p.c_str());
    pf_rmsnorm_kernel<<<(ntok + rows_per_blk - 1) / rows_per_blk, 256, 0, stream>>>(e->pf_x, ln1, e->pf_xn, ntok, H, tc.rms_eps);
    { GemmEpilogue ep{}; ep.out_raw = e->pf_qkv; if (gemm(e->pf_xn, H, p + ".qkv", QKV, ep)) return 1; }
    pf_qkv_post_kernel<<<(ntok * (nh + 2 * nkv) + 7) / 8, 256, 0, stream>>>(e->pf_qkv, ntok, e->pf_seq, e->pf_pos, nh, nkv, qn, kn, tc.rms_eps,
                                                                         e->talker.rope_cos, e->talker.rope_sin, e->talker.kc,
                                                                         e->talker.vc, l, tc.num_layers, e->talker.cap);
    pf_attention_kernel<<<(ntok * nh + 7) / 8, 256, 0, stream>>>(e->pf_qkv, e->pf_attn, ntok, e->pf_seq, e->pf_pos, nh, nkv, e->talker.kc,
                                                                 e->talker.vc, l, tc.num_layers, e->talker.cap);
    { GemmEpilogue ep{}; ep.resid = e->pf_x; ep.out_raw = e->pf_x; if (gemm(e->pf_attn, nh * HD, p + ".o", H, ep)) return 1; }
    pf_rmsnorm_kernel<<<(ntok + rows_per_blk - 1) / rows_per_blk, 256, 0, stream>>>(e->pf_x, ln2, e->pf_xn, ntok, H, tc.rms_eps);
    { GemmEpilogue ep{}; ep.act = ACT_SWIGLU_BLK8; ep.out_act = e->pf_act; if (gemm(e->pf_xn, H, p + ".gate_up", 2 * I, ep)) return 1; }
    { GemmEpilogue ep{}; ep.resid = e->pf_x; ep.out_raw = e->pf_x; if (gemm(e->pf_act, I, p + ".down", H, ep)) return 1; }
  }
  Q3_CUDA(cudaGetLastError());
  for (int b = 0; b < B; ++b)
    Q3_CUDA(cudaMemcpyAsync(e->h_last + (size_t)b * H, e->pf_x + (size_t)last[b] * H, (size_t)H * 2, cudaMemcpyDeviceToDevice, stream));
  // head + sample codebook-0 of frame 0 (codes are materialised by q3_decode's first call via st->c0)
  const int nt_head = B <= 8 ? 1 : B <= 16 ? 2 : 4;
  if (launch_program(e, e->off_head, (int)e->prog_head.size(), 0, 1, nt_head, nullptr, stream)) return 1;
  return 0;
}

extern "C" int q3_decode(q3_engine* e, int32_t max_frames, int32_t* codes_dev, int32_t codes_stride, void* stream_) {
  Q3_REQUIRE(e && e->finalized && e->B > 0, "prefill first");
  Q3_REQUIRE(codes_dev && codes_stride > 0 && max_frames > 0, "bad arguments");
  Q3_CUDA(cudaSetDevice(e->cfg.device));
  cudaStream_t stream = (cudaStream_t)stream_;
  e->codes_stride = codes_stride;
  Q3_REQUIRE(e->max_len0 + e->frames_issued + max_frames <= e->cfg.max_ctx, "KV capacity exceeded: prompt %d + %d frames > max_ctx %d",
             e->max_len0, e->frames_issued + max_frames, e->cfg.max_ctx);
  e->frames_issued += max_frames;
  const int B = e->B;
  const int cols = (2 * B <= MAXCOLS) ? 2 * B : B;
  const int nt = cols <= 8 ? 1 : cols <= 16 ? 2 : 4;
  return launch_program(e, e->off_frame, (int)e->prog_frame.size(), 1, max_frames, nt, codes_dev, stream);
}

extern "C" int q3_get_progress(q3_engine* e, int32_t* frames_done, int32_t* n_valid, int32_t* finished) {
  Q3_REQUIRE(e, "null engine");
  Q3_CUDA(cudaSetDevice(e->cfg.device));
  DevState hs;
  Q3_CUDA(cudaMemcpy(&hs, e->st, sizeof(hs), cudaMemcpyDeviceToHost));
  Q3_REQUIRE(hs.error == 0, "device-side error %d (grid barrier timeout)", hs.error);
  if (frames_done) *frames_done = hs.step;
  for (int b = 0; b < e->B; ++b) {
    if (n_valid) n_valid[b] = hs.finished[b] ? hs.n_valid[b] : hs.step;
    if (finished) finished[b] = hs.finished[b];
  }
  return 0;
}

extern "C" int q3_set_debug(q3_engine* e, const int32_t* forced_dev, int32_t n_frames, float* talker_logits_dev,
                            float* cp_logits_dev) {
  Q3_REQUIRE(e, "null engine");
  e->forced = forced_dev; e->n_forced = forced_dev ? n_frames : 0;
  e->dbg_t = talker_logits_dev; e->dbg_c = cp_logits_dev;
  return 0;
}

extern "C" int q3_set_profile(q3_engine* e, unsigned long long* prof_dev) {
  Q3_REQUIRE(e, "null engine");
  e->prof = prof_dev;
  return 0;
}

// phase kinds of the frame program (0 GEMV, 1 ATTN, 2 SAMPLE) + stack (0 talker, 1 cp) + epilogue, for profiling
extern "C" int q3_describe_frame_program(q3_engine* e, int32_t* kinds, int32_t capacity) {
  Q3_REQUIRE(e && e->prog_B > 0, "no program built yet");
  const int n = (int)e->prog_frame.size();
  if (kinds) {
    for (int i = 0; i < n && i < capacity; ++i) {
      const Phase& p = e->prog_frame[i];
      int stack = p.stack;
      if (p.type == PH_GEMV) stack = (p.kb * 32 == e->cfg.talker.hidden_size || p.kb * 32 == e->cfg.talker.intermediate_size ||
                                      p.kb * 32 == e->cfg.talker.num_heads * HD) && !(p.src == e->cp.h || p.src == e->cp.attn || p.src == e->cp.act) ? 0 : 1;
      if (p.type == PH_SAMPLE) stack = p.group == 0 ? 0 : 1;
      kinds[i] = p.type * 100 + stack * 10 + (p.type == PH_GEMV ? p.epi : 0);
    }
  }
  return -n;  // negative = count (0 is reserved for success elsewhere); callers use abs()
}

// Debug/profiling: run a synthetic program made of `count` repetitions of the frame-program phases
// [first, first+span) (mode 0, one pass) and return the elapsed device time.  Used by tools/icache_probe.py.
extern "C" int q3_debug_set_skip(q3_engine* e, int32_t mask) { if (e) e->dbg_skip = mask; return 0; }

extern "C" int q3_debug_time_phases(q3_engine* e, int32_t first, int32_t span, int32_t count, float* ms_out, void* stream_) {
  Q3_REQUIRE(e && e->prog_B > 0 && e->B > 0, "prefill first");
  Q3_REQUIRE(first >= 0 && span >= 1 && first + span <= (int)e->prog_frame.size() && count >= 1, "bad phase range");
  cudaStream_t stream = (cudaStream_t)stream_;
  std::vector<Phase> prog;
  for (int i = 0; i < count; ++i)
    for (int j = 0; j < span; ++j) {
      Phase p = e->prog_frame[first + j];
      if (p.type == PH_SAMPLE) continue;
      prog.push_back(p);
    }
  const int n = (int)prog.size();
  Phase* dev = nullptr;
  Q3_CUDA(cudaMalloc(&dev, (size_t)n * sizeof(Phase)));
  Q3_CUDA(cudaMemcpy(dev, prog.data(), (size_t)n * sizeof(Phase), cudaMemcpyHostToDevice));
  Phase* saved = e->prog_dev;
  e->prog_dev = dev;
  const int B = e->B;
  const int cols = (2 * B <= MAXCOLS) ? 2 * B : B;
  const int nt = cols <= 8 ? 1 : cols <= 16 ? 2 : 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = launch_program(e, 0, n, 0, 1, nt, nullptr, stream);  // warm
  cudaEventRecord(e0, stream);
  if (!rc) rc = launch_program(e, 0, n, 0, 1, nt, nullptr, stream);
  cudaEventRecord(e1, stream);
  cudaStreamSynchronize(stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = ms;
  e->prog_dev = saved;
  cudaFree(dev);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

extern "C" int q3_algorithmic_bytes(q3_engine* e, int32_t B, int32_t S, double* a_bytes, double* a_stream_bytes) {
  Q3_REQUIRE(e && e->finalized, "engine not finalized");
  const q3_engine_cfg& c = e->cfg;
  const double kv_tok = (double)c.talker.num_layers * 2 * c.talker.num_kv_heads * HD * 2;
  const double kv = (double)B * (S + 1) * kv_tok;
  if (a_bytes) *a_bytes = e->w_talker_bytes + e->w_cp_unique_bytes + kv;
  if (a_stream_bytes) *a_stream_bytes = e->w_talker_bytes + e->w_cp_stream_bytes + kv;
  return 0;
}
