// Grid barrier, block helpers and the GEMV phase (RMSNorm-fused activation staging, batch-in-N mma.sync, fused epilogues).
// Part of the ar_engine.cu translation unit (include order: ar_program, ar_gemv, ar_attention, ar_sampler,
// the persistent kernel in ar_engine.cu, ar_prefill).
#pragma once

namespace {

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

}  // namespace
