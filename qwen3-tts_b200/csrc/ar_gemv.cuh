// Grid barrier, mbarrier / TMA helpers, the per-warp weight ring and the GEMV phase (RMSNorm-fused activation staging,
// batch-in-N mma.sync fed from shared memory, fused epilogues).
// Part of the ar_engine.cu translation unit (include order: ar_program, ar_gemv, ar_attention, ar_sampler,
// the persistent kernel in ar_engine.cu, ar_prefill).
#pragma once

namespace {

using q3ring::PMeta;
using q3ring::ProdIter;
using q3ring::RunGeom;

// ------------------------------------------------------------------------------------------------
// grid barrier.  Two interchangeable implementations (KParams.flags bit 0):
//   counter (default): release-red on one counter + relaxed poll by thread 0 (arrive -> release ~1.0 us measured).
//   flags            : every CTA owns one epoch word; arrive = st.release of the new epoch, wait = warp 0 polls all
//                      gridDim words.  Measured SLOWER on B200 (arrive -> release 3-4 us: 148 pollers x 5 lines), kept
//                      as an A/B knob only (profiles/r02_barrier_ab.txt).
// Polls are RELAXED on purpose: ld.acquire.gpu compiles to LDG.STRONG + CCTL.IVALL (a full L1 invalidation per
// poll iteration).  Correctness does not need it: every cross-CTA read in this kernel is an L2 access
// (ld.global.cg / cp.async.cg / cp.async.bulk), the writers released at gpu scope before their arrival became
// visible, and the GPU does not speculate loads past the poll loop + bar.sync.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void barrier_timeout(DevState* st) {
  st->error = 77;
  __threadfence();
  __trap();
}

__device__ __forceinline__ void grid_arrive(DevState* st, unsigned int& epoch, bool use_counter) {
  // caller: cta_sync() already executed (every thread's global writes happen-before thread 0's release)
  if (threadIdx.x == 0) {
    if (use_counter) {
      epoch += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&st->bar_count) : "memory");
    } else {
      epoch += 1;
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&st->bar_flags[blockIdx.x]), "r"(epoch) : "memory");
    }
  }
}

__device__ __forceinline__ void grid_wait(DevState* st, unsigned int epoch, bool use_counter) {
  if (use_counter) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      unsigned int v;
      while (true) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&st->bar_count) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > 8000000000LL) barrier_timeout(st);  // ~4 s: never hang the box
      }
    }
  } else if (threadIdx.x < 32) {
    const long long t0 = clock64();
    const int n = (int)gridDim.x;
    // thread 0 carries the epoch; broadcast it inside warp 0
    const unsigned int ep = __shfl_sync(0xffffffffu, epoch, 0);
    while (true) {
      bool ok = true;
#pragma unroll 1
      for (int i = threadIdx.x; i < n; i += 32) {
        unsigned int v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&st->bar_flags[i]) : "memory");
        ok = ok && ((int)(v - ep) >= 0);
      }
      if (__all_sync(0xffffffffu, ok)) break;
      if (clock64() - t0 > 8000000000LL) barrier_timeout(st);
    }
  }
}

// CTA-wide barrier of the phase code
__device__ __forceinline__ void cta_sync() { __syncthreads(); }

// fine-grained profiling marks (thread 0 only, first frame of a profiled launch)
__shared__ unsigned long long* g_prof_row;
// (compiled only into the development instantiation: a constexpr bool DEV must be in scope)
#define PROF_MARK(k)                                                                   \
  do {                                                                                 \
    if constexpr (DEV) if (threadIdx.x == 0 && g_prof_row) {                           \
      unsigned long long _t;                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                          \
      g_prof_row[k] = _t;                                                              \
    }                                                                                  \
  } while (0)

// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) primitives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: a waiting thread sleeps in hardware (woken by the phase completion) instead of
// spinning through issue slots that the other warps of its scheduler need
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(2000u)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, DevState* st) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {  // protocol bug: report instead of hanging the box
      st->error = 78;
      __threadfence();
      __trap();
    }
  }
}
// global -> shared bulk copy through the TMA unit; bytes % 16 == 0, both addresses 16-byte aligned.
// `policy` is an L2 eviction-priority descriptor (createpolicy): weights that are re-read within a frame
// (code predictor) are kept, the once-per-frame talker stream is marked evict-first.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// ------------------------------------------------------------------------------------------------
// weight rings.  Every warp owns a private ring of R slots (SB KB each) in shared memory and is its own TMA producer:
// when it has consumed a slot its lane 0 immediately refills it with the piece R positions ahead in the warp's
// deterministic piece sequence — across phases and grid barriers, so weight streaming never waits for the activations'
// dependency chain.  The sequence itself is a TABLE built once on the host with the iterator of ar_ring.cuh
// (build_run_table): per CTA and warp the (address, length) of every contiguous run of one program pass; pieces are
// the <= SB-block chunks of a run.  A refill costs one expect_tx and one cp.async.bulk; the next run's descriptor is
// loaded a whole run ahead, so nothing is computed and no load latency is exposed on the device.
// (Measured alternatives, profiles/r02_phase_breakdown.txt: a device-side iterator inside the consumer loop cost
// ~0.3 us per piece, a dedicated producer warp ~0.7 us per piece of serialised iterator work.)
// ------------------------------------------------------------------------------------------------
struct Ring {
  uint32_t slots;          // shared address of the warp's first slot
  uint32_t full;           // shared address of its first full mbarrier
  int SB, R;               // blocks per slot, slots
  int c_slot, c_par;       // consumer position / parity of the current lap
  int p_slot;              // next slot to fill
  const uint2* list;       // this warp's RUN list of one program pass: (offset from the weight base / 16, 1 KB blocks)
  int len;                 // runs per pass
  int ri;                  // index in `list` of the run held in `nxt`
  long long runs_left;     // runs of this launch not yet started (including the one in `nxt`)
  uint32_t cur_off;        // current run: next block's offset / 16
  int cur_left;            // blocks of the current run not yet requested
  bool cur_keep;           // current run belongs to weights that are re-read within a frame (code-predictor layers)
  uint2 nxt;               // the following run (prefetched a whole run ahead: its L2 latency is never exposed)
  int outstanding;         // pieces requested, not yet consumed
};

__device__ __forceinline__ void ring_init(Ring& rg, const uint2* list, int len, long long total_runs) {
  rg.list = list; rg.len = len; rg.ri = 0; rg.runs_left = total_runs;
  rg.cur_off = 0; rg.cur_left = 0; rg.cur_keep = false;
  rg.nxt = len > 0 ? __ldg(list) : make_uint2(0, 0);
}

// request the next piece (<= SB blocks of the current run) into slot p_slot; lane 0 issues, every lane keeps the books
template <bool DEV>
__device__ __forceinline__ void ring_issue(Ring& rg, const KParams& P, int lane, uint64_t pol_stream, uint64_t pol_keep) {
  if (rg.cur_left == 0) {
    if (rg.runs_left <= 0) return;
    rg.cur_off = rg.nxt.x; rg.cur_left = (int)(rg.nxt.y & 0x7fffffffu); rg.cur_keep = (rg.nxt.y >> 31) != 0;
    --rg.runs_left;
    if (++rg.ri == rg.len) rg.ri = 0;
    rg.nxt = __ldg(rg.list + rg.ri);
  }
  const int nb = q3ring::imin(rg.SB, rg.cur_left);
  if (lane == 0) {
    const uint32_t bytes = (uint32_t)nb << 10;
    const uint32_t bar = rg.full + 8u * rg.p_slot;
    mbar_expect_tx(bar, bytes);
    const char* src = P.wbase + ((size_t)rg.cur_off << 4);
    // L2 priority of the line fill: pol_stream = evict_first (read once per frame-step), pol_keep = a fixed,
    // address-hashed fraction evict_last (the code predictor's layer weights are re-read on each of its 15 passes:
    // what stays in L2 is not re-fetched from HBM; measured traffic: profiles/r02_l2_residency.txt)
    const uint64_t pol = rg.cur_keep ? pol_keep : pol_stream;
    if (DEV && (P.flags & 4)) bulk_g2s_plain(rg.slots + (uint32_t)(rg.p_slot * rg.SB) * 1024u, src, bytes, bar);
    else bulk_g2s(rg.slots + (uint32_t)(rg.p_slot * rg.SB) * 1024u, src, bytes, bar, pol);
  }
  rg.cur_off += (uint32_t)nb << 6;
  rg.cur_left -= nb;
  if (++rg.p_slot == rg.R) rg.p_slot = 0;
  ++rg.outstanding;
}

template <bool DEV>
__device__ __forceinline__ void ring_release(Ring& rg, const KParams& P, int lane, uint64_t pol_stream, uint64_t pol_keep) {
  __syncwarp();  // every lane's reads of the slot are done before the async proxy overwrites it
  if (++rg.c_slot == rg.R) { rg.c_slot = 0; rg.c_par ^= 1; }
  --rg.outstanding;
  ring_issue<DEV>(rg, P, lane, pol_stream, pol_keep);
}

// ------------------------------------------------------------------------------------------------
// block-wide helpers (NTHREADS threads)
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ float block_reduce(float v, float* red, int op /*0 max, 1 sum*/) {
  if (op == 0) v = warp_max(v); else v = warp_sum(v);
  cta_sync();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  cta_sync();
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
  cta_sync();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  cta_sync();
  int r = red[threadIdx.x & (NWARPS - 1)];
#pragma unroll
  for (int o = NWARPS / 2; o > 0; o >>= 1) r = min(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}

// ------------------------------------------------------------------------------------------------
// GEMV phase:  dst[col][row] = epi( sum_k W[row][k] * x[col][k] )
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int xs_stride_bytes(int K) { return K * 2 + 64; }  // +64 B: conflict-free B-fragment reads

__device__ __forceinline__ int phase_nc(int ncmode, const KParams& P) {
  const int B = P.B;
  return ncmode == NC_B ? B : 2 * B;
}

__device__ __forceinline__ uint32_t norm_pair(uint32_t x2, uint32_t w2, float inv) {
  __nv_bfloat162 t = __floats2bfloat162_rn(bf16lo(x2) * inv, bf16hi(x2) * inv);
  __nv_bfloat162 r = __hmul2(t, *reinterpret_cast<const __nv_bfloat162*>(&w2));
  return *reinterpret_cast<uint32_t*>(&r);
}

// stage RMS-normed x into smem as bf16 [col][K] (rows skewed by 64 B).  One warp per column; a lane issues its (up to 8)
// independent 16-byte loads first, the RMSNorm runs on the registers (sum of squares -> warp reduce -> scale) and the
// result is written to smem once.  The norm weights sit in nw_s (fetched before the preceding grid barrier).
// (Un-normed inputs are copied by the TMA unit, see gemv_phase.)  K <= 2048.
__device__ __forceinline__ void stage_columns(const bf16* __restrict__ src, int src_ld, float eps, bf16* __restrict__ save, unsigned int save_mask,
                                              int K, int nc, char* __restrict__ xs, int xstride, const uint4* __restrict__ nw_s,
                                              bf16* __restrict__ hid, int hid_stride, int hid_step, const int* __restrict__ frame0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = K >> 3;
#pragma unroll 1
  for (int col = warp; col < nc; col += NWARPS) {
    const uint4* xr = reinterpret_cast<const uint4*>(src + (size_t)col * src_ld);
    uint4* drow = reinterpret_cast<uint4*>(xs + (size_t)col * xstride);
    // per-step hidden-state capture: row `col`, its own step index (prefill = 0, frame f -> f + 1 - frame0)
    const int hstep = hid ? hid_step - frame0[col] : -1;
    uint4* hrow = (hid && save && ((save_mask >> col) & 1u) && hstep >= 0 && hstep < hid_stride)
                      ? reinterpret_cast<uint4*>(hid + ((size_t)col * hid_stride + hstep) * K) : nullptr;
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (lane + 32 * i < nv) v[i] = ldcg16(xr + lane + 32 * i);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane + 32 * i < nv) {
        const uint32_t w4[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float lo = bf16lo(w4[q]), hi = bf16hi(w4[q]); ss += lo * lo; ss += hi * hi; }
      }
    }
    ss = warp_sum(ss);
    const float inv = rsqrtf(ss / (float)K + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane + 32 * i < nv) {
        // bf16(bf16(x*inv) * w): fp32 scale, one packed RN conversion, then a packed bf16 multiply (HMUL2.BF16
        // rounds the exact product to nearest-even = the reference's bf16 x bf16 -> bf16 multiply)
        const uint4 w = nw_s[lane + 32 * i];
        uint4 o;
        o.x = norm_pair(v[i].x, w.x, inv);
        o.y = norm_pair(v[i].y, w.y, inv);
        o.z = norm_pair(v[i].z, w.z, inv);
        o.w = norm_pair(v[i].w, w.w, inv);
        if (save && ((save_mask >> col) & 1u)) reinterpret_cast<uint4*>(save + (size_t)col * K)[lane + 32 * i] = o;
        if (hrow) hrow[lane + 32 * i] = o;
        drow[lane + 32 * i] = o;
      }
    }
  }
}

// per-CTA scratch of one GEMV round (shared memory): which warps hold partial sums of which tile
struct RoundTab {
  int tl0[NWARPS];   // first tile (round-local) of warp w's run
  int wf[8], wl[8];  // first / last warp contributing to round-local tile tl
};

__device__ __forceinline__ int idiv_small(int a, int b) {  // exact for 0 <= a < 2^20, 0 < b < 2^12
  return __float2int_rz(__fdividef((float)a + 0.5f, (float)b));
}

// One full ring piece (NBLK = 4 or 2 k32-blocks of one tile): A fragments from the ring slot, B fragments already in registers.
// A block is two 512-byte halves; lane l's 16 bytes of a half ARE the four A registers of one m16n8k16 MMA
// (pack_weight_kernel), so no register shuffling sits between the loads and the tensor pipe.
template <int NT, int NACC, int NBLK>
__device__ __forceinline__ void piece_full(float (&acc)[NACC][NT][4], uint32_t sp, const uint4 (&b)[(NT == 4 ? 2 : 4)][NT], int nct) {
  uint4 a1[NBLK], a2[NBLK];
#pragma unroll
  for (int i = 0; i < NBLK; ++i) { a1[i] = lds128(sp + i * 1024); a2[i] = lds128(sp + i * 1024 + 512); }
#pragma unroll
  for (int i = 0; i < NBLK; ++i)
#pragma unroll
    for (int n = 0; n < NT; ++n)
      if (n < nct) {
        mma_bf16_16816(acc[(2 * i) % NACC][n], a1[i].x, a1[i].y, a1[i].z, a1[i].w, b[i][n].x, b[i][n].y);
        mma_bf16_16816(acc[(2 * i + 1) % NACC][n], a2[i].x, a2[i].y, a2[i].z, a2[i].w, b[i][n].z, b[i][n].w);
      }
}

// B fragments of PB (4; 2 at the largest batch class, whose ring pieces are 2 blocks) consecutive k-blocks starting at block kb0 (wrapping at KB) straight from global memory.
// gsrc points at this lane's column g / k offset t*8; `rows_left` = nc - g (columns n*8+g beyond it read zero).
template <int NT>
__device__ __forceinline__ void load_bfrags(uint4 (&dst)[(NT == 4 ? 2 : 4)][NT], const bf16* __restrict__ gsrc, int src_ld, int rows_left,
                                            int nct, int kb0, int KB) {
#pragma unroll
  for (int i = 0; i < (NT == 4 ? 2 : 4); ++i) {
    int kk = kb0 + i;
    if (kk >= KB) kk -= KB;
#pragma unroll
    for (int n = 0; n < NT; ++n)
      dst[i][n] = (n < nct && n * 8 < rows_left) ? ldcg16(gsrc + (size_t)(n * 8) * src_ld + kk * 32) : make_uint4(0, 0, 0, 0);
  }
}

template <int NT, int NACC>
__device__ __forceinline__ void flush_acc(float (&acc)[NACC][NT][4], float* pp, int nct, int t, int g) {
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = acc[0][n][q];
#pragma unroll
      for (int a = 1; a < NACC; ++a) v[q] += acc[a][n][q];
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a][n][q] = 0.f;
    }
    if (n < nct) {
      const int col = n * 8 + 2 * t;
      pp[(col)*PCOL + g] = v[0];
      pp[(col + 1) * PCOL + g] = v[1];
      pp[(col)*PCOL + g + 8] = v[2];
      pp[(col + 1) * PCOL + g + 8] = v[3];
    }
  }
}

// The phase body.  `m` is this CTA's meta of the phase (tiles owned, K blocks); the weights arrive through the ring.
// HID: the kernel instantiation that also captures the per-step hidden states (generate(return_hidden=True)); the
// default instantiation does not carry that code — the phase loop's instruction footprint is a first-order term
// (measured: 4.22 -> 4.14 ms per frame-step at B=8 without it).  DEV: the development instantiation with the
// device-timestamp marks and the ablation / A-B knobs of KParams::flags; production kernels carry none of it.
template <int NT, bool HID, bool DEV>
__device__ __forceinline__ uint32_t gemv_phase(const Phase& ph, const PMeta m, const KParams& P, Ring& rg, unsigned char* smem,
                                               RoundTab* tab, uint32_t xbar, uint32_t xpar, uint64_t pol_stream, uint64_t pol_keep, int frame) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int KB = m.kb, K = KB * 32, epi = ph.epi, ntc = m.ntc;
  const bf16* const src = ph.src;
  const int src_ld = ph.src_ld, dst_ld = ph.dst_ld;
  const bool normed = ph.norm_w != nullptr;
  void* const dst = ph.dst;
  const int nc = phase_nc(ph.ncmode, P);
  const int nct = (nc + 7) >> 3;
  const bool staged = ph.staged != 0;
  int t0, ntc_chk;
  q3ring::cta_tiles(ph.tq, ph.tr, (int)blockIdx.x, t0, ntc_chk);
  bf16* const save = (ph.save_normed != nullptr && blockIdx.x == 0) ? ph.save_normed : nullptr;

  char* xs = reinterpret_cast<char*>(smem + P.plan.x_off);
  float* part = reinterpret_cast<float*>(smem + P.plan.part_off);  // [NWARPS][2][NT*8][PCOL]
  const uint4* nw_s = reinterpret_cast<const uint4*>(smem + P.plan.nw_off);
  const int xstride = xs_stride_bytes(K);
  const uint32_t xs_sh = smem_addr(xs);

  // ---- activations -> shared memory
  if (staged && (ntc > 0 || save) && !(DEV && (P.flags & 32))) {
    if (normed) {
      stage_columns(src, src_ld, ph.eps, save, P.mode == 0 ? P.admit_mask : 0xffffffffu, K, nc, xs, xstride, nw_s,
                    (HID && save) ? P.hid_out : nullptr, P.hid_stride, P.mode == 0 ? 0 : frame + 1, P.frame0);
    } else {
      // plain copy of nc contiguous rows: one bulk (TMA) copy per column, completion on the CTA's x barrier
      if (tid == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");  // earlier generic accesses of the x area / of src's producers
        mbar_expect_tx(xbar, (uint32_t)(nc * K * 2));
#pragma unroll 1
        for (int col = 0; col < nc; ++col)
          bulk_g2s_plain(xs_sh + (uint32_t)(col * xstride), src + (size_t)col * src_ld, (uint32_t)(K * 2), xbar);
      }
      mbar_wait(xbar, xpar, P.st);
      xpar ^= 1;
    }
  }
  PROF_MARK(2);
  if (ntc <= 0) return xpar;  // (the phase loop syncs the CTA right after the body)

  const bool swiglu = epi == EPI_SWIGLU;
  const int rsh = swiglu ? 3 : 4;  // output rows per tile: 8 (gate/up pairs) or 16
#pragma unroll 1
  for (int round = 0; round * 8 < ntc; ++round) {
    if (round) cta_sync();  // the previous round's epilogue still reads the round table and part
    const RunGeom rgm = q3ring::run_geom(ntc, KB, round, warp);
    const int TB = rgm.TB, upw = rgm.upw;
    const int tl0 = idiv_small(rgm.u0, KB);
    if (tid < NWARPS) tab->tl0[tid] = idiv_small(q3ring::imin(TB * KB, tid * upw), KB);
    else if (tid < NWARPS + 8 && tid - NWARPS < TB) {
      const int tl = tid - NWARPS;
      tab->wf[tl] = idiv_small(tl * KB, upw);
      tab->wl[tl] = idiv_small((tl + 1) * KB - 1, upw);
    }
    // residual values of this thread's first two output elements (own rows: stable since the previous barrier)
    int tbl = 0;
    while ((1 << tbl) < TB) ++tbl;
    const int nelem = nc << (rsh + tbl);
    float resid0 = 0.f, resid1 = 0.f;  // (two scalars, not an array: a dynamically indexed array would live in local memory)
    if (epi == EPI_RESID) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e = tid + i * NTHREADS;
        const int tl = (e >> rsh) & ((1 << tbl) - 1);
        if (e < nelem && tl < TB) {
          const int row = (t0 + round * 8 + tl) * 16 + (e & 15);
          const float v = bf2f(ldcg_bf16(reinterpret_cast<const bf16*>(dst) + (size_t)(e >> (rsh + tbl)) * dst_ld + row));
          if (i == 0) resid0 = v; else resid1 = v;
        }
      }
    }
    cta_sync();  // x area + round table visible (also separates rounds: part is free again)

    // ---- main loop: this warp's run, piece by piece, out of its ring.  All fragment loads of a piece are issued
    // before its MMAs, and consecutive MMAs go to independent accumulators (NACC sets), so neither the LDS latency
    // nor the MMA latency serialises (measured before: ~280 cycles per 1 KB block with the naive order)
    {
      constexpr int NACC = NT <= 2 ? 4 : 2;
      float acc[NACC][NT][4];
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[a][n][q] = 0.f;
      long long wait_cycles = 0;
      PROF_MARK(8);
      int u = rgm.u0;
      int kbi = rgm.u0 - tl0 * KB;
      int seg = 0;
      float* pp = part + ((warp * 2) * (NT * 8)) * PCOL;
      const uint32_t xb0 = xs_sh + (uint32_t)(g * xstride + t * 16);  // this lane's B-fragment base (column g of n-tile 0)
      // Un-staged inputs (K too large for the x area: the down projections): B fragments come straight from L2 into
      // registers, one piece AHEAD of the MMAs that use them, so only the first piece of a run exposes the L2 latency
      const bf16* gsrc = src + (size_t)g * src_ld + t * 8;
      constexpr int PB = NT == 4 ? 2 : 4;  // blocks per ring piece (the plan never gives the largest batch class more)
      uint4 bq[PB][NT], bnx[PB][NT];
      if (!staged) load_bfrags<NT>(bq, gsrc, src_ld, nc - g, nct, kbi, KB);
#pragma unroll 1
      while (u < rgm.u1) {
        const int nb = q3ring::imin(rg.SB, rgm.u1 - u);
        const bool full = nb == rg.SB && (nb == PB || nb == 2) && kbi + nb <= KB;  // every shipped shape: runs and tiles are multiples of 4 blocks
        if (!staged && u + nb < rgm.u1) {
          int kn = kbi + nb;
          if (kn >= KB) kn -= KB;
          load_bfrags<NT>(bnx, gsrc, src_ld, nc - g, nct, kn, KB);
        }
        long long w0 = 0;
        if (DEV && g_prof_row) w0 = clock64();
        mbar_wait(rg.full + 8u * rg.c_slot, (uint32_t)rg.c_par, P.st);
        if (DEV && g_prof_row && tid == 0) wait_cycles += clock64() - w0;
        const uint32_t sp = rg.slots + (uint32_t)(rg.c_slot * rg.SB) * 1024u + (uint32_t)lane * 16u;
        if (DEV && (P.flags & 64)) {
          kbi += nb;  // (experiment: consume the ring without touching the data)
          if (kbi > KB) kbi -= KB;
        } else if (full) {
          // a full piece inside one tile: all fragment loads are issued back to back, then 8 MMAs per n-tile on
          // independent accumulators
          if (staged) {
            const uint32_t xb = xb0 + (uint32_t)kbi * 64u;
#pragma unroll
            for (int i = 0; i < PB; ++i)
#pragma unroll
              for (int n = 0; n < NT; ++n)
                if (i < nb && n < nct) bq[i][n] = lds128(xb + (uint32_t)(i * 64 + n * 8 * xstride));
          }
          if (!(DEV && (P.flags & 16))) {  // (flag 16: experiment, skip the tensor work)
            if (PB == 4 && nb == 4) piece_full<NT, NACC, PB>(acc, sp, bq, nct);
            else piece_full<NT, NACC, 2>(acc, sp, bq, nct);
          }
          kbi += nb;
        } else {
          // general path (tiny test shapes, a tile boundary inside the piece): one block at a time, rolled
#pragma unroll 1
          for (int i = 0; i < nb; ++i) {
            const uint4 a1 = lds128(sp + i * 1024), a2 = lds128(sp + i * 1024 + 512);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              if (n < nct) {
                uint4 bb;
                if (staged) bb = lds128(xb0 + (uint32_t)(n * 8 * xstride + kbi * 64));
                else bb = (n * 8 + g < nc) ? ldcg16(gsrc + (size_t)(n * 8) * src_ld + kbi * 32) : make_uint4(0, 0, 0, 0);
                mma_bf16_16816(acc[0][n], a1.x, a1.y, a1.z, a1.w, bb.x, bb.y);
                mma_bf16_16816(acc[1][n], a2.x, a2.y, a2.z, a2.w, bb.z, bb.w);
              }
            }
            if (++kbi == KB && i + 1 < nb) {  // tile boundary inside the piece
              flush_acc<NT, NACC>(acc, pp + seg * (NT * 8) * PCOL, nct, t, g);
              kbi = 0;
              ++seg;
            }
          }
        }
        if (kbi == KB) {  // tile finished with this piece: spill its partial sums
          flush_acc<NT, NACC>(acc, pp + seg * (NT * 8) * PCOL, nct, t, g);
          kbi = 0;
          ++seg;
        }
        ring_release<DEV>(rg, P, lane, pol_stream, pol_keep);
        if (!staged) {
#pragma unroll
          for (int i = 0; i < PB; ++i)
#pragma unroll
            for (int n = 0; n < NT; ++n) bq[i][n] = bnx[i][n];
        }
        u += nb;
      }
      if (kbi != 0) flush_acc<NT, NACC>(acc, pp + seg * (NT * 8) * PCOL, nct, t, g);
      if (DEV && tid == 0 && g_prof_row) {  // warp 0: [5] cycles waiting for ring data; [7] run finished
        g_prof_row[5] = (unsigned long long)wait_cycles;
        PROF_MARK(7);
      }
    }
    cta_sync();
    PROF_MARK(3);
    // ---- cross-warp reduce + epilogue, one output element per thread-iteration
#pragma unroll 1
    for (int e = tid, it = 0; e < ((DEV && (P.flags & 128)) ? 0 : nelem); e += NTHREADS, ++it) {
      const int r = e & ((1 << rsh) - 1);
      const int tl = (e >> rsh) & ((1 << tbl) - 1);
      const int col = e >> (rsh + tbl);
      if (tl >= TB) continue;
      const int tile = t0 + round * 8 + tl;
      const int row = tile * 16 + r;
      float rs = 0.f;
      if (epi == EPI_RESID) rs = it == 0 ? resid0 : it == 1 ? resid1 : bf2f(ldcg_bf16(reinterpret_cast<const bf16*>(dst) + (size_t)col * dst_ld + row));
      float s0 = 0.f, s1 = 0.f;
      const int wf = tab->wf[tl], wl = tab->wl[tl];
#pragma unroll 4
      for (int w = wf; w <= wl; ++w) {
        const float* pq = part + ((w * 2 + (tl - tab->tl0[w])) * (NT * 8) + col) * PCOL;
        s0 += pq[r];
        if (swiglu) s1 += pq[r + 8];
      }
      if (swiglu) {
        // rows 0-7 = gate, 8-15 = up of the same 8 intermediate channels (:853-855, bf16 rounding points)
        const float gt = rbf(s0), up = rbf(s1);
        const float sl = rbf(gt / (1.f + __expf(-gt)));
        reinterpret_cast<bf16*>(dst)[(size_t)col * dst_ld + tile * 8 + r] = f2bf(sl * up);
      } else if (epi == EPI_LOGITS) {  // bf16 linear output, then .float() (HF _sample)
        reinterpret_cast<float*>(dst)[(size_t)col * dst_ld + row] = rbf(s0);
      } else {
        if (epi == EPI_BIAS) s0 += bf2f(ph.bias[row]);
        else if (epi == EPI_RESID) s0 = rs + rbf(s0);
        reinterpret_cast<bf16*>(dst)[(size_t)col * dst_ld + row] = f2bf(s0);
      }
    }
    PROF_MARK(4);
  }
  return xpar;  // (the phase loop syncs the CTA right after the body)
}

}  // namespace
