// Weight-ring geometry of the fused frame-step kernel: which 1 KB weight blocks a warp consumes in which order, and the
// iterator its TMA producer lane runs AHEAD of the consumer (across phases and grid barriers).  Pure integer logic,
// compiled for the device (ar_engine.cu) and for the host (tests/test_ring_model.py builds oracle-free C++ around
// it and checks that producer and consumer enumerate the same pieces for every CTA / warp of the shipped shapes).
//
// Layout recap: a GEMV weight [N][K] is packed as (N/16) tiles x (K/32) k-blocks of 1 KB, tile-major.
// CTA c owns tiles [t0, t0+ntc) of a phase (contiguous bytes).  Tiles are processed in rounds of <= 8; inside a
// round the TB*kb blocks are cut into 8 equal contiguous runs, one per warp.  A run is fetched in pieces of <= SB
// blocks (one ring slot, one cp.async.bulk, one mbarrier phase each).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define Q3_HD __host__ __device__ __forceinline__
#else
#define Q3_HD inline
#endif

namespace q3ring {

struct PMeta {        // per (phase, CTA): 8 bytes in shared memory
  uint32_t woff16;    // (address of this CTA's first block - weight arena base) / 16
  uint16_t ntc;       // tiles of this phase owned by the CTA (0 for non-GEMV phases and idle CTAs)
  uint16_t kb;        // K / 32
};

struct RunGeom { int TB, upw, u0, u1; };

Q3_HD int imin(int a, int b) { return a < b ? a : b; }

// blocks [u0, u1) of round `round` (relative to the round's first block) that warp `warp` consumes
Q3_HD RunGeom run_geom(int ntc, int kb, int round, int warp) {
  RunGeom g;
  g.TB = imin(8, ntc - 8 * round);
  const int units = g.TB * kb;
  g.upw = (units + 7) >> 3;
  g.u0 = imin(units, warp * g.upw);
  g.u1 = imin(units, g.u0 + g.upw);
  return g;
}

// balanced contiguous split of a phase's n_tiles over the grid: CTA c owns tq (+1 if c < tr) consecutive tiles
Q3_HD void cta_tiles(int tq, int tr, int cta, int& t0, int& ntc) {
  t0 = cta * tq + imin(cta, tr);
  ntc = tq + (cta < tr ? 1 : 0);
}

struct ProdIter {
  int it, pi, round;  // iteration of the program, phase, round of the run being fetched
  int u, u1;          // blocks of that run not yet requested: [u, u1)
  int kb, ntc;
  uint32_t woff16;
  int done;
};

Q3_HD void prod_init(ProdIter& p) {
  p.it = 0; p.pi = -1; p.round = 0; p.u = p.u1 = 0; p.kb = 0; p.ntc = 0; p.woff16 = 0; p.done = 0;
}

// advance to the next non-empty run of `warp` in program order (phases cycle `niter` times); sets done at the end
Q3_HD void prod_next_run(ProdIter& p, const PMeta* meta, int n_phases, int niter, int warp) {
  int empty = 0;
  for (;;) {
    ++p.round;
    if (p.round * 8 < p.ntc) {
      const RunGeom g = run_geom(p.ntc, p.kb, p.round, warp);
      if (g.u1 > g.u0) { p.u = g.u0; p.u1 = g.u1; return; }
      continue;
    }
    ++p.pi;
    if (p.pi >= n_phases) { p.pi = 0; ++p.it; }
    if (p.it >= niter || ++empty > n_phases) { p.done = 1; p.u = p.u1 = 0; return; }
    const PMeta m = meta[p.pi];
    p.ntc = m.ntc; p.kb = m.kb; p.woff16 = m.woff16; p.round = -1;
  }
}

// byte offset (from the weight arena base) of block u of the current run's round
Q3_HD uint64_t prod_piece_offset(const ProdIter& p) {
  return ((uint64_t)p.woff16 << 4) + (((uint64_t)(8 * p.round) * (uint64_t)p.kb + (uint64_t)p.u) << 10);
}

}  // namespace q3ring
