// Host model check of the weight-ring geometry (qwen3-tts_b200/csrc/ar_ring.cuh): for every CTA and warp the TMA
// producer iterator and the consumer loop of gemv_phase must enumerate the same (offset, size) pieces in the same
// order, the ring never needs a piece that has not been requested, and the pieces of a phase tile its weight
// matrix exactly once across the grid.  Usage: ring_model <grid> <niter> <SB> <R> <n_phases> {<n_tiles> <kb>}...
// (n_tiles == 0 marks a non-GEMV phase).  Prints "OK <pieces>" or a diagnostic and exits 1.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <map>
#include "../../qwen3-tts_b200/csrc/ar_ring.cuh"
using namespace q3ring;

struct Piece { uint64_t off; int nb; };

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int grid = atoi(argv[1]), niter = atoi(argv[2]), SB = atoi(argv[3]), R = atoi(argv[4]), nph = atoi(argv[5]);
  std::vector<int> ntiles(nph), kbs(nph);
  for (int i = 0; i < nph; ++i) { ntiles[i] = atoi(argv[6 + 2 * i]); kbs[i] = atoi(argv[7 + 2 * i]); }
  // weight arena: phases back to back
  std::vector<uint64_t> wstart(nph);
  uint64_t tot = 0;
  for (int i = 0; i < nph; ++i) { wstart[i] = tot; tot += (uint64_t)ntiles[i] * kbs[i] * 1024; }
  std::vector<std::vector<int>> cover(nph);
  for (int i = 0; i < nph; ++i) cover[i].assign((size_t)ntiles[i] * kbs[i], 0);
  long long pieces = 0;
  for (int cta = 0; cta < grid; ++cta) {
    std::vector<PMeta> meta(nph);
    for (int i = 0; i < nph; ++i) {
      PMeta m{0, 0, 0};
      if (ntiles[i] > 0) {
        int t0, ntc;
        cta_tiles(ntiles[i] / grid, ntiles[i] % grid, cta, t0, ntc);
        m.ntc = (uint16_t)ntc; m.kb = (uint16_t)kbs[i];
        m.woff16 = (uint32_t)((wstart[i] + (uint64_t)t0 * kbs[i] * 1024) >> 4);
      }
      meta[i] = m;
    }
    for (int warp = 0; warp < 8; ++warp) {
      // producer stream
      std::vector<Piece> prod;
      ProdIter p; prod_init(p);
      prod_next_run(p, meta.data(), nph, niter, warp);
      while (!p.done) {
        const int nb = imin(SB, p.u1 - p.u);
        if (nb <= 0) { printf("producer: empty piece cta %d warp %d\n", cta, warp); return 1; }
        prod.push_back({prod_piece_offset(p), nb});
        p.u += nb;
        if (p.u >= p.u1) prod_next_run(p, meta.data(), nph, niter, warp);
      }
      // consumer stream (the loops of the kernel) with the ring occupancy rule: piece i may be consumed only if
      // it was requested, and requests happen R ahead at most (initial fill R, then one per release)
      size_t ci = 0;
      for (int it = 0; it < niter; ++it)
        for (int pi = 0; pi < nph; ++pi) {
          const PMeta m = meta[pi];
          for (int round = 0; round * 8 < m.ntc; ++round) {
            const RunGeom g = run_geom(m.ntc, m.kb, round, warp);
            for (int u = g.u0; u < g.u1;) {
              const int nb = imin(SB, g.u1 - u);
              const uint64_t off = ((uint64_t)m.woff16 << 4) + (((uint64_t)(8 * round) * m.kb + u) << 10);
              if (ci >= prod.size()) { printf("consumer ran past the producer: cta %d warp %d it %d phase %d\n", cta, warp, it, pi); return 1; }
              if (prod[ci].off != off || prod[ci].nb != nb) {
                printf("mismatch cta %d warp %d it %d phase %d round %d u %d: producer (%llu,%d) consumer (%llu,%d)\n", cta, warp, it, pi,
                       round, u, (unsigned long long)prod[ci].off, prod[ci].nb, (unsigned long long)off, nb);
                return 1;
              }
              if (it == 0)
                for (int b = 0; b < nb; ++b) {
                  const uint64_t blk = (off - wstart[pi]) / 1024 + b;
                  if (off < wstart[pi] || blk >= cover[pi].size()) { printf("piece outside its matrix: phase %d\n", pi); return 1; }
                  cover[pi][blk]++;
                }
              ++ci; ++pieces;
              u += nb;
            }
          }
        }
      if (ci != prod.size()) { printf("producer requested %zu pieces, consumer used %zu (cta %d warp %d)\n", prod.size(), ci, cta, warp); return 1; }
    }
  }
  for (int i = 0; i < nph; ++i)
    for (size_t b = 0; b < cover[i].size(); ++b)
      if (cover[i][b] != 1) { printf("phase %d block %zu covered %d times\n", i, b, cover[i][b]); return 1; }
  (void)R;
  printf("OK %lld\n", pieces);
  return 0;
}
