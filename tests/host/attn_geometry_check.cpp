// Host-side exhaustive check of the attention geometry helpers of dalle_pytorch_b200/csrc/attn_common.cuh (the functions every
// attention kernel uses to decide which (query, key) pairs exist, which tiles can be skipped and which need no per-element test).
// The header is plain C++ apart from the CUDA function-space keywords, so it is compiled here with g++ against a stub "common.cuh"
// (tests/test_attn_geometry_cpu.py copies the header next to the stub).  Test infrastructure only.
//
//   check   : row / column bit masks == the element predicate, tile-skip is conservative, tile-full is exact, gather maps are
//             bijections, segment tiles partition the virtual sequence -- over a grid of ragged geometries; prints "OK <cases>"
//   dump ...: prints the allowed(i, j) matrix of one geometry as 0/1 rows (compared with the pinned CPU oracle by the python test)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "attn_common.cuh"

using namespace db200;

static int failures = 0;
#define EXPECT(cond, ...)                                  \
  do {                                                     \
    if (!(cond)) {                                         \
      if (failures < 20) { printf("FAIL: "); printf(__VA_ARGS__); printf("\n"); } \
      ++failures;                                          \
    }                                                      \
  } while (0)

static AttnGeom geom(int pattern, int causal, int T, int fm, int ks, int dil, int n_q, int n_k, const uint8_t* sm) {
  AttnGeom g;
  memset(&g, 0, sizeof g);
  g.pattern = pattern; g.causal = causal; g.text_len = T; g.fmap = fm > 0 ? fm : 1; g.ksize = ks; g.dil = dil > 0 ? dil : 1;
  g.n_q = n_q; g.n_k = n_k; g.static_mask = sm; g.static_ld = n_k; g.n_alloc = n_k; g.n_stat = n_q; g.kv_rows = n_k;
  return g;
}

static unsigned rng_state = 12345u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static long check_one(const AttnGeom& g, const uint8_t* km) {
  long cases = 0;
  const int n_q = g.n_q, n_k = g.n_k, off = n_k - n_q;
  const int widths[3] = {32, 64, 128};
  for (int wi = 0; wi < 3; ++wi) {
    const int W = widths[wi];
    // row masks: query i (absolute position), key tile [k0, k0 + W)
    for (int qi = 0; qi < n_q; ++qi) {
      const int i = qi + off;
      for (int k0 = 0; k0 < n_k; k0 += W) {
        const Mask128 m = attn_row_bits(g, i, k0, km, W);
        for (int b = 0; b < 128; ++b) {
          const int j = k0 + b;
          const bool want = b < W && j < n_k && attn_allowed(g, i, j) && (km == nullptr || km[j] != 0);
          const bool got = (m.w[b >> 5] >> (b & 31)) & 1u;
          EXPECT(want == got, "row_bits pattern %d causal %d T %d fm %d ks %d dil %d n_q %d n_k %d: i %d j %d want %d", g.pattern, g.causal,
                 g.text_len, g.fmap, g.ksize, g.dil, n_q, n_k, i, j, (int)want);
        }
        ++cases;
      }
    }
    // column masks (backward kernels: n_q == n_k)
    if (n_q == n_k) {
      for (int j = 0; j < n_k; ++j)
        for (int q0 = 0; q0 < n_q; q0 += W) {
          const Mask128 m = attn_col_bits(g, j, q0, n_q, W);
          for (int b = 0; b < 128; ++b) {
            const int i = q0 + b;
            const bool want = b < W && i < n_q && attn_allowed(g, i, j);
            const bool got = (m.w[b >> 5] >> (b & 31)) & 1u;
            EXPECT(want == got, "col_bits pattern %d causal %d T %d fm %d ks %d dil %d n %d: i %d j %d want %d", g.pattern, g.causal, g.text_len,
                   g.fmap, g.ksize, g.dil, n_q, i, j, (int)want);
          }
          ++cases;
        }
    }
    // tile predicates over (query tile of 16 / 128 rows) x (key tile of W)
    const int qts[2] = {16, 128};
    for (int qt = 0; qt < 2; ++qt)
      for (int q0 = 0; q0 < n_q; q0 += qts[qt])
        for (int k0 = 0; k0 < n_k; k0 += W) {
          const int q1 = (q0 + qts[qt] < n_q ? q0 + qts[qt] : n_q) - 1, k1 = (k0 + W < n_k ? k0 + W : n_k) - 1;
          bool any = false, all = true;
          for (int i = q0; i <= q1; ++i)
            for (int j = k0; j <= k1; ++j) {
              const bool a = attn_allowed(g, i + off, j);
              any |= a;
              all &= a;
            }
          if (!attn_tile_needed(g, q0 + off, q1 + off, k0, k1)) EXPECT(!any, "tile_needed skipped a live tile: pattern %d T %d fm %d q %d..%d k %d..%d", g.pattern, g.text_len, g.fmap, q0, q1, k0, k1);
          if (attn_tile_full(g, q0 + off, q1 + off, k0, k1)) EXPECT(all, "tile_full on a masked tile: pattern %d T %d fm %d q %d..%d k %d..%d", g.pattern, g.text_len, g.fmap, q0, q1, k0, k1);
          ++cases;
        }
  }
  return cases;
}

static long check_gather(int T, int fm) {
  long cases = 0;
  for (int col = 0; col < 2; ++col) {
    AttnGeom g = geom(DB200_ATTN_AXIAL_ROW, 1, T, fm, 0, 1, T + fm * fm - 1, T + fm * fm - 1, nullptr);
    g.gather = 1; g.col = col; g.n_alloc = gather_n_alloc(T, fm); g.n_stat = gather_n_stat(T, fm); g.t_pad = gather_t_pad(T);
    const int n = T + fm * fm;                         // virtual positions incl. the pad token
    std::vector<int> seen(n, 0);
    for (int u = 0; u < n; ++u) {
      const int p = attn_nat(g, u);
      EXPECT(p >= 0 && p < n, "attn_nat out of range T %d fm %d col %d u %d -> %d", T, fm, col, u, p);
      if (p >= 0 && p < n) seen[p]++;
      EXPECT(attn_virt(g, p) == u, "attn_virt(attn_nat(u)) != u: T %d fm %d col %d u %d", T, fm, col, u);
      const int si = attn_sidx(g, u);
      EXPECT(si >= 0 && si < g.n_stat && (u < T ? si == u : si == g.t_pad + (u - T)), "attn_sidx T %d fm %d u %d -> %d", T, fm, u, si);
    }
    for (int p = 0; p < n; ++p) EXPECT(seen[p] == 1, "attn_nat is not a bijection: T %d fm %d col %d p %d hit %d times", T, fm, col, p, seen[p]);
    // the predicate in virtual coordinates == the natural-order predicate of the axis
    AttnGeom nat = geom(col ? DB200_ATTN_AXIAL_COL : DB200_ATTN_AXIAL_ROW, 1, T, fm, 0, 1, n, n, nullptr);
    for (int u = 0; u < n; ++u)
      for (int w = 0; w < n; ++w)
        EXPECT(attn_allowed(g, u, w) == attn_allowed(nat, attn_nat(g, u), attn_nat(g, w)), "virtual predicate T %d fm %d col %d (%d,%d)", T, fm, col, u, w);
    // segment tiles partition [0, n) for every width that divides the image
    const int widths[3] = {32, 64, 128};
    for (int wi = 0; wi < 3; ++wi) {
      const int W = widths[wi];
      if ((fm * fm) % W) continue;
      const SegTiles S(g, W, n);
      EXPECT(S.count() == seg_tile_count(g, W, n), "seg_tile_count T %d fm %d W %d", T, fm, W);
      std::vector<int> cover(n, 0);
      for (int t = 0; t < S.count(); ++t)
        for (int u = S.origin(t); u < S.limit(t); ++u) {
          EXPECT(u >= 0 && u < n, "tile %d reaches %d of %d", t, u, n);
          if (u >= 0 && u < n) cover[u]++;
          EXPECT((u >= T) == S.is_img(t), "tile %d mixes segments at %d (T %d)", t, u, T);
        }
      for (int u = 0; u < n; ++u) EXPECT(cover[u] == 1, "segment tiles T %d fm %d W %d: position %d covered %d times", T, fm, W, u, cover[u]);
      ++cases;
    }
  }
  return cases;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "dump")) {            // dump pattern causal T fm ks dil n
    const int pattern = atoi(argv[2]), causal = atoi(argv[3]), T = atoi(argv[4]), fm = atoi(argv[5]), ks = atoi(argv[6]), dil = atoi(argv[7]),
              n = atoi(argv[8]);
    const AttnGeom g = geom(pattern, causal, T, fm, ks, dil, n, n, nullptr);
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) putchar(attn_allowed(g, i, j) ? '1' : '0');
      putchar('\n');
    }
    return 0;
  }
  long cases = 0;
  const int Ts[6] = {1, 2, 3, 8, 33, 65}, fms[8] = {1, 2, 3, 4, 5, 6, 8, 9};
  for (int ti = 0; ti < 6; ++ti)
    for (int fi = 0; fi < 8; ++fi) {
      const int T = Ts[ti], fm = fms[fi], n = T + fm * fm - 1;      // training length (the last image token is never an input)
      if (n < 1) continue;
      std::vector<uint8_t> km(n), sm((size_t)n * n);
      for (auto& x : km) x = rnd() % 4 != 0;
      for (auto& x : sm) x = rnd() % 3 != 0;
      const int nqs[3] = {n, 1, n > 5 ? 5 : n};
      for (int qi = 0; qi < 3; ++qi) {
        const int n_q = nqs[qi];
        for (int causal = 0; causal < 2; ++causal) {
          cases += check_one(geom(DB200_ATTN_FULL, causal, T, fm, 0, 1, n_q, n, nullptr), nullptr);
          cases += check_one(geom(DB200_ATTN_FULL, causal, T, fm, 0, 1, n_q, n, nullptr), km.data());
          cases += check_one(geom(DB200_ATTN_STATIC, causal, T, fm, 0, 1, n_q, n, sm.data()), nullptr);
        }
        cases += check_one(geom(DB200_ATTN_AXIAL_ROW, 1, T, fm, 0, 1, n_q, n, nullptr), nullptr);
        cases += check_one(geom(DB200_ATTN_AXIAL_COL, 1, T, fm, 0, 1, n_q, n, nullptr), km.data());
        const int kss[3] = {1, 3, 5}, dils[3] = {1, 2, 3};
        for (int a = 0; a < 3; ++a)
          for (int d = 0; d < 3; ++d) cases += check_one(geom(DB200_ATTN_CONV_LIKE, 1, T, fm, kss[a], dils[d], n_q, n, nullptr), nullptr);
      }
      // shorter-than-full sequences (generation prefix / standalone modules, attention.py:255-258)
      if (n > 4) cases += check_one(geom(DB200_ATTN_AXIAL_COL, 1, T, fm, 0, 1, n - 3, n - 3, nullptr), nullptr);
    }
  {   // benchmark geometry (text 257 incl. <bos>, 32 x 32 image): every pattern once
    const int T = 257, fm = 32, n = T + fm * fm - 1;
    cases += check_one(geom(DB200_ATTN_FULL, 1, T, fm, 0, 1, n, n, nullptr), nullptr);
    cases += check_one(geom(DB200_ATTN_AXIAL_ROW, 1, T, fm, 0, 1, n, n, nullptr), nullptr);
    cases += check_one(geom(DB200_ATTN_AXIAL_COL, 1, T, fm, 0, 1, n, n, nullptr), nullptr);
    cases += check_one(geom(DB200_ATTN_CONV_LIKE, 1, T, fm, 5, 1, n, n, nullptr), nullptr);
    cases += check_one(geom(DB200_ATTN_CONV_LIKE, 1, T, fm, 3, 2, n, n, nullptr), nullptr);
    cases += check_one(geom(DB200_ATTN_FULL, 1, T, fm, 0, 1, 1, n, nullptr), nullptr);          // one decoding step
  }
  const int gT[4] = {1, 7, 64, 257}, gfm[3] = {8, 16, 32};
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 3; ++b) cases += check_gather(gT[a], gfm[b]);
  if (failures) { printf("FAILED %d checks\n", failures); return 1; }
  printf("OK %ld\n", cases);
  return 0;
}
