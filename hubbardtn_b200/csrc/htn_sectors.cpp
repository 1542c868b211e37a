// Host-side sector algebra and SU(2) recoupling coefficients.
// Replaces TensorKitSectors 0.1.4 / WignerSymbols 2.0.0 (Manifest.toml:1176,1302; not
// vendored) for the symmetries of HubbardFunctions.jl:245-255,341-346.  Coefficients are
// obtained by explicit contraction of Clebsch-Gordan tensors (Racah's formula), once per
// distinct spin tuple, and cached.
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "htn_internal.hpp"

namespace htn {

int sdim(int sym, Sector s) { return sym == HTN_SYM_SU2U1 ? s.q + 1 : 1; }

bool allowed(int sym, Sector a, Sector b, Sector c) {
  if (((a.p + b.p) & 1) != c.p || a.n + b.n != c.n) return false;
  if (sym == HTN_SYM_SU2U1)
    return std::abs(a.q - b.q) <= c.q && c.q <= a.q + b.q && ((a.q + b.q + c.q) & 1) == 0;
  return a.q + b.q == c.q;
}

static inline std::pair<int, int> u1key(int n) { return {std::abs(n), n >= 0 ? 0 : 1}; }

// canonical order: last factor most significant; U(1) charges 0,+1,-1,+2,-2,...; SU(2) by j;
// Z2 0<1 (same rule as oracle/sectors.py:sort_key)
bool canonical_less(int sym, Sector a, Sector b) {
  auto ka = u1key(a.n), kb = u1key(b.n);
  if (ka != kb) return ka < kb;
  if (sym == HTN_SYM_SU2U1) {
    if (a.q != b.q) return a.q < b.q;
  } else {
    auto qa = u1key(a.q), qb = u1key(b.q);
    if (qa != qb) return qa < qb;
  }
  return a.p < b.p;
}

static double fact(int n) {
  static double tab[171];
  static bool init = false;
  if (!init) {
    tab[0] = 1.0;
    for (int i = 1; i < 171; ++i) tab[i] = tab[i - 1] * i;
    init = true;
  }
  return (n < 0 || n > 170) ? 0.0 : tab[n];
}

double cg_su2(int tj1, int tm1, int tj2, int tm2, int tj3, int tm3) {
  if (tm1 + tm2 != tm3) return 0.0;
  if (tj3 < std::abs(tj1 - tj2) || tj3 > tj1 + tj2 || ((tj1 + tj2 + tj3) & 1)) return 0.0;
  if (std::abs(tm1) > tj1 || std::abs(tm2) > tj2 || std::abs(tm3) > tj3) return 0.0;
  if (((tj1 + tm1) & 1) || ((tj2 + tm2) & 1) || ((tj3 + tm3) & 1)) return 0.0;
  auto h = [](int x) { return x / 2; };
  double pref = (tj3 + 1) * fact(h(tj3 + tj1 - tj2)) * fact(h(tj3 - tj1 + tj2)) * fact(h(tj1 + tj2 - tj3)) /
                fact(h(tj1 + tj2 + tj3) + 1);
  double rad = pref * fact(h(tj3 + tm3)) * fact(h(tj3 - tm3)) * fact(h(tj1 - tm1)) * fact(h(tj1 + tm1)) *
               fact(h(tj2 - tm2)) * fact(h(tj2 + tm2));
  double sum = 0.0;
  for (int k = 0; k <= tj1 + tj2 + 1; ++k) {
    int a = h(tj1 + tj2 - tj3) - k, b = h(tj1 - tm1) - k, c = h(tj2 + tm2) - k;
    int d = h(tj3 - tj2 + tm1) + k, e = h(tj3 - tj1 - tm2) + k;
    if (a < 0 || b < 0 || c < 0 || d < 0 || e < 0) continue;
    double term = 1.0 / (fact(k) * fact(a) * fact(b) * fact(c) * fact(d) * fact(e));
    sum += (k & 1) ? -term : term;
  }
  return sum * std::sqrt(rad);
}

namespace {
struct Key9 {
  int v[9];
  bool operator<(const Key9& o) const {
    for (int i = 0; i < 9; ++i)
      if (v[i] != o.v[i]) return v[i] < o.v[i];
    return false;
  }
};
std::map<Key9, double> g_cache;
std::mutex g_cache_mu;

// N = sum_m CG(l',s'|r') CG(a,l|l') CG(l,s|r) CG(a,s'|c) CG(s,b|c) CG(b,r|r')
double network_su2(int jlp, int jsp, int jrp, int jl, int js, int jr, int ja, int jb, int jc) {
  double acc = 0.0;
  for (int ml = -jl; ml <= jl; ml += 2)
    for (int ms = -js; ms <= js; ms += 2) {
      int mr = ml + ms;
      double c3 = cg_su2(jl, ml, js, ms, jr, mr);
      if (c3 == 0.0) continue;
      for (int ma = -ja; ma <= ja; ma += 2) {
        int mlp = ma + ml;
        double c2 = cg_su2(ja, ma, jl, ml, jlp, mlp);
        if (c2 == 0.0) continue;
        for (int msp = -jsp; msp <= jsp; msp += 2) {
          int mc = ma + msp, mb = mc - ms, mrp = mlp + msp;
          double c1 = cg_su2(jlp, mlp, jsp, msp, jrp, mrp);
          if (c1 == 0.0) continue;
          double c4 = cg_su2(ja, ma, jsp, msp, jc, mc);
          if (c4 == 0.0) continue;
          double c5 = cg_su2(js, ms, jb, mb, jc, mc);
          if (c5 == 0.0) continue;
          double c6 = cg_su2(jb, mb, jr, mr, jrp, mrp);
          acc += c1 * c2 * c3 * c4 * c5 * c6;
        }
      }
    }
  return acc;
}
}  // namespace

double network(int sym, Sector lp, Sector sp, Sector rp, Sector l, Sector s, Sector r, Sector a, Sector b,
               Sector c) {
  if (!(allowed(sym, lp, sp, rp) && allowed(sym, a, l, lp) && allowed(sym, l, s, r) && allowed(sym, a, sp, c) &&
        allowed(sym, s, b, c) && allowed(sym, b, r, rp)))
    return 0.0;
  if (sym != HTN_SYM_SU2U1) return 1.0;
  Key9 k{{lp.q, sp.q, rp.q, l.q, s.q, r.q, a.q, b.q, c.q}};
  {
    std::lock_guard<std::mutex> g(g_cache_mu);
    auto it = g_cache.find(k);
    if (it != g_cache.end()) return it->second;
  }
  double v = network_su2(lp.q, sp.q, rp.q, l.q, s.q, r.q, a.q, b.q, c.q);
  if (std::fabs(v) < 1e-14) v = 0.0;
  std::lock_guard<std::mutex> g(g_cache_mu);
  g_cache[k] = v;
  return v;
}

}  // namespace htn
