// Unit schedule of the row-unit transposed-conv kernel (convt_tc.cu), shared between the device code, the launcher and a
// host-only unit test (tests/test_schedule_cpu.py compiles this header with g++).
#pragma once
#ifdef __CUDACC__
#define A3D_HD __host__ __device__
#else
#define A3D_HD
#endif

namespace a3d {

// Unit schedule of a CTA (pair).  A unit is (parity class, position) with position = (decode-block group, d, h).  The
// grid is split into `reg` regular workers (a multiple of the class count: worker w owns class w % NPAR for the whole
// launch -- the weight tiles of a class stay hot in L2 -- and walks its positions with stride reg / NPAR) and up to
// NPAR - 1 helpers: 148 SMs are 74 pairs = 72 regular + 2 helpers.  Helper e finishes the positions [pos_reg, n_pos) of
// the classes e, e + H, e + 2H, ... one class after the other; pos_reg balances both kinds of worker.  With the plain
// round-robin walk the two left-over pairs (4 SMs) idled; letting EVERY worker rotate through the classes was slower.
struct Walk {
  int npar, n_pos, pos_reg, reg, helpers;
  // number of schedule slots of `worker` (a slot of the last regular round may be empty: unit() returns false)
  A3D_HD inline int count(int worker) const {
    if (worker < reg) { const int r = reg / npar; return (pos_reg + r - 1) / r; }
    return helpers > 0 ? (npar / helpers) * (n_pos - pos_reg) : 0;
  }
  // k-th slot of `worker`: class and position; false = empty slot.
  // Regular worker p0 = worker / npar takes position k r + (p0 + k) % r in round k (r workers per class): the rotation by
  // k walks every worker through all residues mod r, i.e. through all (d, h) rows.  Units on the grid border skip half or
  // three quarters of their K loop; with the plain stride-r walk (r = 18, W = 8: h advances by 2 per round) the cheap rows
  // all went to the same workers and the others finished 6 % later (ncu: sm__cycles_active avg 7.05 M, max 7.52 M).
  A3D_HD inline bool unit(int worker, int k, int& par, int& pos) const {
    if (worker < reg) {
      const int r = reg / npar, p0 = worker / npar;
      par = worker % npar;
      pos = k * r + (p0 + k) % r;
      return pos < pos_reg;
    }
    const int e = worker - reg, tail = n_pos - pos_reg;
    par = e + (k / tail) * helpers;
    pos = pos_reg + k % tail;
    return true;
  }
};

A3D_HD inline Walk make_walk(int npar, int n_pos, int workers) {
  Walk w;
  w.npar = npar;
  w.n_pos = n_pos;
  w.reg = workers - workers % npar;
  w.helpers = workers - w.reg;
  if (w.helpers > 0 && (npar % w.helpers != 0 || n_pos < 4 * (w.reg / npar))) w.helpers = 0;
  w.pos_reg = n_pos;
  if (w.helpers > 0) {
    // regular: pos_reg / r units each; helper: (npar / helpers) * (n_pos - pos_reg)
    const long long r = w.reg / npar, c = npar / w.helpers;
    w.pos_reg = (int)(((long long)n_pos * c * r + c * r) / (c * r + 1));
    if (w.pos_reg > n_pos) w.pos_reg = n_pos;
    if (w.pos_reg == n_pos) w.helpers = 0;
  }
  return w;
}


// Variant choice of the row-unit kernel per call size (host only).  Cost = units of the busiest worker x relative unit
// time: 1 for the decode pairing, 1 + 1 / (2 W) for the h pairing (its zero rows), `single_cost` for the single-CTA kernel
// (measured: it streams every weight tile to every SM and is bound by the L2 -> SM fabric on the 512->256 layer).
namespace conv_variant {
enum { kVarPair = 0, kVarHp = 1, kVarSingle = 2 };
inline int busiest(int npar, int n_pos, int workers) {      // units of the busiest worker
  if (workers > n_pos * npar) workers = n_pos * npar;
  const Walk w = make_walk(npar, n_pos, workers);
  const int r = w.reg / npar;
  const int reg_rounds = (w.pos_reg + r - 1) / r;
  const int help_rounds = w.helpers > 0 ? (npar / w.helpers) * (n_pos - w.pos_reg) : 0;
  return reg_rounds > help_rounds ? reg_rounds : help_rounds;
}
// n decodes, nt decodes per block, win = input grid, npar parity classes per position
inline int pick(long long n, int nt, int win, int npar, int num_sms, float single_cost) {
  const int nb = (int)((n + nt - 1) / nt);
  const int ww = win * win;
  const float c_pair = (float)busiest(npar, ((nb + 1) / 2) * ww, num_sms / 2);
  const float c_hp = busiest(npar, nb * ww / 2, num_sms / 2) * (1.f + 0.5f / win);
  const float c_single = busiest(npar, nb * ww, num_sms) * single_cost;
  if (c_pair <= c_hp && c_pair <= c_single) return kVarPair;
  return c_hp <= c_single ? kVarHp : kVarSingle;
}
}  // namespace conv_variant

}  // namespace a3d
