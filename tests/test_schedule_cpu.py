"""Host logic of the row-unit transposed-conv kernel's schedule (csrc/walk.h: regular workers with a fixed parity class +
helper workers that finish the tail of the position range), compiled with g++ and checked on the CPU: every (class,
position) unit is executed exactly once by exactly one worker, for the sizes the decoder uses (1 ... 4096+ decodes, 4 / 8
classes, 72 ... 148 workers) and for ragged ones, and the busiest worker carries at most one round more than the mean."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'anytime-3d-reconstruction_b200', 'csrc')

PROGRAM = textwrap.dedent(r'''
    #include <cstdio>
    #include <vector>
    #include "walk.h"
    int main() {
      const int npars[] = {4, 8};
      const int workers_list[] = {4, 8, 37, 72, 74, 144, 148, 150};
      long long cases = 0;
      for (int npar : npars)
        for (int n_pos = 8; n_pos <= 20000; n_pos = n_pos < 600 ? n_pos + 1 : n_pos * 2 + 3)
          for (int workers : workers_list) {
            if (workers < npar) continue;
            int wk = workers > n_pos * npar ? n_pos * npar : workers;
            const a3d::Walk w = a3d::make_walk(npar, n_pos, wk);
            const int launched = w.reg + w.helpers;
            if (launched > wk || w.reg % npar != 0 || w.pos_reg > n_pos) { printf("bad walk %d %d %d\n", npar, n_pos, wk); return 1; }
            std::vector<int> seen((size_t)npar * n_pos, 0);
            long long mx = 0, total = 0;
            for (int id = 0; id < launched; ++id) {
              const int c = w.count(id);
              if (c < 0) { printf("negative count\n"); return 1; }
              if (c > mx) mx = c;
              total += c;
              for (int k = 0; k < c; ++k) {
                int par = -1, pos = -1;
                if (!w.unit(id, k, par, pos)) { --total; continue; }     // empty slot of the last regular round
                if (par < 0 || par >= npar || pos < 0 || pos >= n_pos) { printf("range %d %d %d: worker %d k %d -> %d %d\n", npar, n_pos, wk, id, k, par, pos); return 1; }
                if (seen[(size_t)par * n_pos + pos]++) { printf("twice %d %d %d: %d %d\n", npar, n_pos, wk, par, pos); return 1; }
              }
            }
            if (total != (long long)npar * n_pos) { printf("coverage %d %d %d: %lld\n", npar, n_pos, wk, total); return 1; }
            // every regular worker visits every residue of the position index modulo its stride within r rounds (the
            // rotation that spreads the cheap border rows over all workers)
            if (w.reg >= npar) {
              const int r = w.reg / npar;
              if (w.pos_reg >= r * r) {
                std::vector<int> res(r, 0);
                for (int k = 0; k < r; ++k) { int par, pos; if (w.unit(0, k, par, pos)) res[pos % r]++; }
                for (int v : res) if (v != 1) { printf("rotation %d %d %d\n", npar, n_pos, wk); return 1; }
              }
            }
            // balance: the busiest worker is within one unit (regular) or one class sweep (helper) of the mean
            const double mean = (double)npar * n_pos / launched;
            if (mx > mean + npar + 1) { printf("imbalance %d %d %d: max %lld mean %.1f\n", npar, n_pos, wk, mx, mean); return 1; }
            ++cases;
          }
      printf("OK %lld cases\n", cases);
      return 0;
    }
''')


def test_walk_schedule_covers_every_unit_exactly_once(tmp_path):
    src = tmp_path / 'walk_test.cpp'
    src.write_text(PROGRAM)
    exe = tmp_path / 'walk_test'
    r = subprocess.run(['g++', '-O1', '-std=c++17', '-I', CSRC, str(src), '-o', str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.startswith('OK'), r.stdout
    assert int(r.stdout.split()[1]) > 5000


L4_PROGRAM = textwrap.dedent(r'''
    #include <cstdio>
    #include <vector>
    #include "l4_sched.h"
    using namespace a3d::l4;
    int main() {
      long long cases = 0;
      const int sms_list[] = {8, 16, 132, 144, 148, 160};
      for (int num_sms : sms_list)
        for (int n_pairs = 1; n_pairs <= 600; n_pairs = n_pairs < 40 ? n_pairs + 1 : n_pairs * 2 + 1) {
          const int n_items = n_pairs * WIN;
          const int n_cl = num_clusters(num_sms, n_items);
          if (n_cl < 4 || 2 * n_cl > num_sms + 8 || ((n_cl & 3) != 0 && (n_cl & 3) != 2)) { printf("clusters %d %d -> %d\n", num_sms, n_items, n_cl); return 1; }
          std::vector<int> seen((size_t)4 * n_items, 0);
          long long total = 0, mx = 0;
          for (int c = 0; c < n_cl; ++c) {
            const Sched sc = make_sched(c, n_cl, n_items);
            long long mine = 0;
            for (int sg = 0; sg < 2; ++sg) {
              const Seg S = sc.s[sg];
              for (int k = 0; k < S.count; ++k) {
                const int t = seg_item(S, k);
                if (t < 0) continue;
                if (t >= n_items || S.q < 0 || S.q > 3) { printf("range %d %d: cluster %d -> q %d t %d\n", num_sms, n_items, c, S.q, t); return 1; }
                if (seen[(size_t)S.q * n_items + t]++) { printf("twice %d %d: q %d t %d\n", num_sms, n_items, S.q, t); return 1; }
                const int d = item_depth(t, S.q >> 1);
                if (d < 0 || d >= WIN) { printf("depth\n"); return 1; }
                ++mine;
              }
            }
            total += mine;
            if (mine > mx) mx = mine;
          }
          if (total != 4LL * n_items) { printf("coverage %d %d: %lld of %d\n", num_sms, n_items, total, 4 * n_items); return 1; }
          const double mean = 4.0 * n_items / n_cl;
          if (mx > mean + 3) { printf("imbalance %d %d: max %lld mean %.1f\n", num_sms, n_items, mx, mean); return 1; }
          // per class, the depths 0 .. 15 of every decode-block pair appear exactly once (item t <-> (pair t / 16, depth))
          for (int q = 0; q < 4; ++q) {
            std::vector<int> dseen((size_t)n_items, 0);
            for (int t = 0; t < n_items; ++t) dseen[(size_t)(t / WIN) * WIN + item_depth(t, q >> 1)]++;
            for (int v : dseen) if (v != 1) { printf("depth permutation broken\n"); return 1; }
          }
          ++cases;
        }
      printf("OK %lld cases\n", cases);
      return 0;
    }
''')


def test_l4_item_schedule_covers_every_item_exactly_once(tmp_path):
    """csrc/l4_sched.h: the 128->64 kernel's clusters (18 regular per parity class with the rotating item walk + 2 helper
    clusters on 148 SMs; fewer on small launches) execute every (class, decode-block pair, depth) item exactly once."""
    src = tmp_path / 'l4_test.cpp'
    src.write_text(L4_PROGRAM)
    exe = tmp_path / 'l4_test'
    r = subprocess.run(['g++', '-O1', '-std=c++17', '-I', CSRC, str(src), '-o', str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.startswith('OK'), r.stdout


VARIANT_PROGRAM = textwrap.dedent(r'''
    #include <cstdio>
    #include "walk.h"
    using namespace a3d::conv_variant;
    int main() {
      // (layer geometry: decodes per block, input grid, parity classes, single-CTA cost) x call size -> expected variant
      struct Case { int nt, win, npar; float single; long long n; int want; };
      const Case cases[] = {
        {32, 4, 8, 1.6f, 32, kVarHp},      // 512->256 layer, ONE decode block: two rows of it per pair
        {32, 4, 8, 1.6f, 64, kVarPair},    // two blocks: the decode pairing is full
        {32, 4, 8, 1.6f, 72, kVarHp},      // three blocks (getEval batch 72)
        {32, 4, 8, 1.6f, 4096, kVarPair},
        {32, 4, 8, 1.6f, 8192, kVarPair},
        {16, 8, 4, 1.15f, 32, kVarPair},   // 256->128 layer
        {16, 8, 4, 1.15f, 72, kVarHp},     // five blocks: the decode pairing would run a sixth, empty one
        {16, 8, 4, 1.15f, 4096, kVarPair},
      };
      for (const Case& c : cases) {
        const int got = pick(c.n, c.nt, c.win, c.npar, 148, c.single);
        if (got != c.want) { printf("n %lld (nt %d): variant %d, expected %d\n", c.n, c.nt, got, c.want); return 1; }
      }
      // large even block counts always take the decode pairing (the h pairing pays for its zero rows, the single-CTA kernel
      // for its weight traffic)
      for (long long n = 1024; n <= 16384; n += 1024)
        if (pick(n, 32, 4, 8, 148, 1.6f) != kVarPair || pick(n, 16, 8, 4, 148, 1.15f) != kVarPair) { printf("large n %lld\n", n); return 1; }
      printf("OK\n");
      return 0;
    }
''')


def test_variant_choice_of_the_row_unit_kernel(tmp_path):
    """csrc/walk.h conv_variant::pick: the per-call choice between decode pairing, h pairing and the single-CTA kernel is
    pinned for the reference's call shapes (32 / 72 latents) and for the bench's chunks."""
    src = tmp_path / 'variant_test.cpp'
    src.write_text(VARIANT_PROGRAM)
    exe = tmp_path / 'variant_test'
    r = subprocess.run(['g++', '-O1', '-std=c++17', '-I', CSRC, str(src), '-o', str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == 'OK', r.stdout
