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
                w.unit(id, k, par, pos);
                if (par < 0 || par >= npar || pos < 0 || pos >= n_pos) { printf("range %d %d %d: worker %d k %d -> %d %d\n", npar, n_pos, wk, id, k, par, pos); return 1; }
                if (seen[(size_t)par * n_pos + pos]++) { printf("twice %d %d %d: %d %d\n", npar, n_pos, wk, par, pos); return 1; }
              }
            }
            if (total != (long long)npar * n_pos) { printf("coverage %d %d %d: %lld\n", npar, n_pos, wk, total); return 1; }
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
