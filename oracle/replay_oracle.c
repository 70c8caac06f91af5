/* oracle/replay_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the tape-replay algorithms of Adept 2.1.3 (the
 * reference under /root/reference), written from the algorithm, not from the
 * code.  It exists so that the CUDA path can be checked on machines where the
 * reference sources are absent (the GPU box).  It is pinned against the real
 * reference (oracle/_ref/libadept_ref.so, built by oracle/Makefile from the
 * unmodified sources) and against the committed golden vectors in
 * tests/golden/ by tests/test_oracle.py.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load this library.
 *
 * Compile with -ffp-contract=off: parity is "multiply, round, add, round".
 *
 * Tape layout (include/adept/Statement.h:24-30, StackStorageOrig.h:155-163):
 *   stmt[2*s] = LHS slot, stmt[2*s+1] = end_plus_one; stmt[0..1] = {-1, 0} is
 *   the sentinel pushed by new_recording (Stack.h:528-543); statement s owns
 *   operations [stmt[2*(s-1)+1], stmt[2*s+1]).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LHS(s) stmt[2*(s)]
#define END(s) stmt[2*(s)+1]

/* adept/Stack.cpp:170-190 -- forward sweep of ONE gradient vector, in place:
   a starts at +0.0, operands are accumulated in tape order, no zero skip. */
void oracle_tangent_linear(const int32_t* stmt, int64_t n_stmt,
                           const double* mult, const int32_t* idx, double* g) {
  for (int64_t s = 1; s < n_stmt; ++s) {
    double a = 0.0;
    for (int64_t op = END(s-1); op < END(s); ++op) a += mult[op]*g[idx[op]];
    g[LHS(s)] = a;
  }
}

/* adept/Stack.cpp:139-164 -- reverse sweep of ONE gradient vector, in place:
   a = g[lhs]; g[lhs] = 0; if a is non-zero scatter m*a to the operands in tape order. */
void oracle_adjoint(const int32_t* stmt, int64_t n_stmt,
                    const double* mult, const int32_t* idx, double* g) {
  for (int64_t s = n_stmt-1; s > 0; --s) {
    double a = g[LHS(s)];
    g[LHS(s)] = 0.0;
    if (a != 0.0)
      for (int64_t op = END(s-1); op < END(s); ++op) g[idx[op]] += mult[op]*a;
  }
}

/* Offset defaulting of the raw-pointer entry points (adept/jacobian.cpp:215-220, :473-478):
   a non-positive offset means "the size of the OTHER dimension". */
static void default_offsets(int64_t n_indep, int64_t n_dep, int64_t* dep_offset, int64_t* indep_offset) {
  if (*dep_offset <= 0)   *dep_offset = n_indep;
  if (*indep_offset <= 0) *indep_offset = n_dep;
}

/* adept/jacobian.cpp:127-315 -- forward Jacobian in blocks of P independents.
   Per block: zero G*P workspace, seed g[indep[j]*P + lane] = 1, sweep all
   lanes, copy out out[i*dep_offset + j*indep_offset] = g[dep[i]*P + lane].
   Returns 10 if either list is empty (dependents_or_independents_not_identified, :208-210). */
int oracle_jacobian_forward(const int32_t* stmt, int64_t n_stmt, const double* mult, const int32_t* idx,
                            int64_t max_gradient, const int32_t* indep, int64_t n_indep,
                            const int32_t* dep, int64_t n_dep, double* out,
                            int64_t dep_offset, int64_t indep_offset, int P, int n_threads) {
  if (n_indep == 0 || n_dep == 0) return 10;
  default_offsets(n_indep, n_dep, &dep_offset, &indep_offset);
  int64_t n_block = (n_indep + P - 1)/P;
  int failed = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
#endif
  {
    double* g = (double*)malloc(sizeof(double)*(size_t)max_gradient*P);
    double* a = (double*)malloc(sizeof(double)*P);
    if (!g || !a) failed = 1;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t b = 0; b < n_block; ++b) {
      if (failed) continue;
      int64_t j0 = b*P;
      int w = (int)((n_indep - j0 < P) ? (n_indep - j0) : P);
      memset(g, 0, sizeof(double)*(size_t)max_gradient*P);
      for (int l = 0; l < w; ++l) g[(int64_t)indep[j0+l]*P + l] = 1.0;
      for (int64_t s = 1; s < n_stmt; ++s) {
        for (int l = 0; l < w; ++l) a[l] = 0.0;
        for (int64_t op = END(s-1); op < END(s); ++op) {
          const double m = mult[op];
          const double* src = g + (int64_t)idx[op]*P;
          for (int l = 0; l < w; ++l) a[l] += m*src[l];
        }
        double* dst = g + (int64_t)LHS(s)*P;
        for (int l = 0; l < w; ++l) dst[l] = a[l];
      }
      for (int64_t i = 0; i < n_dep; ++i)
        for (int l = 0; l < w; ++l)
          out[i*dep_offset + (j0+l)*indep_offset] = g[(int64_t)dep[i]*P + l];
    }
    free(g); free(a);
  }
  return failed ? 1 : 0;
}

/* adept/jacobian.cpp:331-647 -- reverse Jacobian in blocks of P dependents.
   Per block: zero, seed g[dep[i]][lane] = 1, backward sweep; a statement is
   skipped only when ALL lanes of the block hold zero (the per-lane sparse branch
   behind "#if MULTIPASS_SIZE > MULTIPASS_SIZE_ZERO_CHECK" is never compiled:
   neither name is a macro, :387), otherwise every lane is updated; copy out
   out[j*indep_offset + i*dep_offset] = g[indep[j]][lane]. */
int oracle_jacobian_reverse(const int32_t* stmt, int64_t n_stmt, const double* mult, const int32_t* idx,
                            int64_t max_gradient, const int32_t* indep, int64_t n_indep,
                            const int32_t* dep, int64_t n_dep, double* out,
                            int64_t dep_offset, int64_t indep_offset, int P, int n_threads) {
  if (n_indep == 0 || n_dep == 0) return 10;
  default_offsets(n_indep, n_dep, &dep_offset, &indep_offset);
  int64_t n_block = (n_dep + P - 1)/P;
  int failed = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
#endif
  {
    double* g = (double*)malloc(sizeof(double)*(size_t)max_gradient*P);
    double* a = (double*)malloc(sizeof(double)*P);
    if (!g || !a) failed = 1;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t b = 0; b < n_block; ++b) {
      if (failed) continue;
      int64_t i0 = b*P;
      int w = (int)((n_dep - i0 < P) ? (n_dep - i0) : P);
      memset(g, 0, sizeof(double)*(size_t)max_gradient*P);
      for (int l = 0; l < w; ++l) g[(int64_t)dep[i0+l]*P + l] = 1.0;
      for (int64_t s = n_stmt-1; s > 0; --s) {
        double* lhs = g + (int64_t)LHS(s)*P;
        int any = 0;
        for (int l = 0; l < w; ++l) { a[l] = lhs[l]; lhs[l] = 0.0; if (a[l] != 0.0) any = 1; }
        if (!any) continue;
        for (int64_t op = END(s-1); op < END(s); ++op) {
          const double m = mult[op];
          double* dst = g + (int64_t)idx[op]*P;
          for (int l = 0; l < w; ++l) dst[l] += m*a[l];
        }
      }
      for (int64_t j = 0; j < n_indep; ++j)
        for (int l = 0; l < w; ++l)
          out[j*indep_offset + (i0+l)*dep_offset] = g[(int64_t)indep[j]*P + l];
    }
    free(g); free(a);
  }
  return failed ? 1 : 0;
}

/* Sequential reference for the dependency analysis the CUDA engine does on the
   device (no counterpart in the reference: it replays strictly in tape order).
   level[s] (1-based; level[0] = 0 for the sentinel) is the earliest step at which
   statement s may run under a constraint set that keeps every bit of every sweep:
     mode 0: RAW, WAR and WAW on slots; concurrent readers of a slot version are free.
             Enough for gather sweeps (forward), where each statement sums its own
             operands in tape order.
     mode 1: every pair of statements touching a common slot keeps its tape order.
     mode 2: mode 0 plus "readers of one slot version have non-decreasing levels in tape
             order".  A reverse sweep that walks levels downwards and, inside a level,
             adds the contributions to each slot in (statement descending, operand
             ascending) order then reproduces the reference's accumulation order exactly.
   `window` > 0 restarts the analysis every `window` statements (levels are then local
   to the window; the schedule is (window, level) lexicographic); with `huge` > 0 a statement
   of more than `huge` operands is a window of its own.  With `pack` every local level of a
   window is mapped to the smallest global step that respects all earlier windows (summarised
   per slot) and stays above the window's previous level, so independent windows share steps;
   without it windows are simply stacked.  level[] holds GLOBAL steps.  Returns the number of
   steps (sum of the windows' depths). */
int64_t oracle_levels(const int32_t* stmt, int64_t n_stmt, const int32_t* idx,
                      int64_t max_gradient, int mode, int64_t window, int64_t huge, int pack, int32_t* level) {
  int32_t* wl = (int32_t*)calloc((size_t)max_gradient, sizeof(int32_t)); /* level of last writer  */
  int32_t* rl = (int32_t*)calloc((size_t)max_gradient, sizeof(int32_t)); /* max level of readers of the current version */
  int64_t* ep = (int64_t*)calloc((size_t)max_gradient, sizeof(int64_t)); /* window tag of the entry */
  int32_t* gw = (int32_t*)calloc((size_t)max_gradient, sizeof(int32_t)); /* packing: global step of the last writer */
  int32_t* gr = (int32_t*)calloc((size_t)max_gradient, sizeof(int32_t)); /* packing: max global step of live readers */
  int64_t steps = 0, base = 0;
  int64_t cur_win = 0;
  level[0] = 0;
  int64_t s = 1;
  while (s < n_stmt) {
    /* the window [s, e): starts every `window` statements; a statement with more than `huge`
       operands is a window of its own */
    int64_t e = s + 1;
    if (window <= 0) e = n_stmt;
    else {
      int big = huge > 0 && (END(s) - END(s-1)) > huge;
      if (!big)
        while (e < n_stmt && (e - 1) % window != 0 && !(huge > 0 && (END(e) - END(e-1)) > huge)) ++e;
    }
    ++cur_win;
    int32_t depth = 0;
#define TOUCH(x) do { if (ep[x] != cur_win) { ep[x] = cur_win; wl[x] = 0; rl[x] = 0; } } while (0)
    for (int64_t t = s; t < e; ++t) {
      int32_t lv = 0;
      int32_t L = LHS(t);
      TOUCH(L);
      for (int64_t op = END(t-1); op < END(t); ++op) {
        int32_t x = idx[op];
        TOUCH(x);
        if (wl[x] + 1 > lv) lv = wl[x] + 1;
        if (mode == 1 && rl[x] + 1 > lv) lv = rl[x] + 1;
        if (mode == 2 && rl[x] > lv) lv = rl[x];
      }
      if (wl[L] + 1 > lv) lv = wl[L] + 1;
      if (rl[L] + 1 > lv) lv = rl[L] + 1;
      if (lv < 1) lv = 1;
      level[t] = lv;
      for (int64_t op = END(t-1); op < END(t); ++op) {
        int32_t x = idx[op];
        if (lv > rl[x]) rl[x] = lv;
      }
      /* the write retires the old version (a statement's own reads belong to it) */
      wl[L] = lv; rl[L] = (mode == 1) ? lv : 0;
      if (lv > depth) depth = lv;
    }
    if (!pack) {
      for (int64_t t = s; t < e; ++t) level[t] += (int32_t)base;
      if (base + depth > steps) steps = base + depth;
    } else {
      /* every local level l of the window gets its own global step g[l], strictly increasing in l:
         g[l] = max(g[l-1]+1, the largest lower bound earlier windows impose on a statement of level l) */
      int32_t* g = (int32_t*)calloc((size_t)depth + 1, sizeof(int32_t));
      for (int64_t t = s; t < e; ++t) {
        int32_t L = LHS(t), need, l = level[t];
        for (int64_t op = END(t-1); op < END(t); ++op) {
          int32_t x = idx[op];
          need = (gw[x] + 1 > gr[x] ? gw[x] + 1 : gr[x]);
          if (need > g[l]) g[l] = need;
        }
        need = (gw[L] > gr[L] ? gw[L] : gr[L]) + 1;
        if (need > g[l]) g[l] = need;
      }
      for (int32_t l = 1; l <= depth; ++l) if (g[l] < g[l-1] + 1) g[l] = g[l-1] + 1;
      /* fold the window's final per-slot state into the summary (idempotent per slot) */
      for (int64_t t = s; t < e; ++t) {
        for (int64_t op = END(t-1); op <= END(t); ++op) {
          int32_t x = (op < END(t)) ? idx[op] : LHS(t);
          if (wl[x] > 0) { gw[x] = g[wl[x]]; gr[x] = rl[x] > 0 ? g[rl[x]] : 0; }
          else if (rl[x] > 0 && g[rl[x]] > gr[x]) gr[x] = g[rl[x]];
        }
      }
      for (int64_t t = s; t < e; ++t) level[t] = g[level[t]];
      if (g[depth] > steps) steps = g[depth];
      free(g);
    }
    base += depth;
    s = e;
  }
  free(wl); free(rl); free(ep); free(gw); free(gr);
  return steps;
}
