// args.cuh -- launch-argument structs shared by the translation units.
#pragma once
#include <stdint.h>

namespace nb {

struct RunArgs {
  const double* m;
  double* q;
  double* v;
  const double* eps;
  double G;
  int B;
  unsigned flags;
  double dt;
  int n_steps;
  int sample_interval;
  int n_megno;
  const int32_t* n_sub;
  const int32_t* perm;
  const int32_t* n_heavy;   // device int: the first *n_heavy entries of perm go to the lane-per-body mapping
  int group_blocks;         // logical CTAs [0, group_blocks) of the main/megno kernels run that mapping
  int block0;               // logical index of this launch's CTA 0 (the main kernel may be split into head + rest)
  int block_count;          // logical CTAs of this launch (0 = all)
  const double* raw_dr;
  const double* raw_dv;
  double* dyn;
  int32_t* status;
  unsigned long long* tstamp;   // optional device [2]: {min start, max end} of the main-phase CTAs in %globaltimer ns
  double* work;             // optional [B][2] counted work: whfast {Newton iterations, Kepler solves}; ham_soft {Jacobi sweeps, S half-flows}
};

struct PrepArgs {
  const double* m;
  const double* q;
  double* v;
  const double* eps;
  double G;
  int B;
  int mode;
  unsigned flags;
  double kick_dt;
  double sched_dt;
  double dt;
  int split_n_max;
  double* h_sub_ref;
  int32_t* n_sub;
  double* stat;
};

}  // namespace nb
