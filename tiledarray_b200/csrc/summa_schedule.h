// summa_schedule.h — host-side description of the SUMMA steps one rank executes.
#pragma once
#include <cstdint>
#include <vector>

struct SummaStep {
  int k = 0;
  std::vector<int> a_rows;  // global tile rows i (i % Pr == r) with A(i,k) non-zero — my row-group's A panel
  std::vector<int> b_cols;  // global tile cols j (j % Pc == c) with B(k,j) non-zero — my col-group's B panel
  bool bcast_a = false;     // the A panel is broadcast along my grid row (root column k % Pc)
  bool bcast_b = false;     // the B panel is broadcast along my grid column (root row k % Pr)
  bool compute = false;     // this rank contracts in this step
  int64_t pair_begin = 0, pair_end = 0;  // range in SummaSchedule::pair_{i,j}
};

struct SummaSchedule {
  std::vector<SummaStep> steps;
  std::vector<int32_t> pair_i, pair_j;  // (i,j) global tile coordinates, row-major within a step
  int64_t nskipped = 0;
};

SummaSchedule make_summa_schedule(int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt, const float* a,
                                  const float* b, const float* cn, float thr);
