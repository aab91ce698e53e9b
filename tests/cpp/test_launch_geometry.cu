// Host-only checks of launch-geometry helpers of the planner kernels (compiled with nvcc, no device
// code is run): the number of classifying CTAs k_scatter is launched with must cover every tile of
// every query window that fits the 256 x 256 grid - a short count would leave cells unclassified.
#include <cstdio>

#include "../../kompass-core_b200/csrc/kc_planner_kernels.cuh"

int main() {
  using namespace kc;
  long long checked = 0;
  int worst_spare = 1 << 30;
  for (int qw = 1; qw <= kGridN; ++qw)
    for (int qh = 1; qh <= kGridN; ++qh) {
      const int tiles = ((qw + kClassCols - 1) / kClassCols) * ((qh + kClassRows - 1) / kClassRows);
      const int launched = class_ctas_for(qw * qh);
      if (launched < tiles) {
        std::printf("FAIL: window %d x %d needs %d tiles, %d CTAs launched\n", qw, qh, tiles, launched);
        return 1;
      }
      if (launched - tiles < worst_spare) worst_spare = launched - tiles;
      ++checked;
    }
  // a batch launches for its largest window: the count must be monotone in the number of cells
  for (int c = 1; c < kGridN * kGridN; ++c)
    if (class_ctas_for(c + 1) < class_ctas_for(c)) {
      std::printf("FAIL: class_ctas_for not monotone at %d\n", c);
      return 1;
    }
  static_assert(kClassCols * kClassRows == 256, "a classifying CTA is one of k_scatter's 256-thread CTAs");
  std::printf("ALL PASSED (%lld windows, tightest margin %d CTAs)\n", checked, worst_spare);
  return 0;
}
