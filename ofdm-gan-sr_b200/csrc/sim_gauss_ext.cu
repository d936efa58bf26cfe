// Fused simulator, Gaussian source: the combinations of generator kind, equaliser rows and late stages (fading channels, Saleh PA,
// DC offset, CFO) that the base translation unit does not build - run_benchmark(channel_type != 'awgn') and the integer generators
// inside the benchmark sweep (benchmark_comparison.py:154,179-250).
#define SIM_TU_EXT 1
#include "sim_kernel.cuh"

namespace og {
int sim_launch_gauss_ext(const SimCall& c) { return sim_launch_ext(c); }
}  // namespace og
