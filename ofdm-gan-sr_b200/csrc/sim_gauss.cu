// Fused simulator instantiations for the reference's synthetic source (Gaussian symbols, N = 16, no CP;
// utils/dataset.py:243-247): simulate-only and + fp32 / Q spec / Q rtl_literal generator.
#include "sim_kernel.cuh"

namespace og {
int sim_launch_gauss(const SimCall& c) { return sim_launch_src<SRC_GAUSS>(c); }
}  // namespace og
