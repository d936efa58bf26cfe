// Fused simulator instantiations for the QPSK + OFDMModulator sources (utils/ofdm_utils.py:105-109,281-329):
// N in {16, 8} x cyclic prefix in {0, 2}, each simulate-only and + fp32 / Q spec / Q rtl_literal generator.
#include "sim_kernel.cuh"

namespace og {
int sim_launch_qpsk(const SimCall& c) {
    switch (c.src) {
        case SRC_Q16_CP0: return sim_launch_src<SRC_Q16_CP0>(c);
        case SRC_Q16_CP2: return sim_launch_src<SRC_Q16_CP2>(c);
        case SRC_Q8_CP0: return sim_launch_src<SRC_Q8_CP0>(c);
        case SRC_Q8_CP2: return sim_launch_src<SRC_Q8_CP2>(c);
        default: return OFDMGAN_E_ARG;
    }
}
}  // namespace og
