// Instantiates the fused closed-loop kernel for env type double, network type double.
#define H3_LOOP_IMPL
#include "h3_launch.cuh"

namespace h3 {
template int launch_closed_loop<double, double>(const H3Config&, const H3Batch&, int, const double*, const double*, int, int,
                                        const H3Traj*, double*, void*, cudaStream_t, int, LoopInfo*);
}
