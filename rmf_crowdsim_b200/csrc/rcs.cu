// rcs.cu -- the C ABI declared in include/rcs.h: host side of the B200 hot path of rmf_crowdsim.
//
// Owns all device memory of a simulation handle and enqueues the per-step kernel pipeline of
// rcs_kernels.cuh / rcs_step_warp.cuh on the handle's stream.  There is no CPU implementation of any
// compute entry point: without a CUDA device every such call fails.
//
// Single translation unit on purpose: every kernel is defined once and everything is compiled with
//   nvcc -std=c++17 -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo
// (--fmad=false: rustc never contracts a*b+c, so neither may the parity-critical arithmetic).
#include "rcs_host.h"

using namespace rcs_host;

#include "rcs_host_core.inl"   // memory, groups, index rebuild, rcs_sync and rollback
#include "rcs_host_step.inl"   // the step pipeline, source sinks, events
#include "rcs_host_api.inl"    // create / destroy, planners, agents, the pub `agents` view
#include "rcs_host_index.inl"  // batched SpatialIndex, trace, options, measurement helpers
#include "rcs_host_dist.inl"   // spatial strips: NCCL and single-process transports
#include "rcs_host_inloop.inl" // study mode: the reference's in-loop index semantic as a fixed-point iteration
