// k_tracer_fast.cu -- tracer step with hoisted stencil coefficients and FMA contraction.
#define CG_TRACER_FAST 1
#include "cg_device.cuh"
namespace cg { static __constant__ GridC c_g; }
#include "k_tracer_body.cuh"
#include "k_tracer_launch.inc"
