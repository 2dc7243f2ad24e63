// k_tracer_strict.cu -- tracer step in the reference's exact operation order (nvcc -fmad=false).
#define CG_TRACER_FAST 0
#include "cg_device.cuh"
namespace cg { static __constant__ GridC c_g; }
#include "k_tracer_body.cuh"
#include "k_tracer_launch.inc"
