// gca_step_common.cuh - helpers shared by the step kernels of gca_step.cu and gca_step_fc.cu (sm_100a).
#pragma once
#include <cstdlib>
#include <utility>

#include "gca_device.cuh"

namespace gca {

// bits of own_b.z (the ownship record the streaming pass reads)
constexpr uint32_t kOwnRuns = 1u;       // the reference's intruder loop runs for this env in this step
constexpr uint32_t kOwnPlane = 2u;      // parity of the env's current position plane
constexpr uint32_t kOwnSkip = 4u;       // forecast step: another role advances this env (hot / resetting): the lane does nothing
constexpr uint32_t kOwnSlotShift = 3u;  // forecast step: bits 3-4 = tick % 3, the forecast slot this step reads
constexpr uint32_t kOwnConf = 32u;      // forecast step: an intruder can be inside minimum_separation after this step
constexpr uint32_t kOwnHot = 64u;       // forecast step: class of the env for the jobs kernel - a warp replays the reference's loop
constexpr uint32_t kOwnReset = 128u;    //                                                   - the env finished: reset()'s spawns

// Programmatic dependent launch: the kernels of a step are launched with programmatic stream serialization, so
// the next grid is staged (and its blocks scheduled as slots free up) while the current one drains.  A kernel
// calls pdl_wait() before it touches anything an earlier kernel of the stream wrote.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ownship state right after reset: (50, 50), min speed, heading pi/4 (PKG/SingleAircraftEnv.py:72-76), or with
// random_start random_pos(), random_speed(), random_heading() drawn in that order BEFORE the intruders
// (Simulators/SingleAircraftDiscrete9HEREnv.py:78-82); f32 position and velocity (Aircraft.__init__ :269-276)
template <bool TAPE>
__device__ __forceinline__ void reset_ownship(const gca_config& c, Draws<TAPE>& d, float2& pos, double2& hs, double2& vel) {
  double sn, cs;
  if (c.random_start) {
    double x, y, speed, heading;
    draw_pos(d, c, GCA_SLOT_OWN_RESET, GCA_BLOCK_POS, x, y);
    draw_speed_heading(d, c, GCA_SLOT_OWN_RESET, speed, heading);
    pos = make_float2((float)x, (float)y);
    hs = make_double2(heading, speed);
  } else {
    pos = make_float2(50.0f, 50.0f);
    hs = make_double2(3.141592653589793 / 4, c.min_speed);
  }
  gca_sincos(hs.x, &sn, &cs);
  vel = make_double2((double)(float)__dmul_rn(hs.y, cs), (double)(float)__dmul_rn(hs.y, sn));
}

template <bool TAPE>
__device__ __forceinline__ Draws<TAPE> make_draws(const StepArgs& a, size_t me, uint32_t tick) {
  Draws<TAPE> d;
  if constexpr (TAPE) {
    d.tape = a.tape + me * (size_t)a.tape_stride;
    d.cur = a.cursor[me];
  } else {
    d.k0 = a.key0; d.k1 = a.key1;
    d.env = a.env_id0 + (uint32_t)me;
    d.tick = tick;
  }
  return d;
}

__device__ __forceinline__ void st_release_pair(float* p, float x, float y) {
  asm volatile("st.volatile.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void st_release_quad(float* p, float x, float y, float z, float w) {
  asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// launch with programmatic stream serialization (see pdl_wait above)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned blocks, unsigned threads, cudaStream_t st, Args&&... args) {
  static const int use_pdl = std::getenv("GCA_NO_PDL") ? 0 : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Kernels that are meant to share SMs must agree on the L1 / shared-memory split: an SM only changes it when it is
// idle, so blocks of a kernel that prefers another split wait until the resident kernel has left the SM (measured:
// the streaming kernel started 21 us late behind a persistent head kernel with the default preference).
inline int shared_carveout_percent() {
  static const int pct = std::getenv("GCA_CARVEOUT") ? std::atoi(std::getenv("GCA_CARVEOUT")) : 72;   // 164 KB of 228 KB
  return pct;
}
template <typename... KArgs>
static void prefer_carveout(void (*kernel)(KArgs...)) {
  cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout, shared_carveout_percent());
}

// plain stream-ordered launch (for kernels that do not call pdl_wait)
template <typename... KArgs, typename... Args>
static cudaError_t launch_plain(void (*kernel)(KArgs...), unsigned blocks, unsigned threads, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.stream = st;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace gca
