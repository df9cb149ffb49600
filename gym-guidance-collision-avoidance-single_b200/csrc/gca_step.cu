// gca_step.cu - the fused step / reset / observe kernels (sm_100a).
//
// Mapping.  A warp owns a tile of TILE consecutive environments and never synchronises with
// any other warp:
//   phase A  lane = env of the tile: ownship kinematics (all the f64 work: Philox + Box-Muller,
//            sincos, clamp) packed 32 envs per warp-instruction so the FP64 pipe stays off the
//            critical path;
//   phase B  for each env of the tile in turn, lanes = intruders: advance, f32/f64 separation,
//            map test, conflict / NMAC; the reference's sequential loop semantics (first NMAC
//            index wins and freezes every later intruder, respawn lands between the distance
//            and the conflict test, the conflict flag never clears) are recovered with
//            __ballot_sync / __ffs / __popc; intruder observations go out from registers as
//            16-byte stores;
//   phase C  lane = env again: respawns of the intruders that left the map, wall / goal /
//            reward / done, the ownship + goal tail of the observation, counters;
//   phase D  VecEnv auto-reset of finished envs, lanes = intruders (PHILOX) so that a rare
//            80-spawn reset costs three warp rounds instead of a serial tail.
// Every intruder row is read once and written once; nothing is staged through global scratch.
#include <cstdio>
#include <type_traits>

#include "gca_device.cuh"
#include "gca_launch.h"

namespace gca {

constexpr int kWarpsPerBlock = 4;

template <bool FAITH>
__device__ __forceinline__ void load_round(const DevState& s, size_t env, int r, int lane, Intruder& it, bool& valid,
                                           uint32_t& fw) {
  const int i = r * 32 + lane;
  valid = i < s.N;
  it.px = it.py = 0.0;
  it.vx = it.vy = 0.0f;
  if (valid) load_intruder<FAITH>(s, env * (size_t)s.Np + i, it);
  fw = s.iflag[env * (size_t)s.W + r];
  uint32_t dw = 0;
  if constexpr (FAITH) dw = s.if64[env * (size_t)s.W + r];
  it.is64 = FAITH && ((dw >> lane) & 1u);
}

// reset(): PKG/SingleAircraftEnv.py:66-98 for env `env`, executed by the whole warp.
// PHILOX: lanes = intruders.  TAPE: lane `owner` replays the reference's sequential draw order.
// Returns (in the owner lane) the new ownship / goal state through the reference arguments.
template <bool FAITH, bool TAPE>
__device__ __forceinline__ void reset_env_warp(const StepArgs& a, size_t env, int lane, int owner, Draws<TAPE>& d,
                                               double2& goal) {
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const float ox = 50.0f, oy = 50.0f;                     // Ownship(position=(50, 50), ...) :72-76
  if constexpr (TAPE) {
    if (lane == owner) {
      for (int r = 0; r < s.W; ++r) {
        uint32_t dw = 0;
        for (int j = 0; j < 32 && r * 32 + j < s.N; ++j) {
          const int i = r * 32 + j;
          Intruder it;
          spawn<FAITH, TAPE>(d, c, GCA_SLOT_RESET | (uint32_t)i, ox, oy, it);
          store_ipos<FAITH>(s, env * (size_t)s.Np + i, it.px, it.py);
          s.ivel[env * (size_t)s.Np + i] = make_float2(it.vx, it.vy);
          dw |= (it.is64 ? 1u : 0u) << j;
          write_obs_intruder<FAITH>(a, env, i, it, it.px, it.py);
        }
        s.iflag[env * (size_t)s.W + r] = 0u;
        if constexpr (FAITH) s.if64[env * (size_t)s.W + r] = dw;
      }
      draw_pos(d, c, GCA_SLOT_GOAL, GCA_BLOCK_POS, goal.x, goal.y);   // Goal(random_pos()) :93
    }
  } else {
    for (int r = 0; r < s.W; ++r) {
      const int i = r * 32 + lane;
      const bool valid = i < s.N;
      Intruder it;
      it.is64 = false;
      if (valid) {
        spawn<FAITH, TAPE>(d, c, GCA_SLOT_RESET | (uint32_t)i, ox, oy, it);
        store_ipos<FAITH>(s, env * (size_t)s.Np + i, it.px, it.py);
        s.ivel[env * (size_t)s.Np + i] = make_float2(it.vx, it.vy);
        write_obs_intruder<FAITH>(a, env, i, it, it.px, it.py);
      }
      const uint32_t dw = __ballot_sync(FULL, valid && it.is64);
      if (lane == 0) {
        s.iflag[env * (size_t)s.W + r] = 0u;
        if constexpr (FAITH) s.if64[env * (size_t)s.W + r] = dw;
      }
    }
    if (lane == owner) draw_pos(d, c, GCA_SLOT_GOAL, GCA_BLOCK_POS, goal.x, goal.y);
  }
}

// ownship state right after reset: min speed, heading pi/4, f32 velocity (Aircraft.__init__ :269-276)
__device__ __forceinline__ void reset_ownship(const gca_config& c, float2& pos, double2& hs, double2& vel) {
  double sn, cs;
  pos = make_float2(50.0f, 50.0f);
  hs = make_double2(3.141592653589793 / 4, c.min_speed);
  gca_sincos(hs.x, &sn, &cs);
  vel = make_double2((double)(float)__dmul_rn(hs.y, cs), (double)(float)__dmul_rn(hs.y, sn));
}

template <bool FAITH, bool TAPE, int TILE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) step_kernel(const StepArgs a) {
  using R = real_t<FAITH>;
  extern __shared__ uint32_t smem_oob[];                 // [warps][TILE][W] out-of-map ballots
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const long long tile = (long long)blockIdx.x * kWarpsPerBlock + warp_in_block;
  const long long env0 = tile * TILE;
  if (env0 >= s.B) return;
  uint32_t* oob_words = smem_oob + (size_t)warp_in_block * TILE * s.W;
  const int n_tile = (int)min((long long)TILE, (long long)s.B - env0);
  const bool has_env = lane < n_tile;
  const size_t me = (size_t)(env0 + (has_env ? lane : 0));

  // ---------------------------------------------------------------- phase A: ownship, lane = env
  float2 pos = make_float2(0.f, 0.f);
  double2 hs = make_double2(0., 0.), vel = make_double2(0., 0.), goal = make_double2(0., 0.);
  int4 cnt = make_int4(0, 0, 0, 0);
  bool maxstep_hit = false;
  Draws<TAPE> d;
  if (has_env) {
    pos = s.own_pos[me];
    hs = s.own_hs[me];
    goal = s.goal[me];
    cnt = s.counters[me];
    if constexpr (TAPE) {
      d.tape = a.tape + me * (size_t)a.tape_stride;
      d.cur = a.cursor[me];
    } else {
      d.k0 = a.key0; d.k1 = a.key1; d.env = a.env_id0 + (uint32_t)me; d.tick = (uint32_t)cnt.z;
    }
    // Ownship.step(a)   PKG/SingleAircraftEnv.py:299-309 (2Env :291-301, DiscreteHER :301-311)
    double f0, f1 = 0.0;
    if (c.action_kind == GCA_ACT_CONTINUOUS2) {
      const R* act = reinterpret_cast<const R*>(a.actions) + 2 * me;
      f0 = (double)act[0];
      f1 = (double)act[1];
    } else {
      const int act = reinterpret_cast<const int*>(a.actions)[me];
      if (c.action_kind == GCA_ACT_DISCRETE9) {
        f0 = (double)(act / 3 - 1);
        f1 = (double)(act % 3 - 1);
      } else {
        f0 = (double)(act - 1);
      }
    }
    double nh, ns, sn, cs;
    draw_own_noise(d, c, nh, ns);
    double heading = __dadd_rn(hs.x, __dmul_rn(c.d_heading, f0));
    heading = __dadd_rn(heading, nh);
    double speed = c.action_kind == GCA_ACT_DISCRETE3 ? __dadd_rn(hs.y, c.speed_sigma)      // reference quirk Q16
                                                      : __dadd_rn(hs.y, __dmul_rn(c.d_speed, f1));
    const double m = c.max_speed < speed ? c.max_speed : speed;     // min(speed, max_speed)
    speed = m > c.min_speed ? m : c.min_speed;                      // max(min_speed, .)
    speed = __dadd_rn(speed, ns);                                   // noise after the clamp (Q5)
    gca_sincos(heading, &sn, &cs);
    vel = make_double2(__dmul_rn(speed, cs), __dmul_rn(speed, sn));
    hs = make_double2(heading, speed);
    pos = make_float2((float)__dadd_rn((double)pos.x, vel.x), (float)__dadd_rn((double)pos.y, vel.y));
    cnt.y += 1;                                                     // StackEnv :118
    maxstep_hit = c.max_steps > 0 && cnt.y >= c.max_steps;          // StackEnv :134-136
    s.own_pos[me] = pos;
    s.own_hs[me] = hs;
    s.own_vel[me] = vel;
    s.own_vel_f32[me] = 0;
  }

  // ---------------------------------------------------------------- phase B: intruders, lanes = intruders
  bool my_nmac = false, my_conf = false;
  int my_newconf = 0;
  for (int e = 0; e < n_tile; ++e) {
    const size_t env = (size_t)(env0 + e);
    const float ox = __shfl_sync(FULL, pos.x, e), oy = __shfl_sync(FULL, pos.y, e);
    bool stop = __shfl_sync(FULL, (int)maxstep_hit, e) != 0;
    bool nmac_hit = false, conf_any = false;
    int newconf = 0;
    for (int r = 0; r < s.W; ++r) {
      Intruder it;
      bool valid;
      uint32_t fw;
      load_round<FAITH>(s, env, r, lane, it, valid, fw);
      // intruder.position += intruder.velocity   :150   (f32 + f32, or f64 + f32 for an f64 position)
      double npx, npy;
      bool oob;
      if (FAITH && it.is64) {
        npx = __dadd_rn(it.px, (double)it.vx);
        npy = __dadd_rn(it.py, (double)it.vy);
        oob = !in_map_f64(c, npx, npy);
      } else {
        const float fx = __fadd_rn((float)it.px, it.vx), fy = __fadd_rn((float)it.py, it.vy);
        npx = (double)fx;
        npy = (double)fy;
        oob = !in_map_f32(c, fx, fy);
      }
      bool lt_sep, lt_nmac, lt_init;
      separation<FAITH>(c, ox, oy, it, npx, npy, lt_sep, lt_nmac, lt_init);   // :151
      const uint32_t b_nmac = stop ? 0u : __ballot_sync(FULL, valid && lt_sep && lt_nmac);
      const int first = b_nmac ? __ffs(b_nmac) - 1 : 31;          // first NMAC index wins (Q9)
      const bool commit = !stop && valid && lane <= first;
      const uint32_t b_conf = __ballot_sync(FULL, commit && lt_sep);
      const uint32_t b_oob = __ballot_sync(FULL, commit && oob);
      newconf += __popc(b_conf & ~fw);                            // False -> True transitions :161-163
      conf_any |= b_conf != 0u;
      // the flag write lands on the old object: a replaced intruder starts with conflict False (Q7, Q8)
      const uint32_t nfw = (fw | b_conf) & ~b_oob;
      if (lane == 0) {
        if (nfw != fw) s.iflag[env * (size_t)s.W + r] = nfw;
        oob_words[e * s.W + r] = b_oob;
      }
      if constexpr (FAITH) {
        if (b_oob && lane == 0) s.if64[env * (size_t)s.W + r] &= ~b_oob;
      }
      if (commit && !oob) store_ipos<FAITH>(s, env * (size_t)s.Np + r * 32 + lane, npx, npy);
      if (valid && !(commit && oob))
        write_obs_intruder<FAITH>(a, env, r * 32 + lane, it, commit ? npx : it.px, commit ? npy : it.py);
      if (b_nmac) {
        nmac_hit = true;
        stop = true;                                              // later intruders are not touched
      }
    }
    if (lane == e) {
      my_nmac = nmac_hit;
      my_conf = conf_any;
      my_newconf = newconf;
    }
  }
  __syncwarp();

  // ---------------------------------------------------------------- phase C: respawn, reward, lane = env
  bool done = false;
  if (has_env) {
    if (!maxstep_hit) {
      // reset_intruder() for every intruder that left the map, in index order   :153-154, :229-238
      for (int r = 0; r < s.W; ++r) {
        uint32_t w = oob_words[lane * s.W + r];
        uint32_t set64 = 0;
        while (w) {
          const int j = __ffs(w) - 1;
          w &= w - 1;
          const int i = r * 32 + j;
          Intruder it;
          spawn<FAITH, TAPE>(d, c, (uint32_t)i, pos.x, pos.y, it);
          store_ipos<FAITH>(s, me * (size_t)s.Np + i, it.px, it.py);
          s.ivel[me * (size_t)s.Np + i] = make_float2(it.vx, it.vy);
          set64 |= (it.is64 ? 1u : 0u) << j;
          write_obs_intruder<FAITH>(a, me, i, it, it.px, it.py);
        }
        if constexpr (FAITH) {
          if (set64) s.if64[me * (size_t)s.W + r] |= set64;
        }
      }
    }
    cnt.x += my_newconf;
    // _terminal_reward()   :143-184 and the variant rows of SURVEY.md 8(a)
    double reward;
    int info;
    if (maxstep_hit) {
      reward = 0.0; done = true; info = GCA_INFO_MAXSTEPS;
    } else if (my_nmac) {
      reward = c.r_nmac; done = true; info = GCA_INFO_NMAC;
    } else if (my_conf) {
      reward = c.r_conflict; info = GCA_INFO_CONFLICT;
    } else if (c.wall_kind != GCA_WALL_NONE && !in_map_f32(c, pos.x, pos.y)) {
      reward = c.r_wall; done = c.wall_kind == GCA_WALL_TERMINAL; info = GCA_INFO_WALL;
    } else {
      const double dg = dist_f64((double)pos.x, (double)pos.y, goal.x, goal.y);
      if (dg < c.goal_radius) {
        reward = c.r_goal; done = true; info = GCA_INFO_GOAL;
      } else {
        reward = c.shaped_default ? __ddiv_rn(-dg, 1200.0) : c.r_default;
        info = GCA_INFO_NONE;
      }
    }
    reinterpret_cast<R*>(a.reward)[me] = (R)reward;
    a.done[me] = done ? 1 : 0;
    a.info[me] = (uint8_t)info;
    write_obs_own<FAITH>(a, me, pos.x, pos.y, vel.x, vel.y, false, hs.x, hs.y, goal.x, goal.y);   // :115-124
  }

  // ---------------------------------------------------------------- phase D: VecEnv auto-reset
  // baselines dummy_vec_env.py:52-55: the observation handed back for a finished env is reset()'s
  uint32_t dmask = a.auto_reset ? __ballot_sync(FULL, has_env && done) : 0u;
  while (dmask) {
    const int e = __ffs(dmask) - 1;
    dmask &= dmask - 1;
    const size_t env = (size_t)(env0 + e);
    __syncwarp();
    Draws<TAPE> de = d;
    if constexpr (!TAPE) {
      de.env = __shfl_sync(FULL, d.env, e);
      de.tick = __shfl_sync(FULL, d.tick, e);
      de.k0 = a.key0; de.k1 = a.key1;
    }
    reset_env_warp<FAITH, TAPE>(a, env, lane, e, de, goal);
    if (lane == e) {
      if constexpr (TAPE) d = de;
      reset_ownship(c, pos, hs, vel);
      s.own_pos[me] = pos;
      s.own_hs[me] = hs;
      s.own_vel[me] = vel;
      s.own_vel_f32[me] = 1;
      s.goal[me] = goal;
      cnt.x = 0;
      cnt.y = 0;
      cnt.w += 1;
      write_obs_own<FAITH>(a, me, pos.x, pos.y, vel.x, vel.y, true, hs.x, hs.y, goal.x, goal.y);
    }
  }
  if (has_env) {
    cnt.z += 1;                                                     // Philox tick
    s.counters[me] = cnt;
    if constexpr (TAPE) a.cursor[me] = d.cur;
  }
}

template <bool FAITH, bool TAPE, int TILE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) reset_kernel(const StepArgs a) {
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const int lane = threadIdx.x & 31;
  const long long tile = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long env0 = tile * TILE;
  if (env0 >= s.B) return;
  const int n_tile = (int)min((long long)TILE, (long long)s.B - env0);
  const bool has_env = lane < n_tile;
  const size_t me = (size_t)(env0 + (has_env ? lane : 0));
  const bool selected = has_env && (a.mask == nullptr || a.mask[me] != 0);
  int4 cnt = make_int4(0, 0, 0, 0);
  Draws<TAPE> d;
  if (has_env) {
    cnt = s.counters[me];
    if constexpr (TAPE) {
      d.tape = a.tape + me * (size_t)a.tape_stride;
      d.cur = a.cursor[me];
    } else {
      d.k0 = a.key0; d.k1 = a.key1; d.env = a.env_id0 + (uint32_t)me; d.tick = (uint32_t)cnt.z;
    }
  }
  uint32_t rmask = __ballot_sync(FULL, selected);
  while (rmask) {
    const int e = __ffs(rmask) - 1;
    rmask &= rmask - 1;
    const size_t env = (size_t)(env0 + e);
    __syncwarp();
    Draws<TAPE> de = d;
    if constexpr (!TAPE) {
      de.env = __shfl_sync(FULL, d.env, e);
      de.tick = __shfl_sync(FULL, d.tick, e);
      de.k0 = a.key0; de.k1 = a.key1;
    }
    double2 goal = make_double2(0., 0.);
    reset_env_warp<FAITH, TAPE>(a, env, lane, e, de, goal);
    if (lane == e) {
      if constexpr (TAPE) d = de;
      float2 pos;
      double2 hs, vel;
      reset_ownship(c, pos, hs, vel);
      s.own_pos[me] = pos;
      s.own_hs[me] = hs;
      s.own_vel[me] = vel;
      s.own_vel_f32[me] = 1;
      s.goal[me] = goal;
      cnt.x = 0;
      cnt.y = 0;
      cnt.z += 1;
      s.counters[me] = cnt;
      if constexpr (TAPE) a.cursor[me] = d.cur;
      if (a.done) a.done[me] = 0;
      if (a.info) a.info[me] = 0;
      write_obs_own<FAITH>(a, me, pos.x, pos.y, vel.x, vel.y, true, hs.x, hs.y, goal.x, goal.y);
    }
  }
}

// _get_ob() of the current state   PKG/SingleAircraftEnv.py:100-126
template <bool FAITH, int TILE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) observe_kernel(const StepArgs a) {
  const DevState& s = a.s;
  const int lane = threadIdx.x & 31;
  const long long tile = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long env0 = tile * TILE;
  if (env0 >= s.B) return;
  const int n_tile = (int)min((long long)TILE, (long long)s.B - env0);
  for (int e = 0; e < n_tile; ++e) {
    const size_t env = (size_t)(env0 + e);
    for (int r = 0; r < s.W; ++r) {
      Intruder it;
      bool valid;
      uint32_t fw;
      load_round<FAITH>(s, env, r, lane, it, valid, fw);
      if (valid) write_obs_intruder<FAITH>(a, env, r * 32 + lane, it, it.px, it.py);
    }
  }
  if (lane < n_tile) {
    const size_t me = (size_t)(env0 + lane);
    const float2 pos = s.own_pos[me];
    const double2 hs = s.own_hs[me], vel = s.own_vel[me], goal = s.goal[me];
    write_obs_own<FAITH>(a, me, pos.x, pos.y, vel.x, vel.y, s.own_vel_f32[me] != 0, hs.x, hs.y, goal.x, goal.y);
  }
}

// ------------------------------------------------------------------------------ launchers
template <int TILE>
static cudaError_t launch_tile(int kind, bool faith, bool tape, const StepArgs& a, cudaStream_t st) {
  const long long tiles = ((long long)a.s.B + TILE - 1) / TILE;
  const unsigned blocks = (unsigned)((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const dim3 grid(blocks), block(kWarpsPerBlock * 32);
  const size_t smem = sizeof(uint32_t) * kWarpsPerBlock * TILE * (size_t)(a.s.W > 0 ? a.s.W : 1);
#define GCA_DISPATCH(KERNEL, SMEM)                                                       \
  do {                                                                                   \
    if (faith) {                                                                         \
      if (tape) KERNEL<true, true, TILE><<<grid, block, SMEM, st>>>(a);                  \
      else KERNEL<true, false, TILE><<<grid, block, SMEM, st>>>(a);                      \
    } else {                                                                             \
      if (tape) KERNEL<false, true, TILE><<<grid, block, SMEM, st>>>(a);                 \
      else KERNEL<false, false, TILE><<<grid, block, SMEM, st>>>(a);                     \
    }                                                                                    \
  } while (0)
  if (kind == 0) GCA_DISPATCH(step_kernel, smem);
  else if (kind == 1) GCA_DISPATCH(reset_kernel, 0);
  else {
    if (faith) observe_kernel<true, TILE><<<grid, block, 0, st>>>(a);
    else observe_kernel<false, TILE><<<grid, block, 0, st>>>(a);
  }
#undef GCA_DISPATCH
  return cudaGetLastError();
}

cudaError_t launch_step(bool faith, bool tape, int tile, const StepArgs& a, cudaStream_t st) {
  if (tile == 8) return launch_tile<8>(0, faith, tape, a, st);
  if (tile == 16) return launch_tile<16>(0, faith, tape, a, st);
  return launch_tile<32>(0, faith, tape, a, st);
}

cudaError_t launch_reset(bool faith, bool tape, const StepArgs& a, cudaStream_t st) {
  return launch_tile<32>(1, faith, tape, a, st);
}

cudaError_t launch_observe(bool faith, const StepArgs& a, cudaStream_t st) {
  return launch_tile<32>(2, faith, false, a, st);
}

}  // namespace gca
