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
// Memory pipeline.  An env's intruder row (positions, velocities, flag words: one contiguous
// 16-byte aligned record, gca_device.cuh) is brought into shared memory by ONE 1-D bulk copy
// (cp.async.bulk, the TMA engine) completing on an mbarrier.  Each warp owns a ring of
// `stages` row buffers and keeps that many envs in flight ahead of phase B, so the kernel's
// memory-level parallelism does not depend on registers or occupancy.  The grid is persistent:
// warps pull tiles from a device-side counter, which keeps every SM busy to the end and
// de-correlates the FP64 phases (A, C) of some warps from the streaming phase (B) of others.
// Every row is read once and written once; nothing is staged through global scratch.
#include <cstdio>
#include <type_traits>

#include "gca_device.cuh"
#include "gca_launch.h"

namespace gca {

// Optional per-warp phase timestamps (build with -DGCA_PHASE_TIMING; tools/phase_timing.py reads them).
#ifdef GCA_PHASE_TIMING
__device__ unsigned long long g_phase_stamps[8192 * 8];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GCA_STAMP(slot)                                                                       \
  do {                                                                                        \
    if (lane == 0 && tile < 8192) g_phase_stamps[tile * 8 + (slot)] = gtime();                \
  } while (0)
#else
#define GCA_STAMP(slot) do { } while (0)
#endif

constexpr int kWarpsPerBlock = 1;   // one warp per block: blocks spread 13-14 per SM instead of 12 or 16 warps
constexpr int kMaxStages = 8;

// one warp-round (32 intruders) of an env row held in shared or global memory
template <bool FAITH>
__device__ __forceinline__ void load_round(const DevState& s, const uint8_t* row, int r, int lane, Intr<FAITH>& it,
                                           bool& valid, uint32_t& fw) {
  const int i = r * 32 + lane;
  valid = i < s.N;
  it.px = it.py = 0;
  it.vx = it.vy = 0.0f;
  if (valid) load_intruder<FAITH>(s, row, i, it);
  fw = reinterpret_cast<const uint32_t*>(row + s.off_flag)[r];
  if constexpr (FAITH) {
    const uint32_t dw = reinterpret_cast<const uint32_t*>(row + s.off_f64)[r];
    it.is64 = (dw >> lane) & 1u;
  }
}

// reset(): PKG/SingleAircraftEnv.py:66-98 for env `env`, executed by the whole warp.
// PHILOX: lanes = intruders.  TAPE: lane `owner` replays the reference's sequential draw order.
template <bool FAITH, bool TAPE>
__device__ __forceinline__ void reset_env_warp(const StepArgs& a, size_t env, int lane, int owner, Draws<TAPE>& d,
                                               double2& goal) {
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const Derived& k = a.k;
  uint8_t* row = env_row(s, env);
  real_t<FAITH>* obase = obs_intruder_base<FAITH>(a, env);
  const float ox = 50.0f, oy = 50.0f;                     // Ownship(position=(50, 50), ...) :72-76
  if constexpr (TAPE) {
    if (lane == owner) {
      for (int r = 0; r < s.W; ++r) {
        uint32_t dw = 0;
        for (int j = 0; j < 32 && r * 32 + j < s.N; ++j) {
          const int i = r * 32 + j;
          Intr<FAITH> it;
          spawn<FAITH, TAPE>(d, c, k, GCA_SLOT_RESET | (uint32_t)i, ox, oy, it);
          store_ipos<FAITH>(row, i, it);
          store_ivel(s, row, i, it.vx, it.vy);
          dw |= (it.is64 ? 1u : 0u) << j;
          write_obs_intruder<FAITH>(a, obase, i, it);
        }
        flag_words(s, row)[r] = 0u;
        if constexpr (FAITH) f64_words(s, row)[r] = dw;
      }
      draw_pos(d, c, GCA_SLOT_GOAL, GCA_BLOCK_POS, goal.x, goal.y);   // Goal(random_pos()) :93
    }
  } else {
    for (int r = 0; r < s.W; ++r) {
      const int i = r * 32 + lane;
      const bool valid = i < s.N;
      Intr<FAITH> it;
      bool wide = false;
      if (valid) {
        spawn<FAITH, TAPE>(d, c, k, GCA_SLOT_RESET | (uint32_t)i, ox, oy, it);
        store_ipos<FAITH>(row, i, it);
        store_ivel(s, row, i, it.vx, it.vy);
        write_obs_intruder<FAITH>(a, obase, i, it);
        wide = it.is64;
      }
      const uint32_t dw = __ballot_sync(FULL, wide);
      if (lane == 0) {
        flag_words(s, row)[r] = 0u;
        if constexpr (FAITH) f64_words(s, row)[r] = dw;
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

// The reference's sequential conflict logic for one warp-round of precomputed per-lane facts
// (PKG/SingleAircraftEnv.py:153-170): first NMAC index wins and freezes every later intruder (Q9),
// the out-of-map respawn lands before the conflict test but the test uses the old object (Q7),
// the conflict flag never clears (Q8).  All lanes call it (ballots).
template <bool FAITH>
__device__ __forceinline__ void commit_round(const StepArgs& a, uint8_t* grow, const uint8_t* srow,
                                             real_t<FAITH>* obase, uint32_t* oob_row, int r, int lane,
                                             const Intr<FAITH>& it, const Intr<FAITH>& nx, bool valid, bool oob,
                                             bool lt_sep, bool lt_nmac, uint32_t fw, bool& stop, bool& nmac_hit,
                                             bool& conf_any, int& newconf, bool& oob_any) {
  const DevState& s = a.s;
  const uint32_t b_nmac = stop ? 0u : __ballot_sync(FULL, valid && lt_sep && lt_nmac);
  const int first = b_nmac ? __ffs(b_nmac) - 1 : 31;
  const bool commit = !stop && valid && lane <= first;
  const uint32_t b_conf = __ballot_sync(FULL, commit && lt_sep);
  const uint32_t b_oob = __ballot_sync(FULL, commit && oob);
  newconf += __popc(b_conf & ~fw);                                // False -> True transitions :161-163
  conf_any |= b_conf != 0u;
  oob_any |= b_oob != 0u;
  const uint32_t nfw = (fw | b_conf) & ~b_oob;                    // a replaced intruder starts with conflict False
  if (lane == 0) {
    if (nfw != fw) flag_words(s, grow)[r] = nfw;
    oob_row[r] = b_oob;
    if constexpr (FAITH) {
      if (b_oob) f64_words(s, grow)[r] = reinterpret_cast<const uint32_t*>(srow + s.off_f64)[r] & ~b_oob;
    }
  }
  if (commit && !oob) store_ipos<FAITH>(grow, r * 32 + lane, nx);
  if (valid && !(commit && oob)) write_obs_intruder<FAITH>(a, obase, r * 32 + lane, commit ? nx : it);
  if (b_nmac) {
    nmac_hit = true;
    stop = true;                                                  // later intruders are not touched
  }
}

// shared memory of one warp: [stages][row_bytes] | mbarrier[kMaxStages] | oob words [TILE][W]
__host__ __device__ inline size_t warp_smem_bytes(const DevState& s, int stages, int tile) {
  const size_t rows = (size_t)stages * (size_t)s.row_bytes;
  const size_t bars = kMaxStages * sizeof(uint64_t);
  const size_t oob = (((size_t)tile * (size_t)(s.W > 0 ? s.W : 1) * 4) + 15) & ~(size_t)15;
  return rows + bars + oob;
}

// WC > 0: the number of warp-rounds per env is the compile-time constant WC (rounds are fully
// unrolled and an env without any conflict / out-of-map event takes a short path); WC == 0: generic.
#ifndef GCA_MINB
#define GCA_MINB 16
#endif
template <bool FAITH, bool TAPE, int TILE, int WC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, GCA_MINB) step_kernel(const StepArgs a, const int n_tiles,
                                                                   const int stages) {
  using R = real_t<FAITH>;
  extern __shared__ __align__(128) uint8_t smem[];
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const Derived& k = a.k;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  uint8_t* wsm = smem + (size_t)warp_in_block * warp_smem_bytes(s, stages, TILE);
  uint8_t* ring = wsm;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)stages * s.row_bytes);
  uint32_t* oob_words = reinterpret_cast<uint32_t*>(bars + kMaxStages);
  const bool use_tma = s.N > 0;
  if (lane == 0) {
    for (int q = 0; q < stages; ++q) mbar_init(&bars[q], 1);
    mbar_fence_init();
  }
  __syncwarp();
  uint32_t phase = 0;                                     // bit q: parity the next wait on slot q expects
  const bool single_wave = (int)(gridDim.x * kWarpsPerBlock) >= n_tiles;
  bool first_pass = true;

  for (;;) {
    // ---- dynamic tile scheduler
    int tile = 0;
    if (single_wave) {                                    // every tile has its own resident warp
      tile = first_pass ? (int)(blockIdx.x * kWarpsPerBlock + warp_in_block) : n_tiles;
      first_pass = false;
    } else {
      if (lane == 0) tile = (int)atomicAdd(&s.sched[0], 1u);
      tile = __shfl_sync(FULL, tile, 0);
    }
    if (tile >= n_tiles) break;
    const long long env0 = (long long)tile * TILE;
    const int n_tile = (int)min((long long)TILE, (long long)s.B - env0);
    const bool has_env = lane < n_tile;
    const size_t me = (size_t)(env0 + (has_env ? lane : 0));

    GCA_STAMP(0);
#ifdef GCA_PHASE_TIMING
    if (lane == 0 && tile < 8192) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_phase_stamps[tile * 8 + 7] = smid;
    }
#endif
    // ---- start streaming the first rows of the tile before any arithmetic
    if (use_tma && lane == 0) {
      const int pre = n_tile < stages ? n_tile : stages;
      for (int q = 0; q < pre; ++q) {
        mbar_expect_tx(&bars[q], (uint32_t)s.row_bytes);
        tma_load_1d(ring + (size_t)q * s.row_bytes, env_row(s, (size_t)(env0 + q)), (uint32_t)s.row_bytes, &bars[q]);
      }
    }

    // -------------------------------------------------------------- phase A: ownship, lane = env
    float2 pos = make_float2(0.f, 0.f);
    double2 hs = make_double2(0., 0.), vel = make_double2(0., 0.), goal = make_double2(0., 0.);
    int4 cnt = make_int4(0, 0, 0, 0);
    bool maxstep_hit = false;
    Draws<TAPE> d;
    if constexpr (TAPE) {
      d.tape = nullptr;
      d.cur = 0;
    } else {
      d.k0 = a.key0; d.k1 = a.key1; d.env = 0; d.tick = 0;
    }
    if (has_env) {
      pos = s.own_pos[me];
      hs = s.own_hs[me];
      goal = s.goal[me];
      cnt = s.counters[me];
      if constexpr (TAPE) {
        d.tape = a.tape + me * (size_t)a.tape_stride;
        d.cur = a.cursor[me];
      } else {
        d.env = a.env_id0 + (uint32_t)me;
        d.tick = (uint32_t)cnt.z;
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

    GCA_STAMP(1);
    // -------------------------------------------------------------- phase B: lanes = intruders
    bool my_nmac = false, my_conf = false, my_oob = false;
    int my_newconf = 0;
    for (int e = 0; e < n_tile; ++e) {
      const size_t env = (size_t)(env0 + e);
      const int q = e & (stages - 1);                               // stages is a power of two
      const uint8_t* srow = ring + (size_t)q * s.row_bytes;        // shared-memory copy of the row
      uint8_t* grow = env_row(s, env);                              // where results go
      real_t<FAITH>* obase = obs_intruder_base<FAITH>(a, env);
      uint32_t* oob_row = oob_words + e * s.W;
      const float ox = __shfl_sync(FULL, pos.x, e), oy = __shfl_sync(FULL, pos.y, e);
      bool stop = __shfl_sync(FULL, (int)maxstep_hit, e) != 0;
      bool nmac_hit = false, conf_any = false, oob_any = false;
      int newconf = 0;
      if (use_tma) {
        mbar_wait(&bars[q], (phase >> q) & 1u);
        phase ^= 1u << q;
      }
      if constexpr (WC > 0) {
        Intr<FAITH> it[WC], nx[WC];
        bool valid[WC], oob[WC], lt_sep[WC], lt_nmac[WC];
        uint32_t fw[WC];
#pragma unroll
        for (int r = 0; r < WC; ++r) {
          bool lt_init;
          load_round<FAITH>(s, srow, r, lane, it[r], valid[r], fw[r]);
          nx[r] = it[r];
          oob[r] = advance<FAITH>(k, nx[r]);                        // intruder.position += velocity :150
          separation<FAITH>(k, ox, oy, nx[r], lt_sep[r], lt_nmac[r], lt_init);   // dist(drone, intruder) :151
        }
        // A round in which nobody left the map and nobody is in conflict only moves positions and
        // writes observations; the reference's sequential bookkeeping is needed for the others.
        bool quiet[WC];
#pragma unroll
        for (int r = 0; r < WC; ++r) quiet[r] = !__any_sync(FULL, valid[r] && (oob[r] || lt_sep[r]));
#pragma unroll
        for (int r = 0; r < WC; ++r) {
          if (!stop && quiet[r]) {
            if (valid[r]) {
              store_ipos<FAITH>(grow, r * 32 + lane, nx[r]);
              write_obs_intruder<FAITH>(a, obase, r * 32 + lane, nx[r]);
            }
            if (lane == 0) oob_row[r] = 0u;
          } else {
            commit_round<FAITH>(a, grow, srow, obase, oob_row, r, lane, it[r], nx[r], valid[r], oob[r], lt_sep[r],
                                lt_nmac[r], fw[r], stop, nmac_hit, conf_any, newconf, oob_any);
          }
        }
      } else {
        for (int r = 0; r < s.W; ++r) {
          Intr<FAITH> it;
          bool valid, lt_sep, lt_nmac, lt_init;
          uint32_t fw;
          load_round<FAITH>(s, srow, r, lane, it, valid, fw);
          Intr<FAITH> nx = it;
          const bool oob = advance<FAITH>(k, nx);
          separation<FAITH>(k, ox, oy, nx, lt_sep, lt_nmac, lt_init);
          commit_round<FAITH>(a, grow, srow, obase, oob_row, r, lane, it, nx, valid, oob, lt_sep, lt_nmac, fw, stop,
                              nmac_hit, conf_any, newconf, oob_any);
        }
      }
      // the slot is free again: stream the row that is `stages` envs ahead into it
      __syncwarp();
      if (use_tma && lane == 0 && e + stages < n_tile) {
        mbar_expect_tx(&bars[q], (uint32_t)s.row_bytes);
        tma_load_1d(ring + (size_t)q * s.row_bytes, env_row(s, (size_t)(env0 + e + stages)), (uint32_t)s.row_bytes,
                    &bars[q]);
      }
      if (lane == e) {
        my_nmac = nmac_hit;
        my_conf = conf_any;
        my_newconf = newconf;
        my_oob = oob_any;
      }
    }
    __syncwarp();

    GCA_STAMP(2);
    // -------------------------------------------------------------- phase C: respawn, reward, lane = env
    bool done = false;
    if (has_env) {
      uint8_t* grow = env_row(s, me);
      real_t<FAITH>* obase = obs_intruder_base<FAITH>(a, me);
      if (my_oob) {
        // reset_intruder() for every intruder that left the map, in index order   :153-154, :229-238
        for (int r = 0; r < s.W; ++r) {
          uint32_t w = oob_words[lane * s.W + r];
          uint32_t set64 = 0;
          while (w) {
            const int j = __ffs(w) - 1;
            w &= w - 1;
            const int i = r * 32 + j;
            Intr<FAITH> it;
            spawn<FAITH, TAPE>(d, c, k, (uint32_t)i, pos.x, pos.y, it);
            store_ipos<FAITH>(grow, i, it);
            store_ivel(s, grow, i, it.vx, it.vy);
            set64 |= (it.is64 ? 1u : 0u) << j;
            write_obs_intruder<FAITH>(a, obase, i, it);
          }
          if constexpr (FAITH) {
            if (set64) f64_words(s, grow)[r] |= set64;
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
      } else if (c.wall_kind != GCA_WALL_NONE && !in_map_f32(k, pos.x, pos.y)) {
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
      if (c.time_limit > 0 && cnt.y >= c.time_limit) done = true;    // gym TimeLimit of the registered ids
      reinterpret_cast<R*>(a.reward)[me] = (R)reward;
      a.done[me] = done ? 1 : 0;
      a.info[me] = (uint8_t)info;
      write_obs_own<FAITH>(a, me, pos.x, pos.y, vel.x, vel.y, false, hs.x, hs.y, goal.x, goal.y);   // :115-124
    }

    GCA_STAMP(3);
    // -------------------------------------------------------------- phase D: VecEnv auto-reset
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
    __syncwarp();
    GCA_STAMP(4);
  }

  // ---- last warp out re-arms the scheduler for the next launch
  if (!single_wave && lane == 0) {
    const unsigned total = gridDim.x * kWarpsPerBlock;
    const unsigned prev = atomicAdd(&s.sched[1], 1u);
    if (prev == total - 1) {
      s.sched[0] = 0u;
      s.sched[1] = 0u;
      __threadfence();
    }
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
  if constexpr (TAPE) {
    d.tape = nullptr;
    d.cur = 0;
  } else {
    d.k0 = a.key0; d.k1 = a.key1; d.env = 0; d.tick = 0;
  }
  if (has_env) {
    cnt = s.counters[me];
    if constexpr (TAPE) {
      d.tape = a.tape + me * (size_t)a.tape_stride;
      d.cur = a.cursor[me];
    } else {
      d.env = a.env_id0 + (uint32_t)me;
      d.tick = (uint32_t)cnt.z;
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
    const uint8_t* row = env_row(s, env);
    real_t<FAITH>* obase = obs_intruder_base<FAITH>(a, env);
    for (int r = 0; r < s.W; ++r) {
      Intr<FAITH> it;
      bool valid;
      uint32_t fw;
      load_round<FAITH>(s, row, r, lane, it, valid, fw);
      if (valid) write_obs_intruder<FAITH>(a, obase, r * 32 + lane, it);
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
namespace {
int g_num_sms = 0;

template <typename K>
cudaError_t persistent_grid(K kernel, size_t smem, int n_tiles, unsigned* blocks) {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpsPerBlock * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  const long long want = ((long long)n_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)g_num_sms * per_sm;
  *blocks = (unsigned)(want < cap ? want : cap);
  return cudaSuccess;
}

template <bool FAITH, bool TAPE, int TILE, int WC>
cudaError_t launch_step_t(const StepArgs& a, int stages, cudaStream_t st) {
  const int n_tiles = (int)(((long long)a.s.B + TILE - 1) / TILE);
  // ring depth: a power of two, never more than the tile holds nor more than ~96 KB per block
  int pow2 = 1;
  while (pow2 * 2 <= stages && pow2 * 2 <= TILE && pow2 * 2 <= kMaxStages) pow2 *= 2;
  stages = pow2;
  while (stages > 1 && kWarpsPerBlock * warp_smem_bytes(a.s, stages, TILE) > 96 * 1024) stages /= 2;
  const size_t smem = kWarpsPerBlock * warp_smem_bytes(a.s, stages, TILE);
  // grid size per (kernel, smem, tiles) is cached: the occupancy query is not free
  static unsigned cached_blocks = 0;
  static size_t cached_smem = 0;
  static int cached_tiles = -1;
  if (cached_blocks == 0 || cached_smem != smem || cached_tiles != n_tiles) {
    cudaError_t e = persistent_grid(step_kernel<FAITH, TAPE, TILE, WC>, smem, n_tiles, &cached_blocks);
    if (e != cudaSuccess) return e;
    cached_smem = smem;
    cached_tiles = n_tiles;
  }
  step_kernel<FAITH, TAPE, TILE, WC><<<cached_blocks, kWarpsPerBlock * 32, smem, st>>>(a, n_tiles, stages);
  return cudaGetLastError();
}

template <int TILE, int WC>
cudaError_t launch_step_w(bool faith, bool tape, const StepArgs& a, int stages, cudaStream_t st) {
  if (faith)
    return tape ? launch_step_t<true, true, TILE, WC>(a, stages, st) : launch_step_t<true, false, TILE, WC>(a, stages, st);
  return tape ? launch_step_t<false, true, TILE, WC>(a, stages, st) : launch_step_t<false, false, TILE, WC>(a, stages, st);
}

template <int TILE>
cudaError_t launch_step_tile(bool faith, bool tape, const StepArgs& a, int stages, cudaStream_t st) {
  switch (a.s.W) {                       // rounds per env: 1 (N <= 32), 3 (N = 65..96, the 80-intruder case) or generic
    case 1: return launch_step_w<TILE, 1>(faith, tape, a, stages, st);
    case 3: return launch_step_w<TILE, 3>(faith, tape, a, stages, st);
    default: return launch_step_w<TILE, 0>(faith, tape, a, stages, st);
  }
}
}  // namespace

cudaError_t launch_step(bool faith, bool tape, int tile, int stages, const StepArgs& a, cudaStream_t st) {
  if (tile == 8) return launch_step_tile<8>(faith, tape, a, stages, st);
  if (tile == 16) return launch_step_tile<16>(faith, tape, a, stages, st);
  return launch_step_tile<32>(faith, tape, a, stages, st);
}

#ifdef GCA_PHASE_TIMING
extern "C" int gca_debug_phase_stamps(unsigned long long* host, int count) {
  return (int)cudaMemcpyFromSymbol(host, g_phase_stamps, sizeof(unsigned long long) * (size_t)count);
}
#endif

cudaError_t launch_reset(bool faith, bool tape, const StepArgs& a, cudaStream_t st) {
  constexpr int TILE = 32;
  const long long tiles = ((long long)a.s.B + TILE - 1) / TILE;
  const unsigned blocks = (unsigned)((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (faith) {
    if (tape) reset_kernel<true, true, TILE><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
    else reset_kernel<true, false, TILE><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  } else {
    if (tape) reset_kernel<false, true, TILE><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
    else reset_kernel<false, false, TILE><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_observe(bool faith, const StepArgs& a, cudaStream_t st) {
  constexpr int TILE = 32;
  const long long tiles = ((long long)a.s.B + TILE - 1) / TILE;
  const unsigned blocks = (unsigned)((tiles + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (faith) observe_kernel<true, TILE><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  else observe_kernel<false, TILE><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gca
