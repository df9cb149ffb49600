// gca_step.cu - the fused step / reset / observe kernels (sm_100a).
//
// Mapping: lane = env.  A warp owns a tile of 32 consecutive environments, lane e owns env
// 32*tile + e for the whole step and never talks to another lane on the hot path:
//   phase A  ownship kinematics (all the f64 work: Philox + Box-Muller, sincos, clamp), 32 envs
//            per warp instruction;
//   phase B  every lane walks ITS env's intruders in index order, two per 16-byte unit: advance,
//            f32 separation on the squared distance, map test.  A unit without any event (nobody
//            left the map, nobody inside the separation radius - the overwhelmingly common case)
//            costs two shared-memory loads, ~25 FP32/integer instructions per intruder, one
//            coalesced 16-byte position store and one 16-byte observation store per intruder.
//            Anything else drops to visit_slow(), which is the reference's loop body verbatim
//            (PKG/SingleAircraftEnv.py:149-170): because a lane visits its intruders sequentially,
//            "first NMAC index wins and freezes every later intruder" (Q9), "respawn lands before
//            the conflict test but the test uses the old object" (Q7) and "the flag never clears"
//            (Q8) need no cross-lane reconstruction;
//   phase C  respawns of the intruders that left the map (their draws come in index order, like
//            the reference's), wall / goal / reward / done, observation tail, counters;
//   phase D  VecEnv auto-reset of finished envs: the warp cooperates, lanes = intruders (PHILOX).
// Memory pipeline.  State is tile-planar (gca_device.cuh): the positions and velocities of units
// [u, u+G) of a tile are two contiguous blocks of G*512 bytes, each brought into shared memory by
// ONE 1-D bulk copy (cp.async.bulk, the TMA engine; SASS UBLKCP) completing on an mbarrier.  Each
// warp owns a ring of `stages` such buffers, so memory-level parallelism does not depend on
// registers or occupancy, lane e's 16-byte shared loads are conflict-free, and every global access
// of the hot path is a full 512-byte line (state) or a private 16-byte store (observation rows).
// Every byte of state is read once and written at most once per step.
#include <cstdio>
#include <cstdlib>
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
// finer: 4 stamps per pipeline stage for every 32nd tile (64 tiles x 16 stages)
__device__ unsigned long long g_stage_stamps[64 * 16 * 4];
#define GCA_STAGE_STAMP(slot)                                                                                   \
  do {                                                                                                          \
    if (lane == 0 && (tile & 31) == 0 && tile < 2048 && st < 16) g_stage_stamps[((tile >> 5) * 16 + st) * 4 + (slot)] = gtime(); \
  } while (0)
#else
#define GCA_STAMP(slot) do { } while (0)
#define GCA_STAGE_STAMP(slot) do { } while (0)
#endif

constexpr int kWarpsPerBlock = 1;   // one warp per block: blocks spread evenly over the SMs
constexpr int kMaxStages = 8;

// shared memory of one warp: [stages][stage bytes] | observation staging [32][row] (OM = 1) |
// mbarrier[kMaxStages] | conflict words [Wd][32] | out-of-map words [Wd][32] | f64 words [Wd][32] (FAITHFUL)
__host__ __device__ inline size_t stage_bytes(bool faith, int G) { return (size_t)G * 512u * (faith ? 3u : 2u); }
// staging row of one lane: the 2G intruders of a stage x 16 bytes of observation entries, plus 16 bytes so that
// the row stride is an odd multiple of 16 (conflict-free 16-byte shared accesses across a quarter warp)
__host__ __device__ constexpr uint32_t obs_row_bytes(int G) { return 32u * (uint32_t)G + 16u; }
__host__ __device__ inline size_t warp_smem_bytes(const DevState& s, bool faith, int G, int stages, int om) {
  return (size_t)stages * stage_bytes(faith, G) + (om ? 32u * obs_row_bytes(G) : 0u) + kMaxStages * sizeof(uint64_t) +
         (size_t)s.Wd * 128u * (faith ? 3u : 2u);
}

// reset(): PKG/SingleAircraftEnv.py:66-98 for env `env`, executed by the whole warp.
// PHILOX: lanes = intruders.  TAPE: lane `owner` replays the reference's sequential draw order.
template <bool FAITH, bool TAPE>
__device__ __forceinline__ void reset_env_warp(const StepArgs& a, size_t env, int lane, int owner, Draws<TAPE>& d,
                                               double2& goal) {
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const Derived& k = a.k;
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
          store_ipos<FAITH>(s, env, i, it);
          store_ivel(s, env, i, it.vx, it.vy);
          dw |= (it.is64 ? 1u : 0u) << j;
          write_obs_intruder<FAITH>(a, obase, i, it);
        }
        s.cflag[flag_index(s, env, r)] = 0u;
        if constexpr (FAITH) s.dflag[flag_index(s, env, r)] = dw;
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
        store_ipos<FAITH>(s, env, i, it);
        store_ivel(s, env, i, it.vx, it.vy);
        write_obs_intruder<FAITH>(a, obase, i, it);
        wide = it.is64;
      }
      const uint32_t dw = __ballot_sync(FULL, wide);
      if (lane == 0) {
        s.cflag[flag_index(s, env, r)] = 0u;
        if constexpr (FAITH) s.dflag[flag_index(s, env, r)] = dw;
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

// G  : units (intruder pairs) per pipeline stage.
// OM : 1 = the observation is GCA_OBS_VECTOR with the one-correction division exact (the registered
//      ids' layout; FAST only): the hot path writes it without run-time layout tests; 0 = generic.
#ifndef GCA_MINB
#define GCA_MINB 14
#endif
template <bool FAITH, bool TAPE, int G, int OM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, GCA_MINB) step_kernel(const StepArgs a, const int stages) {
  using R = real_t<FAITH>;
  static_assert(!(FAITH && OM), "the specialised observation writer is FAST only");
  extern __shared__ __align__(128) uint8_t smem[];
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const Derived& k = a.k;
  constexpr uint32_t kPosUnits = FAITH ? 2u : 1u;                // 16-byte position units per intruder pair
  constexpr uint32_t kStagePos = G * 512u * kPosUnits, kStageBytes = kStagePos + G * 512u;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  static_assert(G == 1 || G == 2 || G == 4 || G == 8, "2G lanes write out one env's entries of a stage");
  constexpr uint32_t kObsRow = obs_row_bytes(G), kObsStage = 32u * kObsRow;
  uint8_t* wsm = smem + (size_t)warp_in_block * warp_smem_bytes(s, FAITH, G, stages, OM);
  uint8_t* ring = wsm;
  uint8_t* stg = wsm + (size_t)stages * kStageBytes;                            // observation staging, row e = lane e's env
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + (size_t)stages * kStageBytes + (OM ? kObsStage : 0u));
  uint32_t* cfw = reinterpret_cast<uint32_t*>(bars + kMaxStages) + lane;       // this lane's word w at [w * 32]
  uint32_t* oobw = cfw + s.Wd * 32;
  uint32_t* dfw = oobw + s.Wd * 32;                                            // FAITHFUL only
  if (lane == 0) {
    for (int q = 0; q < stages; ++q) mbar_init(&bars[q], 1);
    mbar_fence_init();
  }
  __syncwarp();
  uint32_t phase = 0;                                     // bit q: parity the next wait on slot q expects
  const int n_tiles = s.T;
  const int n_pairs = s.N >> 1;                           // units holding two intruders
  const int n_st = (s.U + G - 1) / G;                     // pipeline stages per tile
  const bool single_wave = (int)(gridDim.x * kWarpsPerBlock) >= n_tiles;
  bool first_pass = true;

  for (;;) {
    // ---- tile scheduler
    int tile = 0;
    if (single_wave) {                                    // every tile has its own resident warp
      tile = first_pass ? (int)(blockIdx.x * kWarpsPerBlock + warp_in_block) : n_tiles;
      first_pass = false;
    } else {
      if (lane == 0) tile = (int)atomicAdd(&s.sched[0], 1u);
      tile = __shfl_sync(FULL, tile, 0);
    }
    if (tile >= n_tiles) break;
    const size_t env0 = (size_t)tile * 32;
    const bool has_env = env0 + lane < (size_t)s.B;
    const size_t me = has_env ? env0 + lane : env0;
    const uint8_t* tpos = s.ipos + (size_t)tile * s.U * (512u * kPosUnits);   // this tile's planes
    const uint8_t* tvel = s.ivel + (size_t)tile * s.U * 512u;

    GCA_STAMP(0);
#ifdef GCA_PHASE_TIMING
    if (lane == 0 && tile < 8192) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_phase_stamps[tile * 8 + 7] = smid;
    }
#endif
    // ---- start streaming the first stages of the tile before any arithmetic
    auto issue_stage = [&](int q, int st) {
      const int u0 = st * G;
      const uint32_t nu = (uint32_t)min(G, s.U - u0);
      uint8_t* dst = ring + (size_t)q * kStageBytes;
      mbar_expect_tx(&bars[q], nu * 512u * (kPosUnits + 1u));
      tma_load_1d(dst, tpos + (size_t)u0 * (512u * kPosUnits), nu * 512u * kPosUnits, &bars[q]);
      tma_load_1d(dst + kStagePos, tvel + (size_t)u0 * 512u, nu * 512u, &bars[q]);
    };
    if (lane == 0) {
      const int pre = n_st < stages ? n_st : stages;
      for (int q = 0; q < pre; ++q) issue_stage(q, q);
    }
    // this lane's flag words -> shared memory (each lane only ever touches its own column)
    for (int w = 0; w < s.W; ++w) {
      const size_t fi = ((size_t)tile * s.Wd + w) * 32 + lane;
      cfw[w * 32] = s.cflag[fi];
      oobw[w * 32] = 0u;
      if constexpr (FAITH) dfw[w * 32] = s.dflag[fi];
    }

    // -------------------------------------------------------------- phase A: ownship
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
      // The rest of the tile's planes start moving from HBM into L2 now, behind this
      // lane's own state loads, so that the DRAM channels work through phase A (FP64-bound, no traffic of its own).
      if (!(a.debug_skip & 8) && n_st > stages) {
        const uint32_t total_pos = (uint32_t)s.U * 512u * kPosUnits, total_vel = (uint32_t)s.U * 512u;
        const uint32_t done_pos = (uint32_t)stages * kStagePos, done_vel = (uint32_t)stages * G * 512u;
        // 32 lanes split the remainder in 512-byte aligned pieces
        const uint32_t rem_pos = total_pos - done_pos, rem_vel = total_vel - done_vel;
        const uint32_t piece_pos = ((rem_pos / 32u) + 511u) & ~511u, piece_vel = ((rem_vel / 32u) + 511u) & ~511u;
        const uint32_t o_pos = lane * piece_pos, o_vel = lane * piece_vel;
        if (o_pos < rem_pos) tma_prefetch_l2(tpos + done_pos + o_pos, min(piece_pos, rem_pos - o_pos));
        if (o_vel < rem_vel) tma_prefetch_l2(tvel + done_vel + o_vel, min(piece_vel, rem_vel - o_vel));
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
    // -------------------------------------------------------------- phase B: this lane's intruders, in index order
    bool alive = has_env && !maxstep_hit;   // false once the reference's loop has returned (NMAC, Q9) or never ran (max steps)
    bool nmac = false, conf = false;
    bool dirty = false;                     // a flag word or an out-of-map word of this env changed
    int newconf = 0;                        // False -> True transitions of Aircraft.conflict   :161-163
    R* obase = obs_intruder_base<FAITH>(a, me);
    uint8_t* gpos = s.ipos + ((size_t)tile * s.U * kPosUnits * 32 + lane) * 16;       // unit u at + u * 512 * kPosUnits
    const float ox = pos.x, oy = pos.y;
    const uint32_t wbits = __float_as_uint(k.win_w), hbits = __float_as_uint(k.win_h);

    for (int st = 0; st < n_st; ++st) {
      const int q = st & (stages - 1);                              // stages is a power of two
      GCA_STAGE_STAMP(0);
      mbar_wait(&bars[q], (phase >> q) & 1u);
      phase ^= 1u << q;
      GCA_STAGE_STAMP(1);
      const uint8_t* sp = ring + (size_t)q * kStageBytes + lane * 16;   // unit g of this stage at + g * 512 (* kPosUnits)
      const uint8_t* sv = sp + kStagePos;
      const int u0 = st * G;
      uint32_t gone = 0;                                            // bit j: intruder 2*u0 + j left the map

      // One iteration of the reference's loop body for intruder i (PKG/SingleAircraftEnv.py:149-170), verbatim;
      // taken when the streamlined path below saw a conflict in the unit or the loop has already returned.
      auto visit_slow = [&](int i, const Intr<FAITH>& old) {
        if (!alive) {                                               // not touched this step: observed where it was
          write_obs_intruder<FAITH>(a, obase, i, old);
          return;
        }
        Intr<FAITH> nx = old;
        const bool oob = advance<FAITH>(k, nx);                     // intruder.position += velocity :150, map test :153
        bool lt_sep, lt_nmac, lt_init;
        separation<FAITH>(k, ox, oy, nx, lt_sep, lt_nmac, lt_init); // dist(drone, intruder) :151
        if (oob) {                                                  // replaced by reset_intruder() in phase C :153-154
          gone |= 1u << (i - 2 * u0);
        } else {
          store_ipos<FAITH>(s, me, i, nx);
          write_obs_intruder<FAITH>(a, obase, i, nx);
        }
        if (lt_sep) {                                               // the old object's distance and flag (Q7) :157-163
          conf = true;
          const uint32_t bit = 1u << (i & 31), f = cfw[(i >> 5) * 32];
          if (!(f & bit)) {
            newconf += 1;
            cfw[(i >> 5) * 32] = f | bit;
            dirty = true;
          }
          if (lt_nmac) {                                            // return inside the loop :169-170
            nmac = true;
            alive = false;
          }
        }
      };

      // the exact path for unit g (dynamic index): both intruders through the reference's loop body
      auto careful_unit = [&](int g) {
        const int i0 = 2 * (u0 + g);
        const float4 vv = *reinterpret_cast<const float4*>(sv + g * 512);
        Intr<FAITH> o0, o1;
        o0.vx = vv.x; o0.vy = vv.y; o1.vx = vv.z; o1.vy = vv.w;
        if constexpr (FAITH) {
          const double2 p0 = *reinterpret_cast<const double2*>(sp + (2 * g) * 512);
          const double2 p1 = *reinterpret_cast<const double2*>(sp + (2 * g + 1) * 512);
          const uint32_t dw = dfw[(i0 >> 5) * 32];
          o0.px = p0.x; o0.py = p0.y; o1.px = p1.x; o1.py = p1.y;
          o0.is64 = (dw >> (i0 & 31)) & 1u;
          o1.is64 = (dw >> ((i0 + 1) & 31)) & 1u;
        } else {
          const float4 p = *reinterpret_cast<const float4*>(sp + g * 512);
          o0.px = p.x; o0.py = p.y; o1.px = p.z; o1.py = p.w;
        }
        visit_slow(i0, o0);
        visit_slow(i0 + 1, o1);
      };

      bool streamlined = false;
      if (has_env) {
        const int nu = min(G, n_pairs - u0);                        // full units (two intruders) of this stage
        if (nu == G) {
          // The streamlined path: all 2G intruders of the stage at once, straight-line.  It applies when the
          // reference's loop is still running for this env and no intruder of the stage is inside the separation
          // radius.  An intruder that leaves the map is only recorded (`gone`): its slot is refilled in phase C,
          // which also rewrites its position and observation entries.
          bool ev = !alive;
          uint32_t g_gone = 0;
          if constexpr (FAITH) {
            Intr<FAITH> n[2 * G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const int i0 = 2 * (u0 + g);
              const float4 vv = *reinterpret_cast<const float4*>(sv + g * 512);
              const double2 p0 = *reinterpret_cast<const double2*>(sp + (2 * g) * 512);
              const double2 p1 = *reinterpret_cast<const double2*>(sp + (2 * g + 1) * 512);
              const uint32_t dw = dfw[(i0 >> 5) * 32];
              Intr<FAITH>& n0 = n[2 * g];
              Intr<FAITH>& n1 = n[2 * g + 1];
              n0.vx = vv.x; n0.vy = vv.y; n1.vx = vv.z; n1.vy = vv.w;
              n0.px = p0.x; n0.py = p0.y; n1.px = p1.x; n1.py = p1.y;
              n0.is64 = (dw >> (i0 & 31)) & 1u;
              n1.is64 = (dw >> ((i0 + 1) & 31)) & 1u;
              const bool oob0 = advance<FAITH>(k, n0), oob1 = advance<FAITH>(k, n1);
              bool sep0, sep1, t0, t1;
              separation<FAITH>(k, ox, oy, n0, sep0, t0, t1);
              separation<FAITH>(k, ox, oy, n1, sep1, t0, t1);
              ev |= sep0 | sep1;
              g_gone |= ((oob0 ? 1u : 0u) | (oob1 ? 2u : 0u)) << (2 * g);
            }
            if (!ev) {
              streamlined = true;
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const int u = u0 + g;
                *reinterpret_cast<double2*>(gpos + (size_t)(2 * u) * 512) = make_double2(n[2 * g].px, n[2 * g].py);
                *reinterpret_cast<double2*>(gpos + (size_t)(2 * u + 1) * 512) = make_double2(n[2 * g + 1].px, n[2 * g + 1].py);
                write_obs_intruder<FAITH>(a, obase, 2 * u, n[2 * g]);
                write_obs_intruder<FAITH>(a, obase, 2 * u + 1, n[2 * g + 1]);
              }
            }
          } else {
            float4 np[G], vv[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const float4 p = *reinterpret_cast<const float4*>(sp + g * 512);
              vv[g] = *reinterpret_cast<const float4*>(sv + g * 512);
              np[g] = make_float4(__fadd_rn(p.x, vv[g].x), __fadd_rn(p.y, vv[g].y),      // position += velocity :150
                                  __fadd_rn(p.z, vv[g].z), __fadd_rn(p.w, vv[g].w));
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
              // 0 <= x <= W on f32 bit patterns: a non-negative float is <= W iff its pattern is (as unsigned);
              // negatives and NaN have larger patterns (FAST positions are never -0.0, see gca_set_state).
              const bool oob0 = (__float_as_uint(np[g].x) > wbits) | (__float_as_uint(np[g].y) > hbits);
              const bool oob1 = (__float_as_uint(np[g].z) > wbits) | (__float_as_uint(np[g].w) > hbits);
              g_gone |= ((oob0 ? 1u : 0u) | (oob1 ? 2u : 0u)) << (2 * g);
              ev |= (dist2_f32(ox, oy, np[g].x, np[g].y) < k.sep2_f) | (dist2_f32(ox, oy, np[g].z, np[g].w) < k.sep2_f);
            }
            if (!ev) {
              streamlined = true;
#pragma unroll
              for (int g = 0; g < G; ++g)
                if (!(a.debug_skip & 2)) *reinterpret_cast<float4*>(gpos + (size_t)(u0 + g) * 512) = np[g];
              if constexpr (OM == 1) {
                // the observation entries go to this lane's staging row; the warp writes the rows out below
                float4* row = reinterpret_cast<float4*>(stg + lane * kObsRow);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                  row[2 * g] = obs_intruder_vec(k, np[g].x, np[g].y, vv[g].x, vv[g].y);
                  row[2 * g + 1] = obs_intruder_vec(k, np[g].z, np[g].w, vv[g].z, vv[g].w);
                }
              } else {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                  const int u = u0 + g;
                  Intr<FAITH> n0, n1;
                  n0.px = np[g].x; n0.py = np[g].y; n0.vx = vv[g].x; n0.vy = vv[g].y;
                  n1.px = np[g].z; n1.py = np[g].w; n1.vx = vv[g].z; n1.vy = vv[g].w;
                  write_obs_intruder<FAITH>(a, obase, 2 * u, n0);
                  write_obs_intruder<FAITH>(a, obase, 2 * u + 1, n1);
                }
              }
            }
          }
          if (streamlined) gone = g_gone;
        }
        if (!streamlined) {
#pragma unroll 1
          for (int g = 0; g < nu; ++g) careful_unit(g);
          if ((s.N & 1) && st == n_st - 1) {                        // odd N: the last unit holds one intruder
            const int g = s.U - 1 - u0, i = s.N - 1;
            const float2 vv = *reinterpret_cast<const float2*>(sv + g * 512);
            Intr<FAITH> o;
            o.vx = vv.x; o.vy = vv.y;
            if constexpr (FAITH) {
              const double2 p = *reinterpret_cast<const double2*>(sp + (2 * g) * 512);
              o.px = p.x; o.py = p.y;
              o.is64 = (dfw[(i >> 5) * 32] >> (i & 31)) & 1u;
            } else {
              const float2 p = *reinterpret_cast<const float2*>(sp + g * 512);
              o.px = p.x; o.py = p.y;
            }
            visit_slow(i, o);
          }
        }
        if (gone) {                                                 // a stage never straddles a 32-intruder word (G | 16)
          oobw[((2 * u0) >> 5) * 32] |= gone << ((2 * u0) & 31);
          dirty = true;
        }
      }
      // the slot is free again: stream the stage that is `stages` ahead into it
      __syncwarp();
      GCA_STAGE_STAMP(2);
      if (lane == 0 && st + stages < n_st) issue_stage(q, st + stages);
      if constexpr (OM == 1) {
        // Write-out of the staged observation entries, transposed: 2G consecutive lanes store the 2G x 16 contiguous
        // bytes of ONE env's row, so every store instruction covers whole 32-byte sectors of 16/G rows instead of
        // 32 half sectors 1312 bytes apart.  Rows of lanes that took the exact path were written there.
        const uint32_t staged = __ballot_sync(FULL, streamlined);
        if (staged && !(a.debug_skip & 1)) {
          constexpr int kLanesPerEnv = 2 * G, kEnvsPerStore = 32 / kLanesPerEnv;
          const int sub = lane / kLanesPerEnv, chunk = lane % kLanesPerEnv;
          const uint8_t* src = stg + sub * kObsRow + chunk * 16;
          float* dst = reinterpret_cast<float*>(a.obs) + (env0 + sub) * (size_t)a.D + 4 * (size_t)(2 * u0 + chunk);
          const size_t dstep = (size_t)kEnvsPerStore * a.D;
#pragma unroll
          for (int it = 0; it < kLanesPerEnv; ++it) {
            const float4 val = *reinterpret_cast<const float4*>(src + it * kEnvsPerStore * kObsRow);
            if (a.debug_skip & 4) {   // experiment: same bytes, written as one contiguous 1024*G-byte block per stage
              *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.obs) + env0 * (size_t)a.D * 4 + (size_t)st * (1024 * G) + it * 512 + lane * 16) = val;
            } else
            if ((staged >> (it * kEnvsPerStore + sub)) & 1u) *reinterpret_cast<float4*>(dst + it * dstep) = val;
          }
          __syncwarp();                                             // staging rows are rewritten by the next stage
        }
      }
      GCA_STAGE_STAMP(3);
    }

    GCA_STAMP(2);
    // -------------------------------------------------------------- phase C: respawn, reward
    bool done = false;
    if (has_env) {
      if (dirty) {
        // reset_intruder() for every intruder that left the map, in index order   :153-154, :229-238
        for (int w = 0; w < s.W; ++w) {
          const uint32_t gone = oobw[w * 32];
          uint32_t rest = gone, set64 = 0;
          while (rest) {
            const int j = __ffs(rest) - 1;
            rest &= rest - 1;
            const int i = w * 32 + j;
            Intr<FAITH> it;
            spawn<FAITH, TAPE>(d, c, k, (uint32_t)i, pos.x, pos.y, it);
            store_ipos<FAITH>(s, me, i, it);
            store_ivel(s, me, i, it.vx, it.vy);
            set64 |= (it.is64 ? 1u : 0u) << j;
            write_obs_intruder<FAITH>(a, obase, i, it);
          }
          const size_t fi = ((size_t)tile * s.Wd + w) * 32 + lane;
          s.cflag[fi] = cfw[w * 32] & ~gone;                          // a replaced intruder starts with conflict False
          if constexpr (FAITH) {
            if (gone) s.dflag[fi] = (dfw[w * 32] & ~gone) | set64;
          }
        }
      }
      cnt.x += newconf;
      // _terminal_reward()   :143-184 and the variant rows of SURVEY.md 8(a)
      double reward;
      int info;
      if (maxstep_hit) {
        reward = 0.0; done = true; info = GCA_INFO_MAXSTEPS;
      } else if (nmac) {
        reward = c.r_nmac; done = true; info = GCA_INFO_NMAC;
      } else if (conf) {
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
      const size_t env = env0 + e;
      __syncwarp();                                                   // lane e's phase B/C stores come first
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

template <bool FAITH, bool TAPE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) reset_kernel(const StepArgs a) {
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const int lane = threadIdx.x & 31;
  const long long tile = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long env0 = tile * 32;
  if (env0 >= s.B) return;
  const int n_tile = (int)min(32LL, (long long)s.B - env0);
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

// _get_ob() of the current state   PKG/SingleAircraftEnv.py:100-126; lane = env, plane reads are coalesced
template <bool FAITH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) observe_kernel(const StepArgs a) {
  const DevState& s = a.s;
  const long long me_ll = ((long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * 32 + (threadIdx.x & 31);
  if (me_ll >= s.B) return;
  const size_t me = (size_t)me_ll;
  real_t<FAITH>* obase = obs_intruder_base<FAITH>(a, me);
  for (int i = 0; i < s.N; ++i) {
    Intr<FAITH> it;
    load_intruder<FAITH>(s, me, i, it);
    write_obs_intruder<FAITH>(a, obase, i, it);
  }
  const float2 pos = s.own_pos[me];
  const double2 hs = s.own_hs[me], vel = s.own_vel[me], goal = s.goal[me];
  write_obs_own<FAITH>(a, me, pos.x, pos.y, vel.x, vel.y, s.own_vel_f32[me] != 0, hs.x, hs.y, goal.x, goal.y);
}

// ------------------------------------------------------------------------------ launchers
namespace {
int g_num_sms = 0;

// Grid and dynamic shared memory of the persistent step kernel.  When every tile can have its own resident
// block, the request is padded so that exactly ceil(tiles / SMs) blocks fit on an SM: the hardware block
// scheduler then spreads the single wave evenly instead of packing some SMs to the register limit.
template <typename K>
cudaError_t persistent_grid(K kernel, size_t need, int n_tiles, unsigned* blocks, size_t* smem) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (g_num_sms == 0) {
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  int sm_smem = 0, reserved = 0, max_optin = 0;
  cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpsPerBlock * 32, need);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  const long long want = ((long long)n_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)g_num_sms * per_sm;
  *smem = need;
  static const int forced = std::getenv("GCA_BLOCKS_PER_SM") ? std::atoi(std::getenv("GCA_BLOCKS_PER_SM")) : 0;   // tuning knob
  if (want <= cap || forced > 0) {                        // single wave: balance it
    const int target = forced > 0 ? forced : (int)((want + g_num_sms - 1) / g_num_sms);
    if (target < per_sm && sm_smem > 0) {
      size_t padded = ((size_t)sm_smem / (size_t)target - (size_t)reserved) & ~(size_t)127;
      if (padded > (size_t)max_optin) padded = (size_t)max_optin;
      int chk = 0;
      if (padded > need && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&chk, kernel, kWarpsPerBlock * 32, padded) == cudaSuccess &&
          chk == target)
        *smem = padded, per_sm = target;
    }
  }
  const long long cap2 = (long long)g_num_sms * per_sm;
  *blocks = (unsigned)(want < cap2 ? want : cap2);
  return cudaSuccess;
}

template <bool FAITH, bool TAPE, int G, int OM>
cudaError_t launch_step_t(const StepArgs& a, int stages, cudaStream_t st) {
  const int n_tiles = a.s.T;
  const int n_st = (a.s.U + G - 1) / G;
  // ring depth: a power of two, no deeper than the tile has stages, at most ~96 KB per block
  int pow2 = 1;
  while (pow2 * 2 <= stages && pow2 < n_st && pow2 * 2 <= kMaxStages) pow2 *= 2;
  stages = pow2;
  while (stages > 1 && kWarpsPerBlock * warp_smem_bytes(a.s, FAITH, G, stages, OM) > 96 * 1024) stages /= 2;
  const size_t need = kWarpsPerBlock * warp_smem_bytes(a.s, FAITH, G, stages, OM);
  // grid size per (kernel, smem, tiles) is cached: the occupancy query is not free
  static unsigned cached_blocks = 0;
  static size_t cached_need = 0, cached_smem = 0;
  static int cached_tiles = -1;
  if (cached_blocks == 0 || cached_need != need || cached_tiles != n_tiles) {
    cudaError_t e = persistent_grid(step_kernel<FAITH, TAPE, G, OM>, need, n_tiles, &cached_blocks, &cached_smem);
    if (e != cudaSuccess) return e;
    cached_need = need;
    cached_tiles = n_tiles;
  }
  const size_t smem = cached_smem;
  step_kernel<FAITH, TAPE, G, OM><<<cached_blocks, kWarpsPerBlock * 32, smem, st>>>(a, stages);
  return cudaGetLastError();
}

template <int G>
cudaError_t launch_step_g(bool faith, bool tape, const StepArgs& a, int stages, cudaStream_t st) {
  if (faith)
    return tape ? launch_step_t<true, true, G, 0>(a, stages, st) : launch_step_t<true, false, G, 0>(a, stages, st);
  const bool vec = a.cfg.obs_kind == GCA_OBS_VECTOR && a.k.div1_ok;
  if (vec) return tape ? launch_step_t<false, true, G, 1>(a, stages, st) : launch_step_t<false, false, G, 1>(a, stages, st);
  return tape ? launch_step_t<false, true, G, 0>(a, stages, st) : launch_step_t<false, false, G, 0>(a, stages, st);
}
}  // namespace

cudaError_t launch_step(bool faith, bool tape, int group, int stages, const StepArgs& a, cudaStream_t st) {
  if (group == 2) return launch_step_g<2>(faith, tape, a, stages, st);
  return launch_step_g<4>(faith, tape, a, stages, st);
}

#ifdef GCA_PHASE_TIMING
extern "C" int gca_debug_phase_stamps(unsigned long long* host, int count) {
  return (int)cudaMemcpyFromSymbol(host, g_phase_stamps, sizeof(unsigned long long) * (size_t)count);
}
extern "C" int gca_debug_stage_stamps(unsigned long long* host, int count) {
  return (int)cudaMemcpyFromSymbol(host, g_stage_stamps, sizeof(unsigned long long) * (size_t)count);
}
#endif

cudaError_t launch_reset(bool faith, bool tape, const StepArgs& a, cudaStream_t st) {
  const unsigned blocks = (unsigned)((a.s.T + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (faith) {
    if (tape) reset_kernel<true, true><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
    else reset_kernel<true, false><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  } else {
    if (tape) reset_kernel<false, true><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
    else reset_kernel<false, false><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_observe(bool faith, const StepArgs& a, cudaStream_t st) {
  const unsigned blocks = (unsigned)((a.s.T + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (faith) observe_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  else observe_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gca
