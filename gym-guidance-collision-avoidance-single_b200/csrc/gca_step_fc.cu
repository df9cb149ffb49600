// gca_step_fc.cu - the forecast step: one step = the head kernel + the streaming pass, running CONCURRENTLY, and
// nothing after them (sm_100a).  PHILOX handles with intruders whose reward does not need the step's nearest distance
// and whose intruders do not turn (every registered id and BASELINE config; the others keep gca_step.cu's
// own-role + finish pair).
//
// Why.  A step is: ownship update -> 80 intruders advance (the 220 MB stream) -> the reference's sequential loop
// semantics (respawn what left the map, conflict flags, first NMAC wins, reward, done, auto-reset).  The third part
// needs every intruder of the env, so as a kernel behind the stream it is a 10-15 us latency chain at the end of
// every step, whatever its width (gca_step.cu's step_finish_kernel).  Here nothing of it waits for the stream:
//   * departures are FORECAST one step ahead: the pass that stores a position also makes the f32 sum and the map
//     test the next step will make on it (same operands, same rounding - the forecast is the advance), so the
//     head knows at the START of a step which intruders leave in it.  It spawns their successors (one thread per
//     spawn) while the stream runs; the stream stores nothing for those intruders.
//   * a conflict / NMAC needs an intruder inside minimum_separation.  The stream also keeps the smallest squared
//     distance of the state it writes; with the ownship's own displacement (known once the action is applied) and
//     the bound on an intruder's displacement per step, the triangle inequality tells the head which envs CANNOT
//     see a conflict in this step (~99.6 % at the reference's config).  For those the reward is settled by the
//     ownship alone (wall / goal / default / max steps, PKG/SingleAircraftEnv.py:173-183).  The others ("hot") are
//     advanced by the head itself, a warp per env with lanes = intruders, which replays PKG/SingleAircraftEnv.py
//     :149-170 exactly (first NMAC index wins, later intruders untouched (Q9), a replaced intruder is tested with its
//     old distance (Q7), flags never clear (Q8)); the stream skips them.
//   * an env that finishes without a conflict being possible (wall / goal / max steps / TimeLimit) is reset by the
//     head (VecEnv auto-reset, dummy_vec_env.py:52-55) and skipped by the stream as well.
// The head owns a few groups of 128 envs per block, publishes all their ownship records first (phase A: the stream
// waits for nothing else) and then does the per-group work (phase B).  Records reach the stream through
// DevState::own_b, stamped with the step count as before.
#include <climits>
#include <cstdlib>

#include "gca_launch.h"
#include "gca_step_common.cuh"

namespace gca {

constexpr int kHeadThreads = 128;
constexpr int kHeadMaxGroups = 16;       // groups of 128 envs one head block may own (the grid is sized accordingly)
constexpr int kJobCap = 768;             // respawn records per group; beyond that the env's own lane spawns in place
constexpr uint32_t kInfBits = 0x7f800000u;
enum { CLS_NORMAL = 0, CLS_HOT = 1, CLS_RESET = 2, CLS_NONE = 3, CLS_RUNS = 4 };

template <bool FAITH>
__device__ __forceinline__ bool advance_rt(const Derived& k, Intr<FAITH>& it) {
  return k.has_drift ? advance<FAITH, true>(k, it) : advance<FAITH, false>(k, it);
}

// what _terminal_reward() returns when no intruder event outranks it   PKG/SingleAircraftEnv.py:173-183 and the
// variant rows of SURVEY.md 8(a)
struct Settled {
  double reward;
  int info;
  bool done;
};
__device__ __forceinline__ Settled settle_own(const gca_config& c, const Derived& k, bool maxstep_hit, float2 pos, double2 goal) {
  Settled r;
  r.done = false;
  if (maxstep_hit) {
    r.reward = 0.0; r.done = true; r.info = GCA_INFO_MAXSTEPS;
  } else if (c.wall_kind != GCA_WALL_NONE && !in_map_f32(k, pos.x, pos.y)) {
    r.reward = c.r_wall; r.done = c.wall_kind == GCA_WALL_TERMINAL; r.info = GCA_INFO_WALL;
  } else {
    const double dg = dist_f64((double)pos.x, (double)pos.y, goal.x, goal.y);
    if (dg < c.goal_radius) {
      r.reward = c.r_goal; r.done = true; r.info = GCA_INFO_GOAL;
    } else {
      r.reward = c.shaped_default ? ddiv_prepared(k, -dg, k.dv_shape, k.rc_shape) : c.r_default;
      r.info = GCA_INFO_NONE;
    }
  }
  return r;
}

__device__ __forceinline__ uint32_t slot_of(uint32_t tick) { return tick % 3u; }
__device__ __forceinline__ uint32_t slot_next(uint32_t slot) { return slot == 2u ? 0u : slot + 1u; }

// ------------------------------------------------------------------------------ phase A: thread = env
// Ownship.step(a)   PKG/SingleAircraftEnv.py:299-309 (2Env :291-301, DiscreteHER :301-311), the classification, the
// record, and everything of the step that the ownship alone decides.
template <bool FAITH>
__device__ __forceinline__ int head_own(const StepArgs& a, const size_t me, const uint32_t stamp) {
  using R = real_t<FAITH>;
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const Derived& k = a.k;
  const float2 pos0 = s.own_pos[me];
  double2 hs = s.own_hs[me];
  int4 cnt = s.counters[me];
  double2 goal = s.goal[me];
  const float vmax = s.fc_vmax[me];
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
  const uint32_t slot = slot_of((uint32_t)cnt.z);
  const uint32_t near_bits = s.fc_near[(size_t)slot * ((size_t)s.T * 32) + me];
  Draws<false> d = make_draws<false>(a, me, (uint32_t)cnt.z);
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
  double2 vel = make_double2(__dmul_rn(speed, cs), __dmul_rn(speed, sn));
  hs = make_double2(heading, speed);
  float2 pos = make_float2((float)__dadd_rn((double)pos0.x, vel.x), (float)__dadd_rn((double)pos0.y, vel.y));
  cnt.y += 1;                                                     // StackEnv :118
  const bool maxstep_hit = c.max_steps > 0 && cnt.y >= c.max_steps;   // StackEnv :134-136: the intruder loop never runs
  const bool runs = !maxstep_hit;
  // Can any intruder be inside minimum_separation after this step?  Every intruder was at least sqrt(near) away from
  // the old ownship position; the ownship moved by |pos - pos0|, an intruder moves by at most vmax.  The slack covers
  // the f32 rounding of the distances by orders of magnitude; a NaN anywhere classifies the env as hot (the exact path).
  bool hot = false;
  if (runs) {
    const double dn = sqrt((double)__uint_as_float(near_bits));
    const double dx = (double)pos.x - (double)pos0.x, dy = (double)pos.y - (double)pos0.y;
    const double reach = c.minimum_separation + sqrt(dx * dx + dy * dy) + (double)vmax;
    const double slack = 1.0 + 1e-4 * (fabs((double)pos.x) + fabs((double)pos.y) + fabs((double)pos0.x) + fabs((double)pos0.y));
    hot = !(dn > reach + slack);
  }
  const Settled pre = settle_own(c, k, maxstep_hit, pos, goal);
  const bool done = pre.done || (c.time_limit > 0 && cnt.y >= c.time_limit);   // gym TimeLimit of the registered ids
  const bool resets = !hot && done && a.auto_reset;
  const uint32_t bits = (runs ? kOwnRuns : 0u) | ((uint32_t)(cnt.z & 1) * kOwnPlane) | ((hot || resets) ? kOwnSkip : 0u) |
                        (slot << kOwnSlotShift);
  float* rec = reinterpret_cast<float*>(&s.own_b[me]);
  st_release_pair(rec, pos.x, pos.y);
  __threadfence();                                                // (x, y) are visible before the stamp is
  st_release_pair(rec + 2, __uint_as_float(bits), __uint_as_float(stamp));
  // ---- the stream has what it waits for; the rest of the ownship's step
  if (!hot) {
    reinterpret_cast<R*>(a.reward)[me] = (R)pre.reward;
    a.done[me] = done ? 1 : 0;
    a.info[me] = (uint8_t)pre.info;
  } else {
    s.pre[me] = make_double2(pre.reward, __longlong_as_double((long long)(pre.info | ((pre.done ? 1 : 0) << 8))));
  }
  uint8_t vel_f32 = 0;
  if (resets) {
    // the scalar part of reset() (PKG/SingleAircraftEnv.py:66-98); the N spawns are warp jobs of phase B.  The
    // observation handed back for a finished env is reset()'s (dummy_vec_env.py:52-55).
    draw_goal(d, c, goal.x, goal.y);
    reset_ownship<false>(c, d, pos, hs, vel);
    s.goal[me] = goal;
    cnt.x = 0;
    cnt.y = 0;
    cnt.w += 1;
    vel_f32 = 1;
  }
  write_obs_own<FAITH>(a, me, pos.x, pos.y, vel.x, vel.y, vel_f32 != 0, hs.x, hs.y, goal.x, goal.y);   // :115-124
  cnt.z += 1;                                                     // Philox tick; also flips the current position plane
  s.own_pos[me] = pos;
  s.own_hs[me] = hs;
  s.own_vel[me] = vel;
  s.own_vel_f32[me] = vel_f32;
  s.counters[me] = cnt;
  return (hot ? CLS_HOT : resets ? CLS_RESET : CLS_NORMAL) | (runs ? CLS_RUNS : 0);
}

// ------------------------------------------------------------------------------ phase B pieces
// reset_intruder() for intruder i of env `env` (:153-154, :229-238) in the step with tick z: the successor goes to the
// plane this step writes, with its observation entries, its own departure forecast and its distance.
template <bool FAITH>
__device__ __forceinline__ void respawn_one(const StepArgs& a, const size_t env, const int i, const float2 own, const uint32_t z) {
  const DevState& s = a.s;
  Draws<false> d = make_draws<false>(a, env, z);
  Intr<FAITH> it;
  spawn<FAITH, false>(d, a.cfg, a.k, (uint32_t)i, own.x, own.y, it, ihs_slot(s, env, i));
  store_ipos<FAITH>(s, (int)((z & 1u) ^ 1u), env, i, it);
  store_ivel(s, env, i, it.vx, it.vy);
  write_obs_intruder<FAITH>(a, obs_intruder_base<FAITH>(a, env), i, it);
  const size_t fi = flag_index(s, env, i >> 5);
  if constexpr (FAITH) {
    if (it.is64) atomicOr(&s.dflag[fi], 1u << (i & 31));
  }
  const uint32_t nslot = slot_next(slot_of(z));
  const float d2 = dist2_f32(own.x, own.y, (float)it.px, (float)it.py);
  Intr<FAITH> nx = it;
  if (advance_rt<FAITH>(a.k, nx)) atomicOr(&s.fc_gone[(size_t)nslot * flag_plane_words(s) + fi], 1u << (i & 31));
  atomicMin(&s.fc_near[(size_t)nslot * ((size_t)s.T * 32) + env], __float_as_uint(d2));
}

// reset()'s spawns 32 r .. 32 r + 31 of env `env` (:80-88), lanes = intruders; returns the lane's squared distance
template <bool FAITH>
__device__ __forceinline__ float reset_round(const StepArgs& a, const size_t env, const int r, const int lane, const float2 own,
                                             const uint32_t z) {
  const DevState& s = a.s;
  const int i = r * 32 + lane;
  bool wide = false, out = false;
  float d2 = __uint_as_float(kInfBits);
  if (i < s.N) {
    Draws<false> d = make_draws<false>(a, env, z);
    Intr<FAITH> it;
    spawn<FAITH, false>(d, a.cfg, a.k, GCA_SLOT_RESET | (uint32_t)i, own.x, own.y, it, ihs_slot(s, env, i));
    store_ipos<FAITH>(s, (int)((z & 1u) ^ 1u), env, i, it);
    store_ivel(s, env, i, it.vx, it.vy);
    write_obs_intruder<FAITH>(a, obs_intruder_base<FAITH>(a, env), i, it);
    if constexpr (FAITH) wide = it.is64;
    d2 = dist2_f32(own.x, own.y, (float)it.px, (float)it.py);
    Intr<FAITH> nx = it;
    out = advance_rt<FAITH>(a.k, nx);
  }
  const uint32_t dw = __ballot_sync(FULL, wide), fm = __ballot_sync(FULL, out);
  if (lane == 0) {
    const size_t fi = flag_index(s, env, r);
    s.cflag[fi] = 0u;
    if constexpr (FAITH) s.dflag[fi] = dw;
    s.fc_gone[(size_t)slot_next(slot_of(z)) * flag_plane_words(s) + fi] = fm;
  }
  return d2;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// A hot env: the whole of _terminal_reward() (PKG/SingleAircraftEnv.py:143-184) by one warp, lanes = intruders.
// Phase A already advanced the ownship (state stored, tick incremented) and left what the ownship alone would have
// decided in DevState::pre.
template <bool FAITH>
__device__ __forceinline__ void hot_env(const StepArgs& a, const size_t env, const int lane) {
  using R = real_t<FAITH>;
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const Derived& k = a.k;
  float2 pos = s.own_pos[env];
  int4 cnt = s.counters[env];
  const double2 pre = s.pre[env];
  const uint32_t z = (uint32_t)cnt.z - 1u;                 // the tick of this step
  const int cur = (int)(z & 1u), nxt = cur ^ 1;
  const size_t pw = flag_plane_words(s);
  const uint32_t nslot = slot_next(slot_of(z));
  R* obase = obs_intruder_base<FAITH>(a, env);
  // pass 1: the loop returns right after the first intruder inside NMAC_dist (Q9)
  int stop = INT_MAX;
  for (int r = 0; r < s.W && stop == INT_MAX; ++r) {
    const int i = r * 32 + lane;
    bool hit = false;
    if (i < s.N) {
      Intr<FAITH> it;
      load_intruder<FAITH>(s, cur, env, i, it);
      advance_rt<FAITH>(k, it);                             // :150
      bool lt_sep, lt_nmac, lt_init;
      separation<FAITH>(k, pos.x, pos.y, it, lt_sep, lt_nmac, lt_init);   // :151
      hit = lt_sep && lt_nmac;                              // `if dist < NMAC_dist` sits inside `if dist < minimum_separation`
    }
    const uint32_t m = __ballot_sync(FULL, hit);
    if (m) stop = r * 32 + __ffs(m) - 1;
  }
  const bool nmac = stop != INT_MAX;
  // pass 2: everything up to `stop` happened, nothing after it did
  bool conf_any = false;
  int newconf = 0;
  float near2 = __uint_as_float(kInfBits);
  Draws<false> d = make_draws<false>(a, env, z);
  for (int r = 0; r < s.W; ++r) {
    const int i = r * 32 + lane;
    const bool valid = i < s.N;
    Intr<FAITH> fin;
    bool oob = false, lt_sep = false, vis = false;
    if (valid) {
      load_intruder<FAITH>(s, cur, env, i, fin);
      vis = i <= stop;
      if (vis) {
        oob = advance_rt<FAITH>(k, fin);                    // :150, :153
        bool lt_nmac, lt_init;
        separation<FAITH>(k, pos.x, pos.y, fin, lt_sep, lt_nmac, lt_init);   // the OLD object's distance (Q7)
      }
    }
    const size_t fi = flag_index(s, env, r);
    const uint32_t gone_m = __ballot_sync(FULL, vis && oob), conf_m = __ballot_sync(FULL, vis && lt_sep);
    const uint32_t cf = s.cflag[fi];
    newconf += __popc(conf_m & ~cf);                        // False -> True transitions :161-163 (old object's flag, Q7)
    conf_any |= conf_m != 0u;
    const uint32_t ncf = (cf | conf_m) & ~gone_m;           // the flag never clears (Q8); a replaced intruder starts False
    bool out = false, wide = false;
    if (valid) {
      if (vis && oob) {                                     // reset_intruder() :153-154, :229-238
        spawn<FAITH, false>(d, c, k, (uint32_t)i, pos.x, pos.y, fin, ihs_slot(s, env, i));
        store_ivel(s, env, i, fin.vx, fin.vy);
      }
      // (an intruder after `stop` was never touched: it is carried over to the plane this step writes)
      store_ipos<FAITH>(s, nxt, env, i, fin);
      write_obs_intruder<FAITH>(a, obase, i, fin);
      if constexpr (FAITH) wide = fin.is64;
      near2 = fminf(near2, dist2_f32(pos.x, pos.y, (float)fin.px, (float)fin.py));
      Intr<FAITH> nx = fin;
      out = advance_rt<FAITH>(k, nx);
    }
    const uint32_t dw = __ballot_sync(FULL, wide), fm = __ballot_sync(FULL, out);
    if (lane == 0) {
      if (ncf != cf) s.cflag[fi] = ncf;
      if constexpr (FAITH) s.dflag[fi] = dw;
      s.fc_gone[(size_t)nslot * pw + fi] = fm;
    }
  }
  near2 = warp_min(near2);
  // _terminal_reward()'s return   :165-183
  double reward;
  int info;
  bool done = false;
  if (nmac) {
    reward = c.r_nmac; done = true; info = GCA_INFO_NMAC;
  } else if (conf_any) {
    reward = c.r_conflict; info = GCA_INFO_CONFLICT;
  } else {
    const int bits = (int)__double_as_longlong(pre.y);
    reward = pre.x; info = bits & 0xff; done = (bits >> 8) != 0;
  }
  if (c.time_limit > 0 && cnt.y >= c.time_limit) done = true;      // gym TimeLimit of the registered ids
  cnt.x += newconf;
  if (done && a.auto_reset) {
    // VecEnv auto-reset (dummy_vec_env.py:52-55): reset() PKG/SingleAircraftEnv.py:66-98
    double2 hs, vel, goal;
    if (lane == 0) {
      draw_goal(d, c, goal.x, goal.y);
      reset_ownship<false>(c, d, pos, hs, vel);
      s.own_pos[env] = pos;
      s.own_hs[env] = hs;
      s.own_vel[env] = vel;
      s.own_vel_f32[env] = 1;
      s.goal[env] = goal;
      write_obs_own<FAITH>(a, env, pos.x, pos.y, vel.x, vel.y, true, hs.x, hs.y, goal.x, goal.y);
    }
    pos.x = __shfl_sync(FULL, pos.x, 0);
    pos.y = __shfl_sync(FULL, pos.y, 0);
    __syncwarp();                                           // the loop's stores come before the reset's, lane by lane anyway
    near2 = __uint_as_float(kInfBits);
    for (int r = 0; r < s.W; ++r) near2 = fminf(near2, reset_round<FAITH>(a, env, r, lane, pos, z));
    near2 = warp_min(near2);
    cnt.x = 0;
    cnt.y = 0;
    cnt.w += 1;
  }
  if (lane == 0) {
    reinterpret_cast<R*>(a.reward)[env] = (R)reward;
    a.done[env] = done ? 1 : 0;
    a.info[env] = (uint8_t)info;
    s.counters[env] = cnt;
    s.fc_near[(size_t)nslot * ((size_t)s.T * 32) + env] = __float_as_uint(near2);
  }
}

// ------------------------------------------------------------------------------ the head kernel
template <bool FAITH>
__global__ void __launch_bounds__(kHeadThreads, 2) step_head_kernel(const __grid_constant__ StepArgs a) {
  __shared__ uint8_t cls[kHeadMaxGroups][kHeadThreads];
  __shared__ uint32_t jobs[kJobCap];
  __shared__ int hot_list[kHeadThreads], rst_list[kHeadThreads];
  __shared__ int n_jobs, n_hot, n_rst;
  const DevState& s = a.s;
  pdl_wait();                                               // the previous step (its stream and its head) is complete
  const uint32_t stamp = *s.step_seq + 1u;
  pdl_launch_dependents();                                  // every block of this grid is resident: the stream may start
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const size_t padded = (size_t)s.T * 32;
  const int n_groups = (int)((padded + kHeadThreads - 1) / kHeadThreads);
  // ---- phase A: the records of all groups of this block, in the order the stream will ask for them
  int gi = 0;
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
    const size_t me = (size_t)g * kHeadThreads + tid;
    int kind = CLS_NONE;
    if (me < (size_t)s.B) {
      kind = head_own<FAITH>(a, me, stamp);
    } else if (me < padded) {                               // padding lanes of the last tile: nothing to do for them
      float* rec = reinterpret_cast<float*>(&s.own_b[me]);
      st_release_pair(rec, 0.f, 0.f);
      st_release_pair(rec + 2, __uint_as_float(kOwnSkip), __uint_as_float(stamp));
    }
    cls[gi][tid] = (uint8_t)kind;
  }
  // ---- phase B, group by group: respawns (thread = spawn), hot envs (warp = env), resets (warp = 32 spawns)
  const size_t pw = flag_plane_words(s);
  gi = 0;
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
    if (tid == 0) n_jobs = n_hot = n_rst = 0;
    __syncthreads();
    const size_t me = (size_t)g * kHeadThreads + tid;
    const int kind = cls[gi][tid];
    if (kind != CLS_NONE) {
      const uint32_t z = (uint32_t)s.counters[me].z - 1u;   // the tick of this step (phase A incremented it)
      const uint32_t slot = slot_of(z), cslot = slot_next(slot_next(slot));
      const bool respawns = (kind & 3) == CLS_NORMAL && (kind & CLS_RUNS);
      float2 own = make_float2(0.f, 0.f);
      if (respawns) own = s.own_pos[me];
      for (int w = 0; w < s.W; ++w) {
        const size_t fi = flag_index(s, me, w);
        s.fc_gone[(size_t)cslot * pw + fi] = 0u;            // the slot the NEXT step fills
        if (!respawns) continue;
        uint32_t gone = s.fc_gone[(size_t)slot * pw + fi];
        if (!gone) continue;
        const uint32_t cf = s.cflag[fi];
        if (cf & gone) s.cflag[fi] = cf & ~gone;            // a replaced intruder starts with conflict False
        if constexpr (FAITH) {
          const uint32_t df = s.dflag[fi];
          if (df & gone) s.dflag[fi] = df & ~gone;          // (a retried spawn sets its bit again, atomically)
        }
        const int at0 = atomicAdd(&n_jobs, __popc(gone));
        int at = at0;
        while (gone) {
          const int i = w * 32 + __ffs(gone) - 1;
          gone &= gone - 1;
          if (at < kJobCap) jobs[at] = ((uint32_t)tid << 16) | (uint32_t)i;
          else respawn_one<FAITH>(a, me, i, own, z);        // (list full: > 6 respawns per env on average)
          ++at;
        }
      }
      s.fc_near[(size_t)cslot * padded + me] = kInfBits;
      if ((kind & 3) == CLS_HOT) hot_list[atomicAdd(&n_hot, 1)] = tid;
      if ((kind & 3) == CLS_RESET) rst_list[atomicAdd(&n_rst, 1)] = tid;
    }
    __syncthreads();
    const int nj = n_jobs < kJobCap ? n_jobs : kJobCap;
    for (int j = tid; j < nj; j += kHeadThreads) {
      const uint32_t rec = jobs[j];
      const size_t env = (size_t)g * kHeadThreads + (rec >> 16);
      respawn_one<FAITH>(a, env, (int)(rec & 0xffffu), s.own_pos[env], (uint32_t)s.counters[env].z - 1u);
    }
    for (int h = wib; h < n_hot; h += kHeadThreads / 32) hot_env<FAITH>(a, (size_t)g * kHeadThreads + hot_list[h], lane);
    const int rounds = s.W;
    for (int job = wib; job < n_rst * rounds; job += kHeadThreads / 32) {
      const size_t env = (size_t)g * kHeadThreads + rst_list[job / rounds];
      const uint32_t z = (uint32_t)s.counters[env].z - 1u;
      const float d2 = warp_min(reset_round<FAITH>(a, env, job % rounds, lane, s.own_pos[env], z));
      if (lane == 0) atomicMin(&s.fc_near[(size_t)slot_next(slot_of(z)) * padded + env], __float_as_uint(d2));
    }
    __syncthreads();                                        // (the lists are reused by the next group)
  }
}

// ------------------------------------------------------------------------------ (re)building the forecast
// After reset / set_state / a change of configuration: the forecast words, distance summary and displacement bound
// of the CURRENT state of every env, in the slot its next step reads; thread = env.
template <bool FAITH>
__global__ void __launch_bounds__(128) forecast_kernel(const StepArgs a) {
  const DevState& s = a.s;
  const size_t me = (size_t)blockIdx.x * 128 + threadIdx.x;
  const size_t padded = (size_t)s.T * 32, pw = flag_plane_words(s);
  if (me >= padded) return;
  if (me >= (size_t)s.B) {
    for (int q = 0; q < 3; ++q) s.fc_near[(size_t)q * padded + me] = kInfBits;
    return;
  }
  const uint32_t z = (uint32_t)s.counters[me].z;
  const uint32_t slot = slot_of(z), nslot = slot_next(slot);
  const float2 own = s.own_pos[me];
  // intruders fly at most max_speed (uniform(min_speed, max_speed), constant afterwards) plus the drift per axis;
  // velocities set through gca_set_state may be anything, so the bound is raised to what the state holds
  const float drift = fabsf(a.k.drift_f);
  float vmax = (float)a.cfg.max_speed * 1.0001f + 1.5f * drift;
  float near2 = __uint_as_float(kInfBits);
  for (int w = 0; w < s.W; ++w) {
    uint32_t word = 0u;
    for (int j = 0; j < 32 && w * 32 + j < s.N; ++j) {
      const int i = w * 32 + j;
      Intr<FAITH> it;
      load_intruder<FAITH>(s, (int)(z & 1u), me, i, it);
      near2 = fminf(near2, dist2_f32(own.x, own.y, (float)it.px, (float)it.py));
      const float sp = sqrtf((fabsf(it.vx) + drift) * (fabsf(it.vx) + drift) + (fabsf(it.vy) + drift) * (fabsf(it.vy) + drift));
      if (!(sp * 1.0001f <= vmax)) vmax = sp * 1.0001f;     // (a NaN velocity makes the bound NaN: the env stays hot)
      if (advance_rt<FAITH>(a.k, it)) word |= 1u << j;
    }
    const size_t fi = flag_index(s, me, w);
    s.fc_gone[(size_t)slot * pw + fi] = word;
    s.fc_gone[(size_t)nslot * pw + fi] = 0u;
  }
  s.fc_near[(size_t)slot * padded + me] = __float_as_uint(near2);
  s.fc_near[(size_t)nslot * padded + me] = kInfBits;
  s.fc_vmax[me] = vmax;
}

// ------------------------------------------------------------------------------ launchers
bool forecast_step_applies(const StepArgs& a, bool tape) {
  static const int off = std::getenv("GCA_NO_FORECAST") ? 1 : 0;
  return !off && !tape && a.s.N > 0 && a.s.fc_gone != nullptr && !a.cfg.shaped_nearest && !(a.cfg.intruder_turns && a.s.ihs);
}

cudaError_t launch_forecast(bool faith, const StepArgs& a, cudaStream_t st) {
  const unsigned blocks = (unsigned)(((size_t)a.s.T * 32 + 127) / 128);
  if (faith) forecast_kernel<true><<<blocks, 128, 0, st>>>(a);
  else forecast_kernel<false><<<blocks, 128, 0, st>>>(a);
  return cudaGetLastError();
}

static int head_blocks(const DevState& s) {
  static int sms = 0;
  static int per_sm = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    const char* v = std::getenv("GCA_HEAD_CTAS_PER_SM");
    per_sm = v ? std::atoi(v) : 1;
    if (per_sm < 1) per_sm = 1;
  }
  const int n_groups = (int)(((size_t)s.T * 32 + kHeadThreads - 1) / kHeadThreads);
  int blocks = sms * per_sm;
  if (blocks > n_groups) blocks = n_groups;
  const int need = (n_groups + kHeadMaxGroups - 1) / kHeadMaxGroups;
  return blocks > need ? blocks : need;
}

// ev (nullable): 5 events; the two kernels overlap, so they are timed as one interval (1 -> 2)
cudaError_t launch_step_fc(bool faith, const StepArgs& a0, cudaStream_t st, cudaEvent_t* ev) {
  StepArgs a = a0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (ev) cudaStreamIsCapturing(st, &cap);
  auto mark = [&](int i) {
    if (!ev) return;
    if (cap == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(ev[i], st, cudaEventRecordExternal);
    else cudaEventRecord(ev[i], st);
  };
  mark(0);
  mark(1);
  a.own_blocks = 0;
  a.head_ctas = head_blocks(a.s);
  if (faith) launch_pdl(step_head_kernel<true>, (unsigned)a.head_ctas, kHeadThreads, st, a);
  else launch_pdl(step_head_kernel<false>, (unsigned)a.head_ctas, kHeadThreads, st, a);
  cudaError_t rc = launch_stream_fc(faith, a, st);
  mark(2);
  mark(3);
  if (rc == cudaSuccess) rc = launch_step_tail(faith, a, st);
  mark(4);
  return rc != cudaSuccess ? rc : cudaGetLastError();
}

}  // namespace gca
