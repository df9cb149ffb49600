// gca_step_fc.cuh - the forecast step: one step = the jobs kernel + the main kernel (ownship role + streaming pass),
// running CONCURRENTLY, and nothing after them (sm_100a).  PHILOX handles with intruders whose reward does not need
// the step's nearest distance and whose intruders do not turn (every registered id and BASELINE config; the others
// keep gca_step.cu's own-role + finish pair).
//
// Why.  A step is: ownship update -> 80 intruders advance (the 220 MB stream) -> the reference's sequential loop
// semantics (respawn what left the map, conflict flags, first NMAC wins, reward, done, auto-reset).  The third part
// needs every intruder of the env, so as a kernel behind the stream it is a 10-15 us latency chain at the end of
// every step, whatever its width (gca_step.cu's step_finish_kernel).  Here nothing of it waits for the stream:
//   * departures are FORECAST one step ahead: the pass that stores a position also makes the f32 sum and the map
//     test the next step will make on it (same operands, same rounding - the forecast is the advance), so it is known
//     at the START of a step which intruders leave in it.  The jobs kernel spawns their successors (one thread per
//     spawn) while the stream runs; the stream stores nothing for those intruders.
//   * a conflict needs an intruder inside minimum_separation, an NMAC one inside NMAC_dist.  The stream also keeps the
//     smallest squared distance of the state it writes; with the ownship's own displacement (known once the action is
//     applied) and the bound on an intruder's displacement per step, the triangle inequality tells the ownship role
//     which envs CANNOT see an NMAC in this step.  In those the order of the reference's loop does not matter: a
//     streaming lane that finds an intruder inside minimum_separation writes the conflict return and flag itself.
//     The others ("hot": an NMAC is possible, or a conflict in an env that the ownship alone would finish) are
//     advanced by the jobs kernel, a warp per env with lanes = intruders, which replays PKG/SingleAircraftEnv.py
//     :149-170 exactly (first NMAC index wins, later intruders untouched (Q9), a replaced intruder is tested with its
//     old distance (Q7), flags never clear (Q8)); the stream skips them.
//   * an env that finishes without a conflict being possible (wall / goal / max steps / TimeLimit) is reset by the
//     jobs kernel (VecEnv auto-reset, dummy_vec_env.py:52-55) and skipped by the stream as well.
// Launches: the main kernel - its leading blocks are the HEAD (persistent, one or two per SM: first the ownship role
// of every group of 128 envs the block owns, in the order the stream will ask for them, then their respawns), the
// others the stream; blocks are dispatched in index order, so the head blocks are resident before any streaming block
// waits for a record, and they never wait for anything themselves - and behind it, with a programmatic launch edge
// that the main kernel triggers at once, the tail kernel (hot envs, resets): its blocks are scheduled as the stream
// drains.  (Two separate kernels for head and stream do not overlap reliably: an SM changes its L1 / shared-memory
// split only when idle, and at most two kernels of a stream overlap under programmatic launch.)  Records reach the
// stream through DevState::own_b, stamped with the step count.
// (included by gca_step.cu)
#pragma once

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

// ------------------------------------------------------------------------------ the ownship role: thread = env
// Ownship.step(a)   PKG/SingleAircraftEnv.py:299-309 (2Env :291-301, DiscreteHER :301-311), the classification, the
// record the stream waits for, and everything of the step that the ownship alone decides.  Called by all 128 threads
// of a leading block of the main kernel.  What a streaming lane may touch later (reward / info on a conflict, the
// conflict counter) is stored BEFORE the record is published; an env in which no conflict is possible publishes as
// soon as its new position is known.
template <bool FAITH>
__device__ __forceinline__ void own_role_fc(const StepArgs& a, const uint32_t stamp, const int group) {
  using R = real_t<FAITH>;
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const Derived& k = a.k;
  GCA_KSTAMP_IN(0);
  const size_t me = (size_t)group * kHeadThreads + threadIdx.x;
  float* rec = reinterpret_cast<float*>(&s.own_b[me < (size_t)s.T * 32 ? me : 0]);
  if (me >= (size_t)s.B) {
    if (me < (size_t)s.T * 32)                                // padding lanes of the last tile: nothing to do for them
      st_release_quad(rec, 0.f, 0.f, __uint_as_float(kOwnSkip), __uint_as_float(stamp));
  } else {
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
    const float2 pos = make_float2((float)__dadd_rn((double)pos0.x, vel.x), (float)__dadd_rn((double)pos0.y, vel.y));
    cnt.y += 1;                                                     // StackEnv :118
    const bool maxstep_hit = c.max_steps > 0 && cnt.y >= c.max_steps;   // StackEnv :134-136: the intruder loop never runs
    const bool runs = !maxstep_hit;
    // does the ownship alone end the episode?  (the distance test on the squared distance, as for the intruders)
    bool own_done = maxstep_hit;
    if (!own_done && c.wall_kind == GCA_WALL_TERMINAL && !in_map_f32(k, pos.x, pos.y)) own_done = true;
    if (!own_done && !(c.wall_kind != GCA_WALL_NONE && !in_map_f32(k, pos.x, pos.y)))
      own_done = dist2_f64((double)pos.x, (double)pos.y, goal.x, goal.y) < k.goal2_d;
    const bool done = own_done || (c.time_limit > 0 && cnt.y >= c.time_limit);   // gym TimeLimit of the registered ids
    // How close can an intruder be after this step?  Every intruder was at least sqrt(near) away from the old ownship
    // position; the ownship moved by no more than |speed| (+ the f32 rounding of the position), an intruder moves by
    // at most vmax.  The slack covers the f32 rounding of positions and distances by orders of magnitude; a NaN
    // anywhere makes the env hot (the exact path).
    bool hot = false, conf_possible = false;
    if (runs) {
      const double near2 = (double)__uint_as_float(near_bits);
      const double slack = 1.0 + 1e-4 * (fabs((double)pos.x) + fabs((double)pos.y) + fabs((double)pos0.x) + fabs((double)pos0.y));
      const double reach = fabs(speed) + (double)vmax + slack;
      const double r_nmac = (c.nmac_dist < c.minimum_separation ? c.nmac_dist : c.minimum_separation) + reach;
      const double r_conf = c.minimum_separation + reach;
      conf_possible = !(near2 > r_conf * r_conf);
      hot = !(near2 > r_nmac * r_nmac) || (done && conf_possible);
    }
    const bool resets = !hot && done && a.auto_reset;
    const uint32_t bits = (runs ? kOwnRuns : 0u) | ((uint32_t)(cnt.z & 1) * kOwnPlane) | ((hot || resets) ? kOwnSkip : 0u) |
                          (slot << kOwnSlotShift) | (hot ? kOwnHot : 0u) | (resets ? kOwnReset : 0u) |
                          (conf_possible ? kOwnConf : 0u);
    if (hot) s.fc_queue[4 + atomicAdd(&s.fc_queue[0], 1)] = (int)me;              // the tail kernel's warp jobs
    if (resets) s.fc_queue[4 + (size_t)s.T * 32 + atomicAdd(&s.fc_queue[1], 1)] = (int)me;
    const bool early = !conf_possible;                            // no lane will touch this env's outputs
    if (early) st_release_quad(rec, pos.x, pos.y, __uint_as_float(bits), __uint_as_float(stamp));   // (one 16-byte store)
    GCA_KSTAMP_IN(4);
    const Settled pre = settle_own(c, k, maxstep_hit, pos, goal);
    if (!hot) {
      reinterpret_cast<R*>(a.reward)[me] = (R)pre.reward;
      a.done[me] = done ? 1 : 0;
      a.info[me] = (uint8_t)pre.info;
    } else {
      s.pre[me] = make_double2(pre.reward, __longlong_as_double((long long)(pre.info | ((pre.done ? 1 : 0) << 8))));
    }
    uint8_t vel_f32 = 0;
    float2 pos_new = pos;
    if (resets) {
      // the scalar part of reset() (PKG/SingleAircraftEnv.py:66-98); the N spawns are warp jobs of the jobs kernel.
      // The observation handed back for a finished env is reset()'s (dummy_vec_env.py:52-55).
      draw_goal(d, c, goal.x, goal.y);
      reset_ownship<false>(c, d, pos_new, hs, vel);
      s.goal[me] = goal;
      cnt.x = 0;
      cnt.y = 0;
      cnt.w += 1;
      vel_f32 = 1;
    }
    cnt.z += 1;                                                     // Philox tick; also flips the current position plane
    s.counters[me] = cnt;
    if (!early) {
      __threadfence();                                              // reward / info / counters are visible before the record is
      st_release_quad(rec, pos.x, pos.y, __uint_as_float(bits), __uint_as_float(stamp));
    }
    GCA_KSTAMP_OUT(4);
    // ---- the stream has what it waits for; the rest of the ownship's step
    write_obs_own<FAITH>(a, me, pos_new.x, pos_new.y, vel.x, vel.y, vel_f32 != 0, hs.x, hs.y, goal.x, goal.y);   // :115-124
    s.own_pos[me] = pos_new;
    s.own_hs[me] = hs;
    s.own_vel[me] = vel;
    s.own_vel_f32[me] = vel_f32;
  }
  GCA_KSTAMP_OUT(0);
}

// ------------------------------------------------------------------------------ the jobs kernel: pieces
// One spawn: Aircraft(random_pos(), random_speed(), random_heading()) + rejection loop, stored in the plane this step
// writes together with its velocity and observation entries.  Returns the successor's squared distance to the
// ownship; `out`: it leaves the map at its first advance; `wide`: its position is a true f64 (FAITHFUL, Q3).
template <bool FAITH>
__device__ __noinline__ float spawn_store(const StepArgs& a, const size_t env, const uint32_t slot_tag, const int i,
                                          const float ox, const float oy, const uint32_t z, bool& out, bool& wide) {
  const DevState& s = a.s;
  Draws<false> d = make_draws<false>(a, env, z);
  Intr<FAITH> it;
  spawn<FAITH, false>(d, a.cfg, a.k, slot_tag | (uint32_t)i, ox, oy, it, ihs_slot(s, env, i));
  store_ipos<FAITH>(s, (int)((z & 1u) ^ 1u), env, i, it);
  store_ivel(s, env, i, it.vx, it.vy);
  write_obs_intruder<FAITH>(a, obs_intruder_base<FAITH>(a, env), i, it);
  wide = false;
  if constexpr (FAITH) wide = it.is64;
  const float d2 = dist2_f32(ox, oy, (float)it.px, (float)it.py);
  out = advance_rt<FAITH>(a.k, it);
  return d2;
}

// reset_intruder() for intruder i of a normal env (:153-154, :229-238): first the OLD object's conflict test - the
// reference measures the distance before it replaces the intruder, and a conflict of the replaced object still makes
// the step's return (r_conflict, False, 'c') and counts if its flag was False (Q7) -, then the successor.
// rec: bit 15 the old object's conflict flag, bit 14 the parity of the env's current plane, bit 13 a conflict is possible
template <bool FAITH>
__device__ __forceinline__ void respawn_job(const StepArgs& a, const size_t env, const uint32_t rec) {
  using R = real_t<FAITH>;
  const DevState& s = a.s;
  const int i = (int)(rec & 0x1fffu);
  const int cur = (int)((rec >> 14) & 1u);
  // everything this job reads is requested at once (one round trip)
  const float4 ob = __ldcg(&s.own_b[env]);                  // this step's ownship position
  const uint32_t z = (uint32_t)__ldcg(&s.counters[env].z) - 1u;   // the tick of this step
  if (rec & 0x2000u) {
    Intr<FAITH> old;
    load_intruder<FAITH>(s, cur, env, i, old);
    advance_rt<FAITH>(a.k, old);
    bool lt_sep, lt_nmac, lt_init;
    separation<FAITH>(a.k, ob.x, ob.y, old, lt_sep, lt_nmac, lt_init);
    if (lt_sep) {
      if (lt_nmac) atomicOr(s.error_flag, 4);               // (an env where an NMAC is possible is hot: never)
      reinterpret_cast<R*>(a.reward)[env] = (R)a.cfg.r_conflict;
      a.info[env] = (uint8_t)GCA_INFO_CONFLICT;
      if (!(rec & 0x8000u)) atomicAdd(&s.counters[env].x, 1);
    }
  }
  bool out, wide;
  const float d2 = spawn_store<FAITH>(a, env, 0u, i, ob.x, ob.y, z, out, wide);
  const size_t fi = flag_index(s, env, i >> 5);
  if (wide) atomicOr(&s.dflag[fi], 1u << (i & 31));
  const uint32_t nslot = slot_next(slot_of(z));
  if (out) atomicOr(&s.fc_gone[(size_t)nslot * flag_plane_words(s) + fi], 1u << (i & 31));
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
  if (i < s.N) d2 = spawn_store<FAITH>(a, env, GCA_SLOT_RESET, i, own.x, own.y, z, out, wide);
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
// The ownship kernel already advanced the ownship (state stored, tick incremented) and left what the ownship alone
// would have decided in DevState::pre; the position the intruders are tested against is the record's.
template <bool FAITH>
__device__ __forceinline__ void hot_env(const StepArgs& a, const size_t env, const int lane) {
  using R = real_t<FAITH>;
  const DevState& s = a.s;
  const gca_config& c = a.cfg;
  const Derived& k = a.k;
  const float4 ob = __ldcg(&s.own_b[env]);
  float2 pos = make_float2(ob.x, ob.y);
  int4 cnt = __ldcg(&s.counters[env]);
  const double2 pre = __ldcg(&s.pre[env]);
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
        near2 = fminf(near2, spawn_store<FAITH>(a, env, 0u, i, pos.x, pos.y, z, out, wide));
      } else {
        // (an intruder after `stop` was never touched: it is carried over to the plane this step writes)
        store_ipos<FAITH>(s, nxt, env, i, fin);
        write_obs_intruder<FAITH>(a, obase, i, fin);
        if constexpr (FAITH) wide = fin.is64;
        near2 = fminf(near2, dist2_f32(pos.x, pos.y, (float)fin.px, (float)fin.py));
        out = advance_rt<FAITH>(k, fin);
      }
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
    if (lane == 0) {
      double2 hs, vel, goal;
      Draws<false> d = make_draws<false>(a, env, z);
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
    near2 = __uint_as_float(kInfBits);                      // (same lane, same intruder: the reset's stores come after the loop's)
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

// ------------------------------------------------------------------------------ the head role
// The leading a.head_ctas blocks of the main kernel; block b owns the groups b, b + head_ctas, ... of 128 envs.  After
// the ownship role it runs next to the stream for the rest of the step and only has to be done when the stream is:
// latency does not matter there, footprint does (few, long-lived blocks).  jobs: kJobCap words of shared memory.
template <bool FAITH>
__device__ __forceinline__ void head_role_fc(const StepArgs& a, const uint32_t stamp, uint32_t* jobs) {
  __shared__ int n_jobs;
  const DevState& s = a.s;
  const int tid = threadIdx.x;
  const int head_ctas = a.head_ctas;
  if (tid == 0) n_jobs = 0;
  const size_t padded = (size_t)s.T * 32, pw = flag_plane_words(s);
  const int n_groups = (int)((padded + kHeadThreads - 1) / kHeadThreads);
  // ---- the ownship role of every group of this block: the records the stream waits for
  for (int g = blockIdx.x; g < n_groups; g += head_ctas) own_role_fc<FAITH>(a, stamp, g);
  __threadfence();
  __syncthreads();                                          // (the jobs read what other threads of the block stored)
  GCA_KSTAMP_IN(2);
  // ---- thread = env: this step's departures become respawn records, the next step's forecast slot is cleared
  int gi = 0;
  for (int g = blockIdx.x; g < n_groups; g += head_ctas, ++gi) {
    const size_t me = (size_t)g * kHeadThreads + tid;
    if (me >= (size_t)s.B) continue;
    const uint32_t bits = __float_as_uint(__ldcg(&s.own_b[me]).z);
    const uint32_t slot = (bits >> kOwnSlotShift) & 3u, cslot = slot_next(slot_next(slot));
    const bool respawns = (bits & (kOwnRuns | kOwnSkip)) == kOwnRuns;
    const uint32_t local = (uint32_t)(gi * kHeadThreads + tid);
    const uint32_t tag = (local << 16) | ((bits & kOwnPlane) ? 0x4000u : 0u) | ((bits & kOwnConf) ? 0x2000u : 0u);
    for (int w = 0; w < s.W; ++w) {
      const size_t fi = flag_index(s, me, w);
      s.fc_gone[(size_t)cslot * pw + fi] = 0u;              // the slot the NEXT step fills
      if (!respawns) continue;
      uint32_t gone = s.fc_gone[(size_t)slot * pw + fi];
      if (!gone) continue;
      // a replaced intruder starts with conflict False / an f32 position (a retried spawn sets the bit again); the
      // stream sets OTHER bits of the conflict word atomically at the same time
      const uint32_t cf = atomicAnd(&s.cflag[fi], ~gone);
      if constexpr (FAITH) atomicAnd(&s.dflag[fi], ~gone);
      int at = atomicAdd(&n_jobs, __popc(gone));
      while (gone) {
        const int j = __ffs(gone) - 1;
        gone &= gone - 1;
        const uint32_t rec = tag | (((cf >> j) & 1u) ? 0x8000u : 0u) | (uint32_t)(w * 32 + j);
        if (at < kJobCap) jobs[at] = rec;
        else respawn_job<FAITH>(a, me, rec);                // (list full)
        ++at;
      }
    }
    s.fc_near[(size_t)cslot * padded + me] = kInfBits;
  }
  __syncthreads();
  auto env_of = [&](uint32_t local) { return ((size_t)blockIdx.x + (size_t)(local / kHeadThreads) * head_ctas) * kHeadThreads + local % kHeadThreads; };
  // ---- respawns: thread = spawn
  const int nj = n_jobs < kJobCap ? n_jobs : kJobCap;
  for (int j = tid; j < nj; j += kHeadThreads) respawn_job<FAITH>(a, env_of(jobs[j] >> 16), jobs[j]);
  // the head's completion mark (the tail kernel's jobs wait for it)
  __threadfence();
  __syncthreads();
  if (tid == 0 && atomicAdd(&s.head_sync[0], 1u) == (unsigned)head_ctas - 1u) {
    s.head_sync[0] = 0u;
    __threadfence();
    *reinterpret_cast<volatile uint32_t*>(&s.head_sync[1]) = stamp;
  }
  GCA_KSTAMP_OUT(2);
#ifdef GCA_PHASE_TIMING
  if (threadIdx.x == 0) atomicAdd(&g_kstamp[12], (unsigned long long)n_jobs);
#endif
}

// ------------------------------------------------------------------------------ the tail kernel
// The hot envs and the resets of the step, one warp per job (hot env: the whole env; reset: 32 of its N spawns), pulled
// from the lists the ownship role filled.  The stream skips these envs, so the jobs depend on the head only; the
// kernel is launched behind the stream with a programmatic edge that the stream triggers at once, i.e. its blocks
// are scheduled as soon as the stream's last blocks have been dispatched and the jobs run at full width while the
// stream drains.  Being the last kernel of the step, its last block closes the step: it waits for the stream to be
// complete, clears the queue and advances the step count the next records are stamped with.
constexpr int kTailThreads = 64;
template <bool FAITH>
__global__ void __launch_bounds__(kTailThreads) step_tail_kernel(const __grid_constant__ StepArgs a) {
  const DevState& s = a.s;
  GCA_KSTAMP_IN(5);
  const uint32_t stamp = *s.step_seq + 1u;                  // (stable: only this kernel's last block changes it)
  const int lane = threadIdx.x & 31;
  const size_t padded = (size_t)s.T * 32;
  if (lane == 0) {                                          // the head is complete (it is running or done: it never waits)
    int spins = 0;
    while (*reinterpret_cast<volatile uint32_t*>(&s.head_sync[1]) != stamp) {
      if (++spins > (1 << 22)) {
        atomicOr(s.error_flag, 1);
        break;
      }
      __nanosleep(64);
    }
    __threadfence();
  }
  __syncwarp();
  const int n_hot = __ldcg(&s.fc_queue[0]), n_rst = __ldcg(&s.fc_queue[1]);
  const int rounds = s.W, total = n_hot + n_rst * rounds;
  for (;;) {
    int job = 0;
    if (lane == 0) job = atomicAdd(&s.fc_queue[2], 1);
    job = __shfl_sync(FULL, job, 0);
    if (job >= total) break;
    if (job < n_hot) {
      hot_env<FAITH>(a, (size_t)__ldcg(&s.fc_queue[4 + job]), lane);
    } else {
      const int rj = job - n_hot;
      const size_t env = (size_t)__ldcg(&s.fc_queue[4 + padded + rj / rounds]);
      const uint32_t z = (uint32_t)__ldcg(&s.counters[env].z) - 1u;
      const float2 own = __ldcg(&s.own_pos[env]);
      const float d2 = warp_min(reset_round<FAITH>(a, env, rj % rounds, lane, own, z));
      if (lane == 0) atomicMin(&s.fc_near[(size_t)slot_next(slot_of(z)) * padded + env], __float_as_uint(d2));
    }
  }
#ifdef GCA_PHASE_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) { g_kstamp[13] = (unsigned long long)n_hot; g_kstamp[14] = (unsigned long long)n_rst; }
#endif
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&s.fc_queue[3], 1) == (int)gridDim.x - 1) {
    pdl_wait();                                             // the stream is complete (every block of it has read the step count)
    s.fc_queue[0] = s.fc_queue[1] = s.fc_queue[2] = s.fc_queue[3] = 0;
    *s.step_seq = stamp;
  }
  GCA_KSTAMP_OUT(5);
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
// Opt-in (GCA_FORECAST=1): measured on B200 at 65,536 x 80 the forecast step is correct (bit-exact against the oracle,
// the whole GPU suite) but at 62-70 us per step not yet faster than the own-role + finish pair (51 us): under the
// stream's load every dependent access of the head's latency chains costs 2-3 us and the head blocks take registers
// from the stream (DESIGN.md section 7 has the timelines).
bool forecast_step_applies(const StepArgs& a, bool tape) {
  static const int on = (std::getenv("GCA_FORECAST") && std::atoi(std::getenv("GCA_FORECAST")) != 0) ? 1 : 0;
  return on && !tape && a.s.N > 0 && a.s.fc_gone != nullptr && !a.cfg.shaped_nearest && !(a.cfg.intruder_turns && a.s.ihs);
}

cudaError_t launch_forecast(bool faith, const StepArgs& a, cudaStream_t st) {
  const unsigned blocks = (unsigned)(((size_t)a.s.T * 32 + 127) / 128);
  if (faith) forecast_kernel<true><<<blocks, 128, 0, st>>>(a);
  else forecast_kernel<false><<<blocks, 128, 0, st>>>(a);
  return cudaGetLastError();
}

static int head_blocks_per_sm_unit() {                      // the SM count of the current device
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

static int head_blocks(const DevState& s) {
  static int per_sm = 0;
  const int sms = head_blocks_per_sm_unit();
  if (!per_sm) {
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
  static bool once = false;
  if (!once) {
    prefer_carveout(step_tail_kernel<true>);
    prefer_carveout(step_tail_kernel<false>);
    once = true;
  }
  cudaError_t rc = launch_stream_fc(faith, a, st);
  const unsigned tail_blocks = (unsigned)(head_blocks_per_sm_unit() * 8);
  if (faith) launch_pdl(step_tail_kernel<true>, tail_blocks, kTailThreads, st, a);
  else launch_pdl(step_tail_kernel<false>, tail_blocks, kTailThreads, st, a);
  mark(2);
  mark(3);
  if (rc == cudaSuccess) rc = launch_step_tail(faith, a, st);
  mark(4);
  return rc != cudaSuccess ? rc : cudaGetLastError();
}

}  // namespace gca
