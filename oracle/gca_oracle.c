/* gca_oracle.c - CPU restatement of the reference's reset/step hot path (see gca_oracle.h).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Build with -ffp-contract=off (oracle/Makefile): every
 * arithmetic operation below must round exactly once, like the Python/NumPy scalar code it
 * restates.  Citations are to the reference tree (PKG = gym_guidance_collision_avoidance_single/envs).
 *
 * One environment is advanced at a time, in the reference's own sequential order; nothing here
 * is vectorised or shared with the CUDA kernels except gca_math.h (the bit-reproducible
 * sincos/log used when trig == GCA_TRIG_SHARED).
 */
#include "gca_oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>

#include "gca_math.h"


/* ------------------------------------------------------------------------- Philox4x32-10 */
/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); constants of Random123. */
void gca_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53-bit uniform in [0,1) from two words, numpy's random_sample recipe */
static double u53(uint32_t a, uint32_t b) {
  return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

void gca_oracle_philox_uniform2(uint64_t seed, uint32_t env, uint32_t tick, uint32_t slot, uint32_t block, double u[2]) {
  uint32_t ctr[4] = {env, tick, slot, block};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t w[4];
  gca_oracle_philox4x32_10(ctr, key, w);
  u[0] = u53(w[0], w[1]);
  u[1] = u53(w[2], w[3]);
}

void gca_oracle_sincos(double x, int trig, double* s, double* c) {
  if (trig == GCA_TRIG_SHARED) {
    gca_sincos(x, s, c);
  } else {
    *s = sin(x);
    *c = cos(x);
  }
}

double gca_oracle_log(double x, int trig) { return trig == GCA_TRIG_SHARED ? gca_log(x) : log(x); }

/* Box-Muller on block 0 of `slot`: g0 = r cos(2 pi u1), g1 = r sin(2 pi u1), r = sqrt(-2 ln(1-u0)) */
void gca_oracle_philox_normal2(uint64_t seed, uint32_t env, uint32_t tick, uint32_t slot, int trig, double g[2]) {
  double u[2], s, c;
  gca_oracle_philox_uniform2(seed, env, tick, slot, 0u, u);
  double r = sqrt(-2.0 * gca_oracle_log(1.0 - u[0], trig));
  gca_oracle_sincos(6.283185307179586 * u[1], trig, &s, &c);
  g[0] = r * c;
  g[1] = r * s;
}

/* ------------------------------------------------------------------------- one environment */
typedef struct env_view {
  const gca_config* cfg;
  int n;
  float* own_pos;
  double* own_hs;
  double* own_vel;
  uint8_t* own_vel_is_f32;
  double* goal;
  int32_t* no_conflict;
  int32_t* ep_steps;
  double* ipos;
  uint8_t* is64;
  float* ivel;
  uint8_t* iflag;
  double* ihs;   /* (heading, speed) per intruder; NULL unless the variant keeps them (intruder_turns) */
  /* draws */
  int draws, trig, f32_positions;
  const double* tape;
  int64_t* cursor;
  uint64_t seed;
  uint32_t tick, env_id;
  uint32_t* tick_store;
} env_view;

static double tape_next(env_view* e) { return e->tape[(*e->cursor)++]; }

/* np.random.uniform(low=[0,0], high=[W,H])  PKG/SingleAircraftEnv.py:240-244 */
static void draw_pos(env_view* e, uint32_t slot, uint32_t block, double* x, double* y) {
  if (e->draws == GCA_DRAWS_TAPE) {
    *x = tape_next(e);
    *y = tape_next(e);
  } else {
    double u[2];
    gca_oracle_philox_uniform2(e->seed, e->env_id, e->tick, slot, block, u);
    *x = 0.0 + (e->cfg->window_width - 0.0) * u[0];   /* numpy: low + (high - low) * u */
    *y = 0.0 + (e->cfg->window_height - 0.0) * u[1];
  }
}

/* random_speed(), random_heading()  PKG/SingleAircraftEnv.py:246-250 */
static void draw_speed_heading(env_view* e, uint32_t slot, double* speed, double* heading) {
  if (e->draws == GCA_DRAWS_TAPE) {
    *speed = tape_next(e);
    *heading = tape_next(e);
  } else {
    double u[2];
    gca_oracle_philox_uniform2(e->seed, e->env_id, e->tick, slot, GCA_BLOCK_SPEED_HEADING, u);
    *speed = e->cfg->min_speed + (e->cfg->max_speed - e->cfg->min_speed) * u[0];
    *heading = 0.0 + (6.283185307179586 - 0.0) * u[1];
  }
}

/* the two np.random.normal(0, sigma) calls of Ownship.step  PKG/SingleAircraftEnv.py:301,304 */
static void draw_own_noise(env_view* e, double* nh, double* ns) {
  if (e->draws == GCA_DRAWS_TAPE) {
    *nh = tape_next(e);
    *ns = tape_next(e);
  } else {
    double g[2];
    gca_oracle_philox_normal2(e->seed, e->env_id, e->tick, GCA_SLOT_OWNSHIP, e->trig, g);
    *nh = 0.0 + e->cfg->heading_sigma * g[0];
    *ns = 0.0 + e->cfg->speed_sigma * g[1];
  }
}

/* dist(): np.linalg.norm(p1 - p2)  PKG/SingleAircraftEnv.py:312-313.
 * f32 - f32: sqrtf(fl32(fl32(dx*dx) + fl32(dy*dy))), no FMA (OpenBLAS sdot).
 * anything with an f64 operand: sqrt(fma(dy, dy, fl(dx*dx))) (OpenBLAS ddot tail), SURVEY a6. */
static float dist_f32(float ax, float ay, float bx, float by) {
  float dx = ax - bx, dy = ay - by;
  float xx = dx * dx, yy = dy * dy;
  return sqrtf(xx + yy);
}

static double dist_f64(double ax, double ay, double bx, double by) {
  double dx = ax - bx, dy = ay - by;
  return sqrt(fma(dy, dy, dx * dx));
}

/* distance ownship <-> intruder i, returned as a double holding the exact f32 or f64 result */
static double dist_intruder(const env_view* e, int i) {
  if (e->is64[i]) return dist_f64((double)e->own_pos[0], (double)e->own_pos[1], e->ipos[2 * i], e->ipos[2 * i + 1]);
  return (double)dist_f32(e->own_pos[0], e->own_pos[1], (float)e->ipos[2 * i], (float)e->ipos[2 * i + 1]);
}

/* `d < threshold` where threshold is a Python float: an f32 d compares in f32 (NumPy 2 weak
 * scalars), an f64 d in f64. */
static int lt_thr(double d, int d_is_f64, double thr) { return d_is_f64 ? d < thr : (float)d < (float)thr; }

/* Aircraft(random_pos(), random_speed(), random_heading()) + rejection loop
 * PKG/SingleAircraftEnv.py:229-238,269-278 (reset: :80-88) */
static void spawn(env_view* e, int i, uint32_t slot) {
  double x, y, speed, heading, s, c;
  draw_pos(e, slot, GCA_BLOCK_POS, &x, &y);
  draw_speed_heading(e, slot, &speed, &heading);
  e->ipos[2 * i] = (double)(float)x;
  e->ipos[2 * i + 1] = (double)(float)y;
  e->is64[i] = 0;
  gca_oracle_sincos(heading, e->trig, &s, &c);
  e->ivel[2 * i] = (float)(speed * c);
  e->ivel[2 * i + 1] = (float)(speed * s);
  e->iflag[i] = 0;
  if (e->ihs) {   /* Aircraft.speed / Aircraft.heading  Simulators/SingleAircraftMCTSRandIntruderEnv.py:322-323 */
    e->ihs[2 * i] = heading;
    e->ihs[2 * i + 1] = speed;
  }
  int retries = 0;
  for (;;) {
    double d = dist_intruder(e, i);
    if (!lt_thr(d, e->is64[i], e->cfg->initial_min_dist)) break;
    if (e->draws == GCA_DRAWS_PHILOX && retries >= GCA_MAX_SPAWN_RETRIES) break;
    draw_pos(e, slot, GCA_BLOCK_RETRY0 + (uint32_t)retries, &x, &y);
    ++retries;
    if (e->f32_positions) {       /* FAST-mode storage rule */
      e->ipos[2 * i] = (double)(float)x;
      e->ipos[2 * i + 1] = (double)(float)y;
    } else {                      /* intruder.position = self.random_pos(): a raw f64 array (Q3) */
      e->ipos[2 * i] = x;
      e->ipos[2 * i + 1] = y;
      e->is64[i] = 1;
    }
  }
}

/* position_range.contains(p): low <= p <= high, inclusive, Box of f32 bounds  :38-41,:153 */
static int in_map(const gca_config* cfg, double x, double y) {
  double w = (double)(float)cfg->window_width, h = (double)(float)cfg->window_height;
  return x >= 0.0 && y >= 0.0 && x <= w && y <= h;
}

/* reset()  PKG/SingleAircraftEnv.py:66-98 */
static void reset_env(env_view* e) {
  const gca_config* cfg = e->cfg;
  double s, c;
  if (cfg->random_start) {   /* Ownship(random_pos(), random_speed(), random_heading())
                                Simulators/SingleAircraftDiscrete9HEREnv.py:78-82; Aircraft.__init__ :364-372 */
    double x, y, speed, heading;
    draw_pos(e, GCA_SLOT_OWN_RESET, GCA_BLOCK_POS, &x, &y);
    draw_speed_heading(e, GCA_SLOT_OWN_RESET, &speed, &heading);
    e->own_pos[0] = (float)x;
    e->own_pos[1] = (float)y;
    e->own_hs[0] = heading;
    e->own_hs[1] = speed;
  } else {
    e->own_pos[0] = 50.0f;
    e->own_pos[1] = 50.0f;
    e->own_hs[0] = 3.141592653589793 / 4;
    e->own_hs[1] = cfg->min_speed;
  }
  gca_oracle_sincos(e->own_hs[0], e->trig, &s, &c);
  e->own_vel[0] = (double)(float)(e->own_hs[1] * c);   /* Aircraft.__init__: f32 velocity */
  e->own_vel[1] = (double)(float)(e->own_hs[1] * s);
  *e->own_vel_is_f32 = 1;
  for (int i = 0; i < e->n; ++i) spawn(e, i, GCA_SLOT_RESET | (uint32_t)i);
  draw_pos(e, GCA_SLOT_GOAL, GCA_BLOCK_POS, &e->goal[0], &e->goal[1]);
  if (cfg->goal_margin > 0 && e->draws == GCA_DRAWS_PHILOX) {
    /* random_goal_pos(): uniform(low=[m, m], high=[W - m, H - m])  Simulators/SingleAircraftDiscrete3HEREnv.py:349-353;
     * draw_pos gave W * u: recover u is not exact, so draw again from the same block */
    double u[2];
    gca_oracle_philox_uniform2(e->seed, e->env_id, e->tick, GCA_SLOT_GOAL, GCA_BLOCK_POS, u);
    const double m = cfg->goal_margin;
    e->goal[0] = m + ((cfg->window_width - m) - m) * u[0];
    e->goal[1] = m + ((cfg->window_height - m) - m) * u[1];
  }
  *e->no_conflict = 0;
  *e->ep_steps = 0;
}

/* normalize_velocity()  PKG/SingleAircraftEnv.py:104-106 */
static double norm_vel_f32(const gca_config* cfg, float v) {
  float t = v + (float)cfg->max_speed;                 /* f32 + python float -> f32 */
  return (double)(t / (float)(cfg->max_speed * 2));
}

static double norm_vel_f64(const gca_config* cfg, double v) { return (v + cfg->max_speed) / (cfg->max_speed * 2); }

/* _get_ob()  PKG/SingleAircraftEnv.py:100-126, PKG/SingleAircraftHEREnv.py:103-139,
 * PKG/SingleAircraftDiscreteHEREnv.py:103-133, Simulators/SingleAircraftMCTSEnv.py:98-124 */
static void observe_nearest(const env_view* e, double* obs, double* ag, double* dg);

static void observe_env(const env_view* e, double* obs, double* ag, double* dg) {
  const gca_config* cfg = e->cfg;
  const int kind = cfg->obs_kind;
  if (kind == GCA_OBS_NONE) return;
  if (kind == GCA_OBS_NEAREST) {
    observe_nearest(e, obs, ag, dg);
    return;
  }
  const int raw = kind == GCA_OBS_RAW || kind == GCA_OBS_RAW6;
  const int per = kind == GCA_OBS_RAW6 ? 6 : 4;
  const int own_first = kind == GCA_OBS_HER || kind == GCA_OBS_DHER;
  double* oi = obs + (own_first ? 6 : 0);
  double* oo = obs + (own_first ? 0 : per * e->n);
  for (int i = 0; i < e->n; ++i) {
    double px = e->ipos[2 * i], py = e->ipos[2 * i + 1];
    float vx = e->ivel[2 * i], vy = e->ivel[2 * i + 1];
    if (kind == GCA_OBS_RAW6) {   /* (x, y, vx, vy, speed, heading)  Simulators/SingleAircraftMCTSRandIntruderEnv.py:133-140 */
      oi[6 * i] = px; oi[6 * i + 1] = py; oi[6 * i + 2] = vx; oi[6 * i + 3] = vy;
      oi[6 * i + 4] = e->ihs ? e->ihs[2 * i + 1] : 0.0;
      oi[6 * i + 5] = e->ihs ? e->ihs[2 * i] : 0.0;
    } else if (raw) {
      oi[4 * i] = px; oi[4 * i + 1] = py; oi[4 * i + 2] = vx; oi[4 * i + 3] = vy;
    } else {
      if (e->is64[i]) {
        oi[4 * i] = px / cfg->ob_window_width;
        oi[4 * i + 1] = py / cfg->ob_window_height;
      } else {
        oi[4 * i] = (double)((float)px / (float)cfg->ob_window_width);
        oi[4 * i + 1] = (double)((float)py / (float)cfg->ob_window_height);
      }
      oi[4 * i + 2] = norm_vel_f32(cfg, vx);
      oi[4 * i + 3] = norm_vel_f32(cfg, vy);
    }
  }
  if (raw) {
    oo[0] = e->own_pos[0]; oo[1] = e->own_pos[1];
    oo[2] = e->own_vel[0]; oo[3] = e->own_vel[1];
    oo[4] = e->own_hs[1]; oo[5] = e->own_hs[0];
  } else {
    oo[0] = (double)(e->own_pos[0] / (float)cfg->ob_window_width);
    oo[1] = (double)(e->own_pos[1] / (float)cfg->ob_window_height);
    if (*e->own_vel_is_f32) {
      oo[2] = norm_vel_f32(cfg, (float)e->own_vel[0]);
      oo[3] = norm_vel_f32(cfg, (float)e->own_vel[1]);
    } else {
      oo[2] = norm_vel_f64(cfg, e->own_vel[0]);
      oo[3] = norm_vel_f64(cfg, e->own_vel[1]);
    }
    oo[4] = (e->own_hs[1] - cfg->ob_min_speed) / (cfg->ob_max_speed - cfg->ob_min_speed);
    oo[5] = e->own_hs[0] / (2 * 3.141592653589793);
  }
  if (!own_first) {
    double* og = obs + per * e->n + 6;
    og[0] = raw ? e->goal[0] : e->goal[0] / cfg->ob_window_width;
    og[1] = raw ? e->goal[1] : e->goal[1] / cfg->ob_window_height;
  } else if (kind == GCA_OBS_HER) {
    ag[0] = (double)(e->own_pos[0] / (float)cfg->ob_window_width);
    ag[1] = (double)(e->own_pos[1] / (float)cfg->ob_window_height);
    dg[0] = e->goal[0] / cfg->ob_window_width;
    dg[1] = e->goal[1] / cfg->ob_window_height;
  } else {
    ag[0] = e->own_pos[0]; ag[1] = e->own_pos[1];
    dg[0] = e->goal[0]; dg[1] = e->goal[1];
  }
}

/* _get_ob() of Simulators/SingleAircraftDiscrete9HEREnv.py:106-165 (3HER identical): ownship (x, y, vx, vy), then the
 * Config.n nearest intruders, nearest first, each (x, y, vx, vy, dist / Config.diagonal).  dist_array is
 * np.array(dist_list) of f32 / f64 scalars: ordering by value; np.argpartition + argsort of distinct values = the n
 * smallest in ascending order (ties: lowest index first - they do not occur in the recorded traces). */
static void observe_nearest(const env_view* e, double* obs, double* ag, double* dg) {
  const gca_config* cfg = e->cfg;
  const int k = cfg->nearest_n;
  int idx[8];
  double dd[8];
  int m = 0;
  for (int i = 0; i < e->n; ++i) {
    const double d = dist_intruder(e, i);
    int pos = m;
    while (pos > 0 && d < dd[pos - 1]) --pos;             /* strict: an equal distance stays behind the earlier index */
    if (pos >= k) continue;
    const int last = m < k ? m : k - 1;
    for (int j = last; j > pos; --j) { dd[j] = dd[j - 1]; idx[j] = idx[j - 1]; }
    dd[pos] = d;
    idx[pos] = i;
    if (m < k) ++m;
  }
  obs[0] = (double)(e->own_pos[0] / (float)cfg->ob_window_width);
  obs[1] = (double)(e->own_pos[1] / (float)cfg->ob_window_height);
  if (*e->own_vel_is_f32) {
    obs[2] = norm_vel_f32(cfg, (float)e->own_vel[0]);
    obs[3] = norm_vel_f32(cfg, (float)e->own_vel[1]);
  } else {
    obs[2] = norm_vel_f64(cfg, e->own_vel[0]);
    obs[3] = norm_vel_f64(cfg, e->own_vel[1]);
  }
  for (int j = 0; j < m; ++j) {
    const int i = idx[j];
    double* o = obs + 4 + 5 * j;
    if (e->is64[i]) {
      o[0] = e->ipos[2 * i] / cfg->ob_window_width;
      o[1] = e->ipos[2 * i + 1] / cfg->ob_window_height;
      o[4] = dd[j] / cfg->ob_diagonal;
    } else {
      o[0] = (double)((float)e->ipos[2 * i] / (float)cfg->ob_window_width);
      o[1] = (double)((float)e->ipos[2 * i + 1] / (float)cfg->ob_window_height);
      o[4] = (double)((float)dd[j] / (float)cfg->ob_diagonal);
    }
    o[2] = norm_vel_f32(cfg, e->ivel[2 * i]);
    o[3] = norm_vel_f32(cfg, e->ivel[2 * i + 1]);
  }
  ag[0] = (double)(e->own_pos[0] / (float)cfg->ob_window_width);
  ag[1] = (double)(e->own_pos[1] / (float)cfg->ob_window_height);
  dg[0] = e->goal[0] / cfg->ob_window_width;
  dg[1] = e->goal[1] / cfg->ob_window_height;
}

/* Ownship.step(a)  PKG/SingleAircraftEnv.py:299-309 (+ 2Env :291-301, DiscreteHER :301-311) */
static void ownship_step(env_view* e, const double* action) {
  const gca_config* cfg = e->cfg;
  double f0, f1 = 0.0, nh, ns, s, c;
  if (cfg->action_kind == GCA_ACT_DISCRETE9) {
    int a = (int)action[0];
    f0 = (double)(a / 3) - 1.0;
    f1 = (double)(a % 3) - 1.0;
  } else if (cfg->action_kind == GCA_ACT_CONTINUOUS2) {
    f0 = action[0];
    f1 = action[1];
  } else {
    f0 = (double)((int)action[0] - 1);                  /* DISCRETE3 and DISCRETE3_HEADING */
  }
  draw_own_noise(e, &nh, &ns);
  double heading = e->own_hs[0], speed = e->own_hs[1];
  heading += cfg->d_heading * f0;
  heading += nh;
  if (cfg->action_kind == GCA_ACT_DISCRETE3) speed += cfg->speed_sigma;  /* reference quirk Q16 */
  else speed += cfg->d_speed * f1;
  double m = cfg->max_speed < speed ? cfg->max_speed : speed;           /* min(speed, max_speed) */
  speed = m > cfg->min_speed ? m : cfg->min_speed;                       /* max(min_speed, .) */
  speed += ns;
  gca_oracle_sincos(heading, e->trig, &s, &c);
  double vx = speed * c, vy = speed * s;
  e->own_hs[0] = heading;
  e->own_hs[1] = speed;
  e->own_vel[0] = vx;
  e->own_vel[1] = vy;
  *e->own_vel_is_f32 = 0;
  e->own_pos[0] = (float)((double)e->own_pos[0] + vx);                  /* f32 array += f64 array */
  e->own_pos[1] = (float)((double)e->own_pos[1] + vy);
}

/* _terminal_reward()  PKG/SingleAircraftEnv.py:143-184 and the variant rows of SURVEY.md 8(a) */
static void terminal_reward(env_view* e, double* reward, uint8_t* done, uint8_t* info, double* nearest) {
  const gca_config* cfg = e->cfg;
  /* self.dist_nearest_intruder = 9999 (Simulators/SingleAircraftDiscrete3HEREnv.py:185): value + dtype of the holder */
  double dnear = 9999.0;
  int near_is64 = 1, near_set = 0;
  if (nearest) *nearest = dnear;
  if (cfg->max_steps > 0 && *e->ep_steps >= cfg->max_steps) {            /* StackEnv :134-136 */
    *reward = 0.0; *done = 1; *info = GCA_INFO_MAXSTEPS;
    return;
  }
  int conflict = 0;
  for (int i = 0; i < e->n; ++i) {
    int is64 = e->is64[i];
    float vx = e->ivel[2 * i], vy = e->ivel[2 * i + 1];
    if (cfg->position_drift != 0.0) {
      /* intruder.position += intruder.velocity + self.position_sigma: the f32 velocity array plus a Python float stays
       * f32 (NumPy 2 weak scalars)  Simulators/SingleAircraftMCTSRandIntruderEnv.py:183 */
      vx = vx + (float)cfg->position_drift;
      vy = vy + (float)cfg->position_drift;
    }
    if (is64) {
      e->ipos[2 * i] = e->ipos[2 * i] + (double)vx;
      e->ipos[2 * i + 1] = e->ipos[2 * i + 1] + (double)vy;
    } else {
      e->ipos[2 * i] = (double)((float)e->ipos[2 * i] + vx);
      e->ipos[2 * i + 1] = (double)((float)e->ipos[2 * i + 1] + vy);
    }
    double d = dist_intruder(e, i);
    if (!(dnear < d)) {            /* min(dist_intruder, self.dist_nearest_intruder): the first argument wins ties :191 */
      dnear = d; near_is64 = is64; near_set = 1;
    }
    if (nearest) *nearest = dnear;
    int old_flag = e->iflag[i], replaced = 0;
    if (!in_map(cfg, e->ipos[2 * i], e->ipos[2 * i + 1])) {
      spawn(e, i, (uint32_t)i);                                          /* :153-154 */
      replaced = 1;
    }
    if (lt_thr(d, is64, cfg->minimum_separation)) {                      /* old distance, old object (Q7) */
      conflict = 1;
      if (!old_flag) {
        *e->no_conflict += 1;
        if (!replaced) e->iflag[i] = 1;                                  /* the write lands on the old object */
      }
      if (lt_thr(d, is64, cfg->nmac_dist)) {
        *reward = cfg->r_nmac; *done = 1; *info = GCA_INFO_NMAC;         /* later intruders untouched (Q9) */
        return;
      }
    }
  }
  if (conflict) {
    *reward = cfg->r_conflict; *done = 0; *info = GCA_INFO_CONFLICT;
    return;
  }
  if (cfg->wall_kind != GCA_WALL_NONE && !in_map(cfg, e->own_pos[0], e->own_pos[1])) {
    *reward = cfg->r_wall; *done = cfg->wall_kind == GCA_WALL_TERMINAL; *info = GCA_INFO_WALL;
    return;
  }
  double dg = dist_f64((double)e->own_pos[0], (double)e->own_pos[1], e->goal[0], e->goal[1]);
  if (dg < cfg->goal_radius) {
    *reward = cfg->r_goal; *done = 1; *info = GCA_INFO_GOAL;
    return;
  }
  *reward = cfg->shaped_default ? -dg / 1200 : cfg->r_default;
  if (cfg->shaped_nearest) {       /* :225-232 - NumPy 2 weak scalars: an f32 distance keeps the arithmetic in f32 */
    const double thr = 3 * cfg->minimum_separation;
    if (near_set ? lt_thr(dnear, near_is64, thr) : dnear < thr) {
      if (near_is64) {
        const double r = cfg->conflict_coeff * dnear - 0.1;
        *reward = cfg->shaped_default ? -dg / 1200 + r : cfg->r_default + r;
      } else {
        const float r = (float)((float)cfg->conflict_coeff * (float)dnear) - (float)0.1;
        *reward = cfg->shaped_default ? -dg / 1200 + (double)r : (double)((float)cfg->r_default + r);
      }
    }
  }
  *done = 0;
  *info = GCA_INFO_NONE;
}

/* _update_headings()  Simulators/SingleAircraftMCTSRandIntruderEnv.py:166-174 and Aircraft.change_heading :332-336:
 * for every intruder of the CURRENT list (a respawned one included) p = np.random.uniform(); if p < .1 the heading
 * moves by math.radians(np.random.uniform(-10, 10)) and the f32 velocity is rebuilt from (speed, heading).
 * math.radians(x) is x * (pi / 180) in C doubles. */
static void update_headings(env_view* e) {
  const gca_config* cfg = e->cfg;
  if (!e->ihs) return;
  for (int i = 0; i < e->n; ++i) {
    double p, raw;
    if (e->draws == GCA_DRAWS_TAPE) {
      p = tape_next(e);
      if (!(p < cfg->turn_prob)) continue;
      raw = tape_next(e);
    } else {
      double u[2];   /* one block serves the p of an intruder pair; the (rare) turn draws its angle from a block of its own */
      gca_oracle_philox_uniform2(e->seed, e->env_id, e->tick, GCA_SLOT_TURN | (uint32_t)(i >> 1), 0u, u);
      p = u[i & 1];
      if (!(p < cfg->turn_prob)) continue;
      gca_oracle_philox_uniform2(e->seed, e->env_id, e->tick, GCA_SLOT_TURN | (uint32_t)i, 1u, u);
      raw = -cfg->turn_max_deg + (cfg->turn_max_deg - -cfg->turn_max_deg) * u[0];   /* numpy: low + (high - low) * u */
    }
    double s, c;
    const double heading = e->ihs[2 * i] + raw * (3.141592653589793 / 180.0);
    const double speed = e->ihs[2 * i + 1];
    e->ihs[2 * i] = heading;
    gca_oracle_sincos(heading, e->trig, &s, &c);
    e->ivel[2 * i] = (float)(speed * c);
    e->ivel[2 * i + 1] = (float)(speed * s);
  }
}

static void bind(env_view* e, const gca_config* cfg, gca_oracle_batch* b, int i) {
  const int n = b->n_intr;
  e->cfg = cfg;
  e->n = n;
  e->own_pos = b->st.own_pos + 2 * (size_t)i;
  e->own_hs = b->st.own_hs + 2 * (size_t)i;
  e->own_vel = b->st.own_vel + 2 * (size_t)i;
  e->own_vel_is_f32 = b->st.own_vel_is_f32 + i;
  e->goal = b->st.goal + 2 * (size_t)i;
  e->no_conflict = b->st.no_conflict + i;
  e->ep_steps = b->st.ep_steps + i;
  e->ipos = b->st.ipos + 2 * (size_t)i * n;
  e->is64 = b->st.ipos_is_f64 + (size_t)i * n;
  e->ivel = b->st.ivel + 2 * (size_t)i * n;
  e->iflag = b->st.iflag + (size_t)i * n;
  e->ihs = b->st.ihs ? b->st.ihs + 2 * (size_t)i * n : NULL;
  e->draws = b->draws;
  e->trig = b->trig;
  e->f32_positions = b->f32_positions;
  e->tape = b->tape ? b->tape + (size_t)i * b->tape_stride : NULL;
  e->cursor = b->cursor ? b->cursor + i : NULL;
  e->seed = b->seed;
  e->tick_store = b->st.tick ? b->st.tick + i : NULL;
  e->tick = e->tick_store ? *e->tick_store : 0u;
  e->env_id = b->env_id0 + (uint32_t)i;
}

static double* row(double* p, int i, int w) { return p ? p + (size_t)i * w : NULL; }

int gca_oracle_step(const gca_config* cfg, gca_oracle_batch* b, const double* actions) {
  const int D = gca_oracle_obs_dim(cfg, b->n_intr);
  if (b->draws == GCA_DRAWS_TAPE && (!b->tape || !b->cursor)) return GCA_ERR_INVALID;
  for (int i = 0; i < b->n_envs; ++i) {
    env_view e;
    bind(&e, cfg, b, i);
    *e.ep_steps += 1;                                                    /* StackEnv :118 (a TimeLimit counter elsewhere) */
    ownship_step(&e, actions + 2 * (size_t)i);
    terminal_reward(&e, &b->reward[i], &b->done[i], &b->info[i], b->nearest ? &b->nearest[i] : NULL);
    if (cfg->time_limit > 0 && *e.ep_steps >= cfg->time_limit) b->done[i] = 1;   /* gym TimeLimit (registered ids) */
    if (cfg->intruder_turns) update_headings(&e);                        /* step(): after _terminal_reward, before _get_ob */
    observe_env(&e, row(b->obs, i, D), row(b->achieved, i, 2), row(b->desired, i, 2));
    if (b->term_obs && D) memcpy(row(b->term_obs, i, D), row(b->obs, i, D), sizeof(double) * D);
    if (b->auto_reset && b->done[i]) {                                   /* baselines dummy_vec_env.py:52-55 */
      reset_env(&e);
      observe_env(&e, row(b->obs, i, D), row(b->achieved, i, 2), row(b->desired, i, 2));
    }
    if (e.tick_store) *e.tick_store = e.tick + 1u;   /* an auto-reset shares the step's tick (RESET slots) */
  }
  return GCA_OK;
}

int gca_oracle_reset(const gca_config* cfg, gca_oracle_batch* b, const uint8_t* mask) {
  const int D = gca_oracle_obs_dim(cfg, b->n_intr);
  if (b->draws == GCA_DRAWS_TAPE && (!b->tape || !b->cursor)) return GCA_ERR_INVALID;
  for (int i = 0; i < b->n_envs; ++i) {
    if (mask && !mask[i]) continue;
    env_view e;
    bind(&e, cfg, b, i);
    reset_env(&e);
    observe_env(&e, row(b->obs, i, D), row(b->achieved, i, 2), row(b->desired, i, 2));
    if (e.tick_store) *e.tick_store = e.tick + 1u;
    if (b->done) b->done[i] = 0;
    if (b->info) b->info[i] = 0;
  }
  return GCA_OK;
}

int gca_oracle_observe(const gca_config* cfg, gca_oracle_batch* b) {
  const int D = gca_oracle_obs_dim(cfg, b->n_intr);
  for (int i = 0; i < b->n_envs; ++i) {
    env_view e;
    bind(&e, cfg, b, i);
    observe_env(&e, row(b->obs, i, D), row(b->achieved, i, 2), row(b->desired, i, 2));
  }
  return GCA_OK;
}

/* compute_reward()  PKG/SingleAircraftHEREnv.py:194-196, PKG/SingleAircraftDiscreteHEREnv.py:184-186:
 * d = np.linalg.norm(ag - g, axis=-1) -> sqrt(dx*dx + dy*dy) reduced pairwise in f64 (no BLAS) */
int gca_oracle_compute_reward_f32(const float* ag, const float* g, int64_t m, double radius, int kind, float* out) {
  for (int64_t i = 0; i < m; ++i) {   /* f32 inputs: numpy keeps the whole norm in f32 and compares against f32(radius) */
    float dx = ag[2 * i] - g[2 * i], dy = ag[2 * i + 1] - g[2 * i + 1];
    float xx = dx * dx, yy = dy * dy;
    float d = sqrtf(xx + yy);
    if (kind == GCA_OBS_HER) out[i] = -(float)(d > (float)radius);
    else out[i] = (float)(d < (float)radius);
  }
  return GCA_OK;
}

int gca_oracle_compute_reward(const double* ag, const double* g, int64_t m, double radius, int kind, float* out) {
  for (int64_t i = 0; i < m; ++i) {
    double dx = ag[2 * i] - g[2 * i], dy = ag[2 * i + 1] - g[2 * i + 1];
    double d = sqrt(dx * dx + dy * dy);
    if (kind == GCA_OBS_HER) out[i] = -(float)(d > radius);
    else out[i] = (float)(d < radius);
  }
  return GCA_OK;
}

/* the oracle carries its own copy of the product's gca_obs_dim so that it never links against libgca */
int gca_oracle_obs_dim(const gca_config* cfg, int n_intruders) {
  switch (cfg->obs_kind) {
    case GCA_OBS_VECTOR:
    case GCA_OBS_RAW: return 4 * n_intruders + 8;
    case GCA_OBS_HER:
    case GCA_OBS_DHER: return 4 * n_intruders + 6;
    case GCA_OBS_NEAREST: return 4 + 5 * cfg->nearest_n;
    case GCA_OBS_RAW6: return 6 * n_intruders + 8;
    default: return 0;
  }
}
