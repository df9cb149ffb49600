/* gca_oracle_mcts.c - CPU restatement of the MCTS forward model and search of the reference
 * (Algorithms/MCTS/nodes_single.py, search_single.py, common.py, config_single.py).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see gca_oracle.h).  Pinned against tests/golden/mcts_n*.npz,
 * recorded from the unmodified reference: every move(), rollout() and whole best_action() search
 * is replayed from the recorded numpy draws.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gca_oracle.h"
#include "gca_math.h"

typedef struct mdraws {
  int mode;                 /* GCA_DRAWS_TAPE / GCA_DRAWS_PHILOX */
  int trig;
  const double* tape;
  int64_t* cursor;
  uint64_t seed;
  uint32_t root, playout;
} mdraws;

static double mt_next(mdraws* d) { return d->tape[(*d->cursor)++]; }

/* np.random.normal(0, sigma) of the model; `what`/`idx` address the Philox block */
static double draw_normal(mdraws* d, double sigma, uint32_t what, uint32_t idx) {
  if (d->mode == GCA_DRAWS_TAPE) return mt_next(d);
  if (sigma == 0.0) return 0.0;                         /* the reference still draws; the value is +-0 (Q26) */
  double u[2], s, c;
  gca_oracle_philox_uniform2(d->seed, d->root, d->playout, what, idx, u);
  const double r = sqrt(-2.0 * gca_oracle_log(1.0 - u[0], d->trig));
  gca_oracle_sincos(6.283185307179586 * u[1], d->trig, &s, &c);
  return 0.0 + sigma * (r * c);
}

/* np.random.randint(9) of rollout_policy (common.py:54-55) */
static int draw_action(mdraws* d, uint32_t move) {
  if (d->mode == GCA_DRAWS_TAPE) return (int)mt_next(d);
  double u[2];
  gca_oracle_philox_uniform2(d->seed, d->root, d->playout, GCA_MCTS_DRAW_ACTION, move, u);
  int a = (int)(9.0 * u[0]);
  return a > 8 ? 8 : a;
}

/* np.random.random() < turn_prob -> math.radians(np.random.uniform(-10, 10))  nodes_single_randintru.py:64-65.
 * Returns 1 and the heading change when the intruder turns. */
static int draw_turn(mdraws* d, const gca_mcts_config* c, uint32_t intruder, uint32_t gf, double* delta) {
  double p, raw;
  if (d->mode == GCA_DRAWS_TAPE) {
    p = mt_next(d);
    if (!(p < c->turn_prob)) return 0;
    raw = mt_next(d);
  } else {
    double u[2];
    gca_oracle_philox_uniform2(d->seed, d->root, d->playout, GCA_MCTS_DRAW_TURN + intruder, gf, u);
    p = u[0];
    if (!(p < c->turn_prob)) return 0;
    raw = -c->turn_max_deg + (c->turn_max_deg - -c->turn_max_deg) * u[1];
  }
  *delta = raw * (3.141592653589793 / 180.0);
  return 1;
}

/* length of a state vector and entries per intruder of the two models */
static int per_intruder(const gca_mcts_config* c) { return c->random_intruders ? 6 : 4; }
static int state_len(const gca_mcts_config* c, int n) { return per_intruder(c) * n + 8; }

static double metric(double x1, double y1, double x2, double y2) {   /* nodes_single.py:123-126 */
  double dx = x1 - x2, dy = y1 - y2;
  return sqrt(dx * dx + dy * dy);
}

/* SingleAircraftState.move(action)  nodes_single.py:39-100.  `frame0` numbers the sub-frames of a
 * playout globally (Philox index); returns the GCA_MCTS_* flags of the successor. */
static int model_move(const gca_mcts_config* c, int n, double* st, int a0, int a1, mdraws* d, int frame0) {
  const int per = per_intruder(c);
  const int L = state_len(c, n);
  /* (len - 9) // 4 = N - 1: the last intruder is ignored (Q22); nodes_single_randintru.py:47 has (len - 8) // 6 = N */
  const int near = c->random_intruders ? n : (L >= 9 ? (L - 9) / 4 : 0);
  const double d_heading = (double)(a0 - 1) * c->d_heading;
  const double accel = (double)(a1 - 1) * c->d_speed;
  double* own = st + per * n;                            /* x y vx vy speed heading */
  double* goal = st + per * n + 6;
  int flags = 0;
  for (int f = 0; f < c->simulate_frame; ++f) {
    const uint32_t gf = (uint32_t)(frame0 + f);
    for (int i = 0; i < near; ++i) {                     /* :54-57 (nodes_single_randintru.py:54-71) */
      double* it = st + per * i;
      it[0] += it[2] + draw_normal(d, c->position_sigma, GCA_MCTS_DRAW_INTRUDER + (uint32_t)i, 2 * gf);
      it[1] += it[3] + draw_normal(d, c->position_sigma, GCA_MCTS_DRAW_INTRUDER + (uint32_t)i, 2 * gf + 1);
      double delta;
      if (c->random_intruders && draw_turn(d, c, (uint32_t)i, gf, &delta)) {
        double s, co;
        const double heading = it[5] + delta;
        gca_oracle_sincos(heading, d->trig, &s, &co);
        it[2] = it[4] * co;
        it[3] = it[4] * s;
        it[5] = heading;
      }
    }
    own[4] += accel;                                     /* :59 - overwritten by the next line in nodes_single.py */
    {
      const double v = c->random_intruders ? own[4] : own[3];           /* min(state[-5], max): own vy (Q23); the
                                                                            random-intruder model clamps the speed (:74) */
      const double m = c->max_speed < v ? c->max_speed : v;
      own[4] = m > c->min_speed ? m : c->min_speed;
    }
    own[4] += draw_normal(d, c->speed_sigma, GCA_MCTS_DRAW_SPEED, gf);
    const double speed = own[4];
    own[5] += d_heading;
    own[5] += draw_normal(d, c->heading_sigma, GCA_MCTS_DRAW_HEADING, gf);
    const double heading = own[5];
    double s, co;
    gca_oracle_sincos(heading, d->trig, &s, &co);
    const double vx = speed * co, vy = speed * s;
    own[0] += vx;
    own[1] += vy;
    own[2] = vx;
    own[3] = vy;
    const double ox = own[0], oy = own[1];
    if (!(0 < ox && ox < c->window_width) || !(0 < oy && oy < c->window_height)) {   /* strict (Q6) */
      flags |= GCA_MCTS_WALL;
      break;
    }
    int conflict = 0;
    for (int i = 0; i < near; ++i)
      if (metric(st[per * i], st[per * i + 1], ox, oy) < c->minimum_separation) {
        conflict = 1;
        break;
      }
    if (conflict) {
      flags |= GCA_MCTS_CONFLICT;
      break;
    }
    if (metric(ox, oy, goal[0], goal[1]) < c->minimum_separation) {      /* goal radius = separation (Q24) */
      flags |= GCA_MCTS_GOAL;
      break;
    }
  }
  return flags;
}

/* reward()  nodes_single.py:25-32 */
static double model_reward(const gca_mcts_config* c, int n, const double* st, int flags) {
  if (flags & (GCA_MCTS_WALL | GCA_MCTS_CONFLICT)) return 0.0;
  if (flags & GCA_MCTS_GOAL) return 1.0;
  const double* own = st + per_intruder(c) * n;
  const double dx = own[0] - own[6], dy = own[1] - own[7];
  return 1 - sqrt(dx * dx + dy * dy) / 1200.0;
}

int gca_oracle_mcts_move(const gca_mcts_config* c, int n, double* state, int action, int draws, int trig,
                         const double* tape, int64_t* cursor, uint64_t seed, uint32_t root, int first_frame,
                         uint8_t* flags, double* reward) {
  mdraws d = {draws, trig, tape, cursor, seed, root, 0u};
  const int f = model_move(c, n, state, action / 3, action % 3, &d, first_frame);
  if (flags) *flags = (uint8_t)f;
  if (reward) *reward = model_reward(c, n, state, f);
  return GCA_OK;
}

/* rollout(search_depth) from a state at depth `depth0`   nodes_single.py:198-204 */
static double model_rollout(const gca_mcts_config* c, int n, double* st, int flags, int depth0, int depth_limit,
                            mdraws* d, int forced_first, int* first_out, int* flags_out) {
  int depth = depth0, first = -1;
  while (!(flags || depth == depth_limit)) {
    int a = (depth == depth0 && forced_first >= 0) ? forced_first : draw_action(d, (uint32_t)depth);
    if (first < 0) first = a;
    flags = model_move(c, n, st, a / 3, a % 3, d, depth * c->simulate_frame);
    ++depth;
  }
  if (first_out) *first_out = first;
  if (flags_out) *flags_out = flags;
  return model_reward(c, n, st, flags);
}

int gca_oracle_mcts_rollout(const gca_mcts_config* c, int n, const double* root, int depth_limit, int draws, int trig,
                            const double* tape, int64_t* cursor, uint64_t seed, uint32_t root_id, uint32_t playout,
                            int forced_first, double* reward, int8_t* first_out, uint8_t* flags_out) {
  const int L = state_len(c, n);
  double* st = (double*)malloc(sizeof(double) * L);
  if (!st) return GCA_ERR_ALLOC;
  memcpy(st, root, sizeof(double) * L);
  mdraws d = {draws, trig, tape, cursor, seed, root_id, playout};
  int first = -1, fl = 0;
  *reward = model_rollout(c, n, st, 0, 0, depth_limit, &d, forced_first, &first, &fl);
  if (first_out) *first_out = (int8_t)first;
  if (flags_out) *flags_out = (uint8_t)fl;
  free(st);
  return GCA_OK;
}

/* the batched-playout contract of gca_mcts_playouts (include/gca.h), on host arrays */
int gca_oracle_mcts_playouts(const gca_mcts_config* c, int n, const double* roots, int64_t n_roots, int playouts,
                             int depth, const int8_t* first_action, uint64_t seed, uint32_t root_id0, int trig,
                             double* rewards, int8_t* first_out, uint8_t* flags) {
  const int L = state_len(c, n);
  for (int64_t r = 0; r < n_roots; ++r)
    for (int p = 0; p < playouts; ++p) {
      const int64_t k = r * playouts + p;
      int rc = gca_oracle_mcts_rollout(c, n, roots + r * L, depth, GCA_DRAWS_PHILOX, trig, NULL, NULL, seed,
                                       root_id0 + (uint32_t)r, (uint32_t)p, first_action ? first_action[k] : -1,
                                       &rewards[k], first_out ? &first_out[k] : NULL, flags ? &flags[k] : NULL);
      if (rc) return rc;
    }
  return GCA_OK;
}

/* ------------------------------------------------------------------------------ the search */
typedef struct node {
  int parent, depth, flags, action;
  int n_children, children[9];
  int untried;                /* actions 0..untried-1 are still untried; expand() pops from the end (Q27) */
  double q, n;
  double* state;
} node;

static int is_terminal(const node* v, int search_depth) { return v->flags || v->depth == search_depth; }

/* best_child(c_param)  common.py:47-52 */
static int best_child(const node* nodes, int v, double c_param, int shared_log, int trig) {
  int best = -1;
  double best_w = 0;
  for (int k = 0; k < nodes[v].n_children; ++k) {
    const node* ch = &nodes[nodes[v].children[k]];
    /* np.log of the reference; the device search shares gca_math.h's log, so the Philox search does too */
    const double lg = shared_log ? gca_oracle_log(nodes[v].n, trig) : log(nodes[v].n);
    const double w = (ch->q / ch->n) + c_param * sqrt((2 * lg / ch->n));
    if (best < 0 || w > best_w) {     /* np.argmax: first maximum */
      best = nodes[v].children[k];
      best_w = w;
    }
  }
  return best;
}

/* MCTS(root).best_action(simulations, search_depth)  search_single.py:8-22 */
static int search_impl(const gca_mcts_config* c, int n, const double* root, int sims, int search_depth, mdraws d,
                       int* best_action, double* child_n, double* child_q, int* child_action) {
  const int L = state_len(c, n);
  const int philox = d.mode == GCA_DRAWS_PHILOX;
  node* nodes = (node*)calloc((size_t)sims + 2, sizeof(node));
  double* scratch = (double*)malloc(sizeof(double) * L);
  if (!nodes || !scratch) return GCA_ERR_ALLOC;
  int count = 1;
  nodes[0].parent = -1;
  nodes[0].untried = 9;
  nodes[0].action = -1;
  nodes[0].state = (double*)malloc(sizeof(double) * L);
  memcpy(nodes[0].state, root, sizeof(double) * L);
  for (int s = 0; s < sims; ++s) {
    d.playout = (uint32_t)s;                                       /* Philox: simulation s is "playout" s of the root */
    int v = 0;                                                     /* tree_policy */
    while (!is_terminal(&nodes[v], search_depth)) {
      if (nodes[v].untried > 0) {                                  /* expand(): nodes_single.py:188-193 */
        const int a = --nodes[v].untried;
        node* ch = &nodes[count];
        ch->parent = v;
        ch->depth = nodes[v].depth + 1;
        ch->action = a;
        ch->untried = 9;
        ch->state = (double*)malloc(sizeof(double) * L);
        memcpy(ch->state, nodes[v].state, sizeof(double) * L);
        /* sub-frames are numbered from the root (Philox index; the tape ignores it) */
        ch->flags = model_move(c, n, ch->state, a / 3, a % 3, &d, nodes[v].depth * c->simulate_frame);
        nodes[v].children[nodes[v].n_children++] = count;
        v = count++;
        break;
      }
      v = best_child(nodes, v, 1.4, philox, d.trig);
    }
    memcpy(scratch, nodes[v].state, sizeof(double) * L);           /* rollout */
    const double r = model_rollout(c, n, scratch, nodes[v].flags, nodes[v].depth, search_depth, &d, -1, NULL, NULL);
    for (int u = v; u >= 0; u = nodes[u].parent) {                 /* backpropagate: nodes_single.py:206-210 */
      nodes[u].n += 1.;
      nodes[u].q += r;
    }
  }
  const int b = best_child(nodes, 0, 0., philox, d.trig);
  *best_action = b >= 0 ? nodes[b].action : -1;
  for (int k = 0; k < 9; ++k) {
    const int has = k < nodes[0].n_children;
    if (child_n) child_n[k] = has ? nodes[nodes[0].children[k]].n : 0;
    if (child_q) child_q[k] = has ? nodes[nodes[0].children[k]].q : 0;
    if (child_action) child_action[k] = has ? nodes[nodes[0].children[k]].action : -1;
  }
  for (int i = 0; i < count; ++i) free(nodes[i].state);
  free(nodes);
  free(scratch);
  return GCA_OK;
}

int gca_oracle_mcts_search(const gca_mcts_config* c, int n, const double* root, int sims, int search_depth,
                           const double* tape, int64_t* cursor, int trig, int* best_action, double* child_n,
                           double* child_q, int* child_action) {
  mdraws d = {GCA_DRAWS_TAPE, trig, tape, cursor, 0, 0, 0};
  return search_impl(c, n, root, sims, search_depth, d, best_action, child_n, child_q, child_action);
}

/* the contract of gca_mcts_search (include/gca.h): Philox draws keyed (seed, root id, simulation, kind, sub-frame) */
int gca_oracle_mcts_search_philox(const gca_mcts_config* c, int n, const double* roots, int64_t n_roots, int sims,
                                  int search_depth, uint64_t seed, uint32_t root_id0, int trig, int32_t* best_action,
                                  double* child_n, double* child_q, int32_t* child_action) {
  const int L = state_len(c, n);
  for (int64_t r = 0; r < n_roots; ++r) {
    mdraws d = {GCA_DRAWS_PHILOX, trig, NULL, NULL, seed, root_id0 + (uint32_t)r, 0};
    int best = -1, ca[9];
    int rc = search_impl(c, n, roots + r * L, sims, search_depth, d, &best, child_n ? child_n + 9 * r : NULL,
                         child_q ? child_q + 9 * r : NULL, ca);
    if (rc) return rc;
    best_action[r] = best;
    if (child_action)
      for (int k = 0; k < 9; ++k) child_action[9 * r + k] = ca[k];
  }
  return GCA_OK;
}
