/* gca_oracle.h - CPU restatement of the reference's step hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  Nothing under gym-guidance-collision-avoidance-single_b200/ links,
 * imports or calls it.
 *
 * Pinning: the restatement is checked bit-for-bit (flags, counters, positions, rewards,
 * observations) against traces recorded from the unmodified reference by
 * tests/golden/make_golden.py (tests/test_oracle_golden.py).  The reference has no tests or
 * golden vectors of its own.
 */
#ifndef GCA_ORACLE_H_
#define GCA_ORACLE_H_

#include <stdint.h>
#include "gca.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { GCA_TRIG_LIBM = 0,    /* cos/sin/log from libm: what CPython's math module calls */
       GCA_TRIG_SHARED = 1 };/* gca_math.h: bit-identical to the CUDA kernels */

typedef struct gca_oracle_batch {
  int32_t n_envs, n_intr;
  gca_host_state st;          /* in/out, canonical layout (include/gca.h) */
  /* draw source */
  int32_t draws;              /* GCA_DRAWS_* */
  int32_t trig;               /* GCA_TRIG_* */
  const double* tape;         /* [B][tape_stride] */
  int64_t tape_stride;
  int64_t* cursor;            /* [B] in/out */
  uint64_t seed;            /* Philox key; the per-env tick lives in st.tick */
  uint32_t env_id0;
  uint32_t reserved0;
  int32_t f32_positions;      /* 1: FAST-mode storage rule (a retried spawn is rounded to f32) */
  int32_t auto_reset;
  /* outputs (always f64-valued; an element the reference computes in f32 is that f32 widened) */
  double* obs;                /* [B][obs_dim] */
  double* achieved;           /* [B][2] or NULL */
  double* desired;            /* [B][2] or NULL */
  double* reward;             /* [B] */
  uint8_t* done;              /* [B] */
  uint8_t* info;              /* [B] */
  double* term_obs;           /* [B][obs_dim] or NULL: observation of the step itself, before any auto-reset */
  double* nearest;            /* [B] or NULL: dist_nearest_intruder (shaped_nearest variants) */
} gca_oracle_batch;

/* actions: double[B][2]; discrete kinds use actions[b][0] as the integer code. */
int gca_oracle_step(const gca_config* cfg, gca_oracle_batch* b, const double* actions);
int gca_oracle_reset(const gca_config* cfg, gca_oracle_batch* b, const uint8_t* mask);
/* observation of the current state without stepping (PKG/SingleAircraftEnv.py:100-126) */
int gca_oracle_observe(const gca_config* cfg, gca_oracle_batch* b);

/* HER relabel reward, same contract as gca_compute_reward but on host arrays */
int gca_oracle_compute_reward(const double* ag, const double* g, int64_t m, double radius, int kind, float* out);
int gca_oracle_compute_reward_f32(const float* ag, const float* g, int64_t m, double radius, int kind, float* out);
int gca_oracle_obs_dim(const gca_config* cfg, int n_intruders);

/* building blocks exposed for unit tests */
void gca_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void gca_oracle_sincos(double x, int trig, double* s, double* c);
double gca_oracle_log(double x, int trig);
/* the two uniforms / the Box-Muller pair a Philox block yields */
void gca_oracle_philox_uniform2(uint64_t seed, uint32_t env, uint32_t tick, uint32_t slot, uint32_t block, double u[2]);
void gca_oracle_philox_normal2(uint64_t seed, uint32_t env, uint32_t tick, uint32_t slot, int trig, double g[2]);

/* ---- MCTS forward model and search (oracle/gca_oracle_mcts.c) ---- */
int gca_oracle_mcts_move(const gca_mcts_config* c, int n, double* state, int action, int draws, int trig,
                         const double* tape, int64_t* cursor, uint64_t seed, uint32_t root, int first_frame,
                         uint8_t* flags, double* reward);
int gca_oracle_mcts_rollout(const gca_mcts_config* c, int n, const double* root, int depth_limit, int draws, int trig,
                            const double* tape, int64_t* cursor, uint64_t seed, uint32_t root_id, uint32_t playout,
                            int forced_first, double* reward, int8_t* first_out, uint8_t* flags_out);
int gca_oracle_mcts_playouts(const gca_mcts_config* c, int n, const double* roots, int64_t n_roots, int playouts,
                             int depth, const int8_t* first_action, uint64_t seed, uint32_t root_id0, int trig,
                             double* rewards, int8_t* first_out, uint8_t* flags);
int gca_oracle_mcts_search(const gca_mcts_config* c, int n, const double* root, int sims, int search_depth,
                           const double* tape, int64_t* cursor, int trig, int* best_action, double* child_n,
                           double* child_q, int* child_action);
int gca_oracle_mcts_search_philox(const gca_mcts_config* c, int n, const double* roots, int64_t n_roots, int sims,
                                  int search_depth, uint64_t seed, uint32_t root_id0, int trig, int32_t* best_action,
                                  double* child_n, double* child_q, int32_t* child_action);

/* ---- StackEnv image observation (oracle/gca_oracle_raster.c); trig of `b` selects libm / shared sincos ---- */
int gca_oracle_raster(const gca_config* cfg, const gca_oracle_batch* b, const unsigned char* sprites,
                      unsigned char* frames, unsigned char* rgb);

#ifdef __cplusplus
}
#endif
#endif
