/* pgtg_oracle.c -- CPU restatement of the reference PGTG tick, reset and map build.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * `--impl reference` legs may load this. The product (pgtg_b200) never links or imports it.
 *
 * Parity status: PINNED. tests/test_oracle_golden.py replays traces recorded from the unmodified
 * reference run in the build container (tests/golden/make_golden.py, using the recorded
 * np_random draws) and requires bit-equal observations, rewards, flags, agent and car state.
 * The only third-party algorithm restated is graph-theory 2022.4.3 (module `graph`, pinned in
 * the reference's poetry.lock:1115-1116; source unavailable offline): insertion-ordered
 * adjacency, edges() order, FIFO BFS, Dijkstra with (cost, push counter) heap keys. It is
 * anchored on the reference's own golden trajectory (tests/test_data/reproducibility_data.py)
 * and on the recorded traces, whose maps come from the shim in oracle/shims/graph.py; the shim
 * and this file agree, the real package could only differ in Dijkstra tie-breaking on maps with
 * non-default start/goal.
 *
 * Style: deliberately literal -- a dense per-square feature grid, explicit starter/spawner lists
 * built by the same x-major scan, ordered adjacency lists -- i.e. the reference's data
 * structures, not the product's packed tile descriptors. Every function cites the reference
 * lines it follows (paths relative to /root/reference/pgtg/).
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, so the float64
 * rounding in _decompose_velocity matches CPython/numpy).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pgtg_b200.h"
#include "pgtg_oracle_tables.h"

#define TW 9
#define TH 9

/* feature bits of one square (the reference's set[str]) */
enum {
  F_WALL = 1, F_SUBGOAL = 2, F_USED = 4, F_START = 8, F_FINAL = 16, F_ICE = 32, F_BROKEN = 64,
  F_SAND = 128, F_LIGHT = 256, F_SPAWNER = 512, F_LANE = 1024
};

typedef struct { int id, x, y, route, profile, patience, delay; } ora_car;

typedef struct {
  /* MapPlan (map_generator.py:10-17) */
  int W, H;
  unsigned char exits[PGTG_MAX_TILES], otype[PGTG_MAX_TILES], omask[PGTG_MAX_TILES];
  int sx, sy, sdir, gx, gy, gdir;
  /* EpisodeMap (map.py:8-42) */
  int width, height;
  uint16_t* grid;          /* [x * height + y] feature bits */
  unsigned char* sq_type;  /* tile type of the square's tile (lane lookups) */
  int num_subgoals;
  signed char tile_dir[PGTG_MAX_TILES]; /* tile_coordinates_to_subgoal_directions, -1 none */
  int n_starters, n_spawnable, n_spawners;
  int16_t (*starters)[2];
  int16_t (*spawnable)[2];
  int16_t (*spawners)[2];
  /* PGTGEnv episode state (environment.py:631-650) */
  double individual_subgoal_reward;
  int x, y, vx, vy;
  int terminated, truncated, flat_tire;
  int light_counter;
  int elapsed;
  int n_cars, next_car_id;
  ora_car* cars;
  int cars_cap;
  unsigned char* visited; /* positions_path as a bitmap over [-1,width] x [-1,height] */
  /* rng */
  uint64_t seed;     /* philox key */
  uint32_t episode;
  uint32_t draw_k[5];       /* words consumed per stream this tick */
  uint32_t blk[4], blk_index; /* cached Philox block of blk_stream */
  int blk_stream;
  uint32_t cblk[4], cblk_index; /* cached car-stream block of car slot cblk_slot */
  int cblk_slot;
  /* numpy mode: the five PCG64 children of this reset (environment.py:593-599) */
  struct { unsigned __int128 state, inc; uint32_t buf; int has; } pcg[5];
  int64_t cursor, tape_end;
  int error;
  /* outputs of the last step */
  int braking_applied;
  int outcome; /* of the last tick: 0 running, 1 crash, 2 final goal */
  double ep_return;
} ora_env;

typedef struct ora_batch {
  pgtg_config cfg;
  int N, P, C, T;
  ora_env* envs;
  /* fixed map */
  int have_fixed;
  pgtg_tile fixed_tiles[PGTG_MAX_TILES];
  int fw, fh, fsx, fsy, fsdir, fgx, fgy, fgdir;
  /* tape */
  const double* tape_values;
  const uint8_t* tape_tags;
  double stats[8];
  int threads;
} ora_batch;

/* ------------------------------------------------------------------------------------------ */
/* Random draws. Semantic API shared with the product (include/pgtg_b200.h):
 *   tape mode   : values recorded from the reference's five np_random children
 *                 (environment.py:593-599), one tape per env in program order;
 *   philox mode : Philox4x32-10, key = env seed, counter = (k, tick, episode, stream).     */

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

void ora_philox(uint32_t ctr[4], uint32_t k0, uint32_t k1) { philox4x32_10(ctr, k0, k1); }

typedef struct { ora_batch* b; ora_env* e; } rctx;

/* Philox word stream: block b = philox(counter = (b, tick, episode, stream), key = seed), 4 words per
 * block; an index draw takes one word, a double two consecutive words. */
static uint32_t philox_word(rctx* r, int stream) {
  ora_env* e = r->e;
  uint32_t pos = e->draw_k[stream]++;
  uint32_t b = pos >> 2;
  if (e->blk_stream != stream || e->blk_index != b) {
    e->blk[0] = b; e->blk[1] = (uint32_t)e->elapsed; e->blk[2] = e->episode; e->blk[3] = (uint32_t)stream;
    philox4x32_10(e->blk, (uint32_t)e->seed, (uint32_t)(e->seed >> 32));
    e->blk_stream = stream; e->blk_index = b;
  }
  return e->blk[pos & 3u];
}

/* Philox CAR stream (product specification, pgtg_b200/csrc/pgtg_device.cuh CW_*): every car owns one block
 * sequence per tick, block b of car slot s = philox(counter = (b, tick, episode, STREAM_CAR | (s + 1) << 8), key),
 * s = the car's list index when the tick starts, with fixed word meanings; a car-stream uniform is one word
 * * 2^-32. Initial traffic: car slot j takes lane square perm(j) (4-round Feistel over [0, 4^h), keys = the
 * block of slot -1, cycle-walked into [0, n)); profile and route from words 0 and 1 of its own tick-0 block. */
enum { CW_DELAY = 0, CW_SPEED = 1, CW_IDX = 2, CW_PUSH = 3, CW_LIGHT = 4, CW_SPAWNER = 5, CW_SPAWN_ROUTE = 6, CW_PROFILE = 7,
       CW0_PROFILE = 0, CW0_ROUTE = 1 };
static uint32_t philox_car_word(rctx* r, int slot, int pos) {
  ora_env* e = r->e;
  uint32_t b = (uint32_t)pos >> 2;
  if (e->cblk_slot != slot || e->cblk_index != b) {
    e->cblk[0] = b; e->cblk[1] = (uint32_t)e->elapsed; e->cblk[2] = e->episode;
    e->cblk[3] = (uint32_t)PGTG_STREAM_CAR | (uint32_t)(slot + 1) << 8;
    philox4x32_10(e->cblk, (uint32_t)e->seed, (uint32_t)(e->seed >> 32));
    e->cblk_slot = slot; e->cblk_index = b;
  }
  return e->cblk[pos & 3];
}
static uint32_t feistel_mix(uint32_t x, uint32_t k) {
  uint32_t t = (x ^ k) * 0x9E3779B1u;
  t ^= t >> 15; t *= 0x85EBCA6Bu; t ^= t >> 13;
  return t;
}
/* Feistel network over [0, 2^m), m = bits of n - 1 (at least 2), halves of a = m / 2 and m - a bits, cycle-walked into [0, n) */
static int initial_car_position(const uint32_t keys[4], int n, int slot) {
  int m = 2;
  while ((1 << m) < n) m++;
  int a = m >> 1, b = m - a;
  uint32_t ma = (1u << a) - 1u, mb = (1u << b) - 1u, v = (uint32_t)slot;
  do {
    uint32_t lo = v & ma, hi = v >> a;
    lo ^= feistel_mix(hi, keys[0]) & ma;
    hi ^= feistel_mix(lo, keys[1]) & mb;
    lo ^= feistel_mix(hi, keys[2]) & ma;
    hi ^= feistel_mix(lo, keys[3]) & mb;
    v = hi << a | lo;
  } while (v >= (uint32_t)n);
  return (int)v;
}

/* ---- numpy mode (PGTG_RNG_NUMPY): SeedSequence + PCG64 + Generator methods, restated from numpy's
 * published algorithms (bit_generator.pyx SeedSequence, pcg64.h, distributions.c bounded Lemire,
 * _generator.pyx choice) for exactly the calls the reference makes; pinned against numpy itself by
 * tests/test_numpy_rng.py through the golden traces. ------------------------------------------------ */
static uint32_t ss_hashmix(uint32_t v, uint32_t* hc) { v ^= *hc; *hc *= 0x931e8875u; v *= *hc; v ^= v >> 16; return v; }
static uint32_t ss_mix(uint32_t x, uint32_t y) { uint32_t q = 0xca01f9ddu * x - 0x4973f715u * y; q ^= q >> 16; return q; }
#define PCG_MULT ((((unsigned __int128)0x2360ED051FC65DA4ull) << 64) | 0x4385DF649FCCF645ull)
static void np_seed_child(ora_env* e, int stream, uint32_t child) {
  uint32_t ent[5] = {(uint32_t)e->seed, (uint32_t)(e->seed >> 32), 0, 0, child}, pool[4], hc = 0x43b0d7e5u;
  for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(ent[i], &hc);
  for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) if (a != b) pool[b] = ss_mix(pool[b], ss_hashmix(pool[a], &hc));
  for (int b = 0; b < 4; b++) pool[b] = ss_mix(pool[b], ss_hashmix(ent[4], &hc));
  uint32_t w[8], hb = 0x8b51f9ddu;
  for (int i = 0; i < 8; i++) { uint32_t v = pool[i & 3] ^ hb; hb *= 0x58f38dedu; v *= hb; v ^= v >> 16; w[i] = v; }
  unsigned __int128 initstate = ((unsigned __int128)((uint64_t)w[0] | (uint64_t)w[1] << 32) << 64) | ((uint64_t)w[2] | (uint64_t)w[3] << 32);
  unsigned __int128 initseq = ((unsigned __int128)((uint64_t)w[4] | (uint64_t)w[5] << 32) << 64) | ((uint64_t)w[6] | (uint64_t)w[7] << 32);
  e->pcg[stream].inc = (initseq << 1) | 1u;
  e->pcg[stream].state = 0;
  e->pcg[stream].state = e->pcg[stream].state * PCG_MULT + e->pcg[stream].inc;
  e->pcg[stream].state += initstate;
  e->pcg[stream].state = e->pcg[stream].state * PCG_MULT + e->pcg[stream].inc;
  e->pcg[stream].has = 0; e->pcg[stream].buf = 0;
}
static uint64_t np_next64(ora_env* e, int s) {
  e->pcg[s].state = e->pcg[s].state * PCG_MULT + e->pcg[s].inc;
  uint64_t hi = (uint64_t)(e->pcg[s].state >> 64), lo = (uint64_t)e->pcg[s].state, x = hi ^ lo;
  unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}
static uint32_t np_next32(ora_env* e, int s) {
  if (e->pcg[s].has) { e->pcg[s].has = 0; return e->pcg[s].buf; }
  uint64_t n = np_next64(e, s);
  e->pcg[s].has = 1; e->pcg[s].buf = (uint32_t)(n >> 32);
  return (uint32_t)n;
}
static uint32_t np_bounded(ora_env* e, int s, uint32_t rng) { /* inclusive bound, Lemire 32-bit */
  if (rng == 0) return 0;
  uint32_t ex = rng + 1u;
  uint64_t m = (uint64_t)np_next32(e, s) * ex;
  uint32_t left = (uint32_t)m;
  if (left < ex) {
    uint32_t thr = (0u - ex) % ex;
    while (left < thr) { m = (uint64_t)np_next32(e, s) * ex; left = (uint32_t)m; }
  }
  return (uint32_t)(m >> 32);
}

static double tape_next(rctx* r, int stream, int kind) {
  ora_env* e = r->e;
  if (e->cursor >= e->tape_end) { e->error |= 1; return 0.0; }
  if (r->b->tape_tags[e->cursor] != (uint8_t)(stream * 8 + kind)) { e->error |= 2; }
  return r->b->tape_values[e->cursor++];
}

/* Generator.random() */
static double rng_double(rctx* r, int stream) {
  if (r->b->cfg.rng_mode == PGTG_RNG_TAPE) return tape_next(r, stream, PGTG_DRAW_DOUBLE);
  if (r->b->cfg.rng_mode == PGTG_RNG_NUMPY) return (double)(np_next64(r->e, stream) >> 11) * (1.0 / 9007199254740992.0);
  uint32_t w0 = philox_word(r, stream), w1 = philox_word(r, stream);
  return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) / 9007199254740992.0;
}

/* Generator.integers(0, n) / Generator.choice over n items; numpy consumes nothing for n == 1 */
static int rng_index(rctx* r, int stream, int n) {
  if (n <= 1) return 0;
  if (r->b->cfg.rng_mode == PGTG_RNG_TAPE) {
    int v = (int)tape_next(r, stream, PGTG_DRAW_INDEX);
    if (v < 0 || v >= n) { r->e->error |= 4; v = 0; }
    return v;
  }
  if (r->b->cfg.rng_mode == PGTG_RNG_NUMPY) return (int)np_bounded(r->e, stream, (uint32_t)n - 1u);
  return (int)(((uint64_t)philox_word(r, stream) * (uint64_t)n) >> 32);
}

/* Generator.choice(items, p=...): one uniform double, cdf.searchsorted(u, side="right") */
static int rng_choice_cdf(rctx* r, int stream, const double* cdf, int n) {
  if (r->b->cfg.rng_mode == PGTG_RNG_TAPE) {
    int v = (int)tape_next(r, stream, PGTG_DRAW_INDEX);
    if (v < 0 || v >= n) { r->e->error |= 4; v = 0; }
    return v;
  }
  double u = rng_double(r, stream);
  int i = 0;
  while (i < n - 1 && cdf[i] <= u) i++;
  return i;
}

/* car-stream draws: the reference's sequential car_rng order in tape / numpy mode, the per-car blocks in Philox mode */
static double rng_car_double(rctx* r, int slot, int pos) {
  if (r->b->cfg.rng_mode != PGTG_RNG_PHILOX) return rng_double(r, PGTG_STREAM_CAR);
  return (double)philox_car_word(r, slot, pos) * (1.0 / 4294967296.0);
}
static int rng_car_index(rctx* r, int slot, int pos, int n) {
  if (n <= 1) return 0;
  if (r->b->cfg.rng_mode != PGTG_RNG_PHILOX) return rng_index(r, PGTG_STREAM_CAR, n);
  return (int)(((uint64_t)philox_car_word(r, slot, pos) * (uint64_t)n) >> 32);
}
static int rng_car_choice_cdf(rctx* r, int slot, int pos, const double* cdf, int n) {
  if (r->b->cfg.rng_mode != PGTG_RNG_PHILOX) return rng_choice_cdf(r, PGTG_STREAM_CAR, cdf, n);
  double u = rng_car_double(r, slot, pos);
  int i = 0;
  while (i < n - 1 && cdf[i] <= u) i++;
  return i;
}

/* Generator.choice(n, size=k, replace=False): k distinct indices in returned order.
 * Philox mode: the keyed Feistel permutation of the car-stream specification above. */
static void rng_distinct(rctx* r, int stream, int n, int k, int* out) {
  if (r->b->cfg.rng_mode == PGTG_RNG_PHILOX) {
    uint32_t keys[4];
    for (int i = 0; i < 4; i++) keys[i] = philox_car_word(r, -1, i);
    for (int j = 0; j < k; j++) out[j] = initial_car_position(keys, n, j);
    return;
  }
  if (r->b->cfg.rng_mode == PGTG_RNG_NUMPY) { /* Floyd's sampling + shuffle (_generator.pyx choice) */
    for (int t = 0; t < k; t++) {
      uint32_t j = (uint32_t)(n - k + t), v = np_bounded(r->e, stream, j);
      int dup = 0;
      for (int q = 0; q < t; q++) if (out[q] == (int)v) { dup = 1; break; }
      out[t] = dup ? (int)j : (int)v;
    }
    for (int i = k - 1; i >= 1; i--) { int jj = (int)np_bounded(r->e, stream, (uint32_t)i); int a = out[i]; out[i] = out[jj]; out[jj] = a; }
    return;
  }
  for (int j = 0; j < k; j++) {
    int v = (int)tape_next(r, stream, PGTG_DRAW_INDEX);
    if (v < 0 || v >= n) { r->e->error |= 4; v = 0; }
    out[j] = v;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* graph-theory restatement: ordered adjacency (see oracle/shims/graph.py and the header). */

#define GMAXN (PGTG_MAX_TILES + 2)
typedef struct {
  int n_nodes;
  int order[GMAXN];        /* node insertion order */
  unsigned char known[GMAXN];
  int deg[GMAXN];
  int nbr[GMAXN][6];       /* successors in insertion order */
} ograph;

static void g_init(ograph* g) { memset(g, 0, sizeof *g); }
static void g_add_node(ograph* g, int a) {
  if (!g->known[a]) { g->known[a] = 1; g->order[g->n_nodes++] = a; }
}
static void g_add_edge1(ograph* g, int a, int b) {
  g_add_node(g, a); g_add_node(g, b);
  for (int i = 0; i < g->deg[a]; i++) if (g->nbr[a][i] == b) return;
  g->nbr[a][g->deg[a]++] = b;
}
static void g_del_edge(ograph* g, int a, int b) {
  for (int i = 0; i < g->deg[a]; i++) if (g->nbr[a][i] == b) {
    for (int j = i; j + 1 < g->deg[a]; j++) g->nbr[a][j] = g->nbr[a][j + 1];
    g->deg[a]--; return;
  }
}
static int g_edge_count(const ograph* g) {
  int c = 0;
  for (int i = 0; i < g->n_nodes; i++) c += g->deg[g->order[i]];
  return c;
}
/* breadth_first_search(start, end): FIFO, first-discovered predecessor; returns path length */
static int g_bfs(const ograph* g, int s, int t, int* path) {
  int prev[GMAXN], q[GMAXN], qh = 0, qt = 0;
  for (int i = 0; i < GMAXN; i++) prev[i] = -2;
  prev[s] = -1; q[qt++] = s;
  while (qh < qt) {
    int n = q[qh++];
    if (n == t) {
      int len = 0, tmp[GMAXN];
      while (n != -1) { tmp[len++] = n; n = prev[n]; }
      for (int i = 0; i < len; i++) path[i] = tmp[len - 1 - i];
      return len;
    }
    for (int i = 0; i < g->deg[n]; i++) {
      int m = g->nbr[n][i];
      if (prev[m] == -2) { prev[m] = n; q[qt++] = m; }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* map_generator.py */

static int card_dirs_at(int W, int H, int x, int y, int* out) {
  /* chose_random_start_or_goal_direction (map_generator.py:571-597): north,east,south,west */
  int n = 0;
  if (y == 0) out[n++] = 0;
  if (x == W - 1) out[n++] = 1;
  if (y == H - 1) out[n++] = 2;
  if (x == 0) out[n++] = 3;
  return n;
}

static void random_border_position(rctx* r, int W, int H, int* x, int* y) {
  /* chose_random_start_or_goal_position (map_generator.py:600-626) */
  switch (rng_index(r, PGTG_STREAM_MAP, 4)) {
    case 0: *x = rng_index(r, PGTG_STREAM_MAP, W); *y = 0; break;
    case 1: *x = W - 1; *y = rng_index(r, PGTG_STREAM_MAP, H); break;
    case 2: *x = rng_index(r, PGTG_STREAM_MAP, W); *y = H - 1; break;
    default: *x = 0; *y = rng_index(r, PGTG_STREAM_MAP, H); break;
  }
}

static int random_direction(rctx* r, int W, int H, int x, int y) {
  int d[4];
  int n = card_dirs_at(W, H, x, y, d);
  return d[rng_index(r, PGTG_STREAM_MAP, n)];
}

static void choose_start_goal(rctx* r, ora_env* e) {
  /* chose_random_start_and_goal_position_and_direction (map_generator.py:475-568).
   * "len == 2" in the reference means "direction still missing". */
  const pgtg_config* c = &r->b->cfg;
  int W = e->W, H = e->H;
  int sx = c->start_x, sy = c->start_y, sd = c->start_dir, s_has_dir = (c->start_mode == 0);
  int gx = c->goal_x, gy = c->goal_y, gd = c->goal_dir, g_has_dir = (c->goal_mode == 0);
  if (c->start_mode == 2) random_border_position(r, W, H, &sx, &sy);
  if (c->goal_mode == 2) random_border_position(r, W, H, &gx, &gy);
  if (c->min_start_goal_distance >= 0) {
    while (abs(sx - gx) + abs(sy - gy) < c->min_start_goal_distance) {
      random_border_position(r, W, H, &sx, &sy);
      random_border_position(r, W, H, &gx, &gy);
    }
  }
  if (!s_has_dir) { sd = random_direction(r, W, H, sx, sy); }
  if (!g_has_dir) { gd = random_direction(r, W, H, gx, gy); }
  while (sx == gx && sy == gy && sd == gd) {
    if (c->start_mode == 2) random_border_position(r, W, H, &sx, &sy);
    if (c->start_mode == 2 || c->start_mode == 1) sd = random_direction(r, W, H, sx, sy);
    if (c->goal_mode == 2) random_border_position(r, W, H, &gx, &gy);
    if (c->goal_mode == 2 || c->goal_mode == 1) gd = random_direction(r, W, H, gx, gy);
  }
  e->sx = sx; e->sy = sy; e->sdir = sd; e->gx = gx; e->gy = gy; e->gdir = gd;
}

static void generate_map(rctx* r, ora_env* e) {
  /* generate_map (map_generator.py:43-189) */
  const pgtg_config* c = &r->b->cfg;
  int W = c->map_w, H = c->map_h;
  e->W = W; e->H = H;
  choose_start_goal(r, e);

  /* generate_map_graph (map_generator.py:192-266). Node id = x * H + y for (x, y);
   * "start" = W*H, "end" = W*H + 1. */
  static __thread ograph g;
  g_init(&g);
  for (int x = 0; x < W; x++)
    for (int y = 0; y < H; y++) {
      if (x < W - 1) { g_add_edge1(&g, x * H + y, (x + 1) * H + y); g_add_edge1(&g, (x + 1) * H + y, x * H + y); }
      if (y < H - 1) { g_add_edge1(&g, x * H + y, x * H + y + 1); g_add_edge1(&g, x * H + y + 1, x * H + y); }
    }
  /* removable_edges = edges() order: outer node insertion order, inner successor order (:227) */
  static __thread int rem[4 * PGTG_MAX_TILES][2];
  int n_rem = 0;
  for (int i = 0; i < g.n_nodes; i++) {
    int a = g.order[i];
    for (int j = 0; j < g.deg[a]; j++) { rem[n_rem][0] = a; rem[n_rem][1] = g.nbr[a][j]; n_rem++; }
  }
  int S = W * H, E = W * H + 1;
  int sn = e->sx * H + e->sy, gn = e->gx * H + e->gy;
  g_add_edge1(&g, S, sn); g_add_edge1(&g, sn, S);
  g_add_edge1(&g, E, gn); g_add_edge1(&g, gn, E);
  int keep = c->edges_to_keep;
  int path[GMAXN];
  int plen = g_bfs(&g, S, E, path);
  if (c->rng_mode == PGTG_RNG_PHILOX) {
    /* Philox specification (pgtg_logic.cuh remove_edges_tabled): the pick is over the grid edges not tried yet (what the
     * reference's draw over removable_edges amounts to: both directions of an edge are listed and leave the list
     * together), kept in an array that starts in the order of the product's connectivity bits -- horizontal edges row by
     * row, then vertical edges by tile index y * W + x. A trip takes one word of the map stream (also when one edge is
     * left), i = (word * n) >> 32, tries edge a[i] and closes the gap with the last one: a[i] = a[n-1], n -= 1. */
    n_rem = 0;
    for (int y = 0; y < H; y++) for (int x = 0; x + 1 < W; x++) { rem[n_rem][0] = x * H + y; rem[n_rem][1] = (x + 1) * H + y; n_rem++; }
    for (int y = 0; y + 1 < H; y++) for (int x = 0; x < W; x++) { rem[n_rem][0] = x * H + y; rem[n_rem][1] = x * H + y + 1; n_rem++; }
  }
  while (g_edge_count(&g) - 4 > keep && n_rem > 0) { /* :245 */
    int a, b;
    if (c->rng_mode == PGTG_RNG_PHILOX) {
      int idx = (int)(((uint64_t)philox_word(r, PGTG_STREAM_MAP) * (uint32_t)n_rem) >> 32); /* :249 */
      a = rem[idx][0]; b = rem[idx][1];
      rem[idx][0] = rem[n_rem - 1][0]; rem[idx][1] = rem[n_rem - 1][1];
      n_rem--;
    } else {
      int idx = rng_index(r, PGTG_STREAM_MAP, n_rem); /* :249 */
      a = rem[idx][0]; b = rem[idx][1];
      /* removable_edges.remove(chosen); .remove(reverse) (:252-253) */
      int w = 0;
      for (int i = 0; i < n_rem; i++) {
        if ((rem[i][0] == a && rem[i][1] == b) || (rem[i][0] == b && rem[i][1] == a)) continue;
        rem[w][0] = rem[i][0]; rem[w][1] = rem[i][1]; w++;
      }
      n_rem = w;
    }
    g_del_edge(&g, a, b); g_del_edge(&g, b, a);
    int ina = 0, inb = 0;
    for (int i = 0; i < plen; i++) { if (path[i] == a) ina = 1; if (path[i] == b) inb = 1; }
    if (ina && inb) { /* :258-264 */
      int np[GMAXN];
      int nl = g_bfs(&g, S, E, np);
      if (nl > 0) { plen = nl; memcpy(path, np, sizeof(int) * nl); }
      else { g_add_edge1(&g, a, b); g_add_edge1(&g, b, a); }
    }
  }
  /* map_graph_to_tile_map_object (map_generator.py:269-334) */
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      int ex = 0, a = x * H + y;
      for (int j = 0; j < g.deg[a]; j++) {
        int m = g.nbr[a][j];
        if (m >= S) continue;
        int mx = m / H, my = m % H;
        if (mx == x && my == y - 1) ex |= 1;
        if (mx == x + 1 && my == y) ex |= 2;
        if (mx == x && my == y + 1) ex |= 4;
        if (mx == x - 1 && my == y) ex |= 8;
      }
      e->exits[y * W + x] = (unsigned char)ex;
      e->otype[y * W + x] = 0; e->omask[y * W + x] = 0;
    }
  e->exits[e->sy * W + e->sx] |= (unsigned char)(1 << e->sdir);
  e->exits[e->gy * W + e->gx] |= (unsigned char)(1 << e->gdir);

  /* add_connections_to_borders (map_generator.py:337-371): rows (tile_y, tile_x, dir) */
  static __thread int bc[4 * 64][3];
  int nb = 0;
  for (int x = 0; x < W; x++) { bc[nb][0] = 0; bc[nb][1] = x; bc[nb][2] = 0; nb++; }
  for (int y = 0; y < H; y++) { bc[nb][0] = y; bc[nb][1] = W - 1; bc[nb][2] = 1; nb++; }
  for (int x = 0; x < W; x++) { bc[nb][0] = H - 1; bc[nb][1] = x; bc[nb][2] = 2; nb++; }
  for (int y = 0; y < H; y++) { bc[nb][0] = y; bc[nb][1] = 0; bc[nb][2] = 3; nb++; }
  for (int pass = 0; pass < 2; pass++) { /* always the DEFAULT start/goal slots (:359-360) */
    int ty = pass == 0 ? H - 1 : 0, tx = pass == 0 ? 0 : W - 1, d = pass == 0 ? 3 : 1;
    for (int i = 0; i < nb; i++) if (bc[i][0] == ty && bc[i][1] == tx && bc[i][2] == d) {
      for (int j = i; j + 1 < nb; j++) memcpy(bc[j], bc[j + 1], sizeof bc[0]);
      nb--; break;
    }
  }
  if (c->rng_mode == PGTG_RNG_PHILOX) {
    /* Philox specification (pgtg_logic.cuh choose_border_slots): the picks are a uniformly random subset of the slots,
     * drawn by rejection: the stream's next words are cut into chunks of ceil(log2 n) bits, lowest first, floor(32 / bits)
     * per word, the rest of a word dropped; a chunk >= n or naming a chosen slot is skipped. */
    int left = c->border_connections < nb ? c->border_connections : nb, bbits = 1, bleft = 0;
    unsigned char chosen[4 * 64] = {0};
    uint32_t bw = 0;
    while (bbits < 31 && (1 << bbits) < nb) bbits++;
    while (left > 0) {
      if (bleft == 0) { bw = philox_word(r, PGTG_STREAM_MAP); bleft = 32 / bbits; }
      int v = (int)(bw & ((1u << bbits) - 1u)); /* :367 */
      bw >>= bbits; bleft--;
      if (v >= nb || chosen[v]) continue;
      chosen[v] = 1; left--;
      e->exits[bc[v][0] * W + bc[v][1]] |= (unsigned char)(1 << bc[v][2]);
    }
  } else
  for (int k = 0; k < c->border_connections; k++) {
    int idx = rng_index(r, PGTG_STREAM_MAP, nb); /* :367 */
    e->exits[bc[idx][0] * W + bc[idx][1]] |= (unsigned char)(1 << bc[idx][2]);
    for (int j = idx; j + 1 < nb; j++) memcpy(bc[j], bc[j + 1], sizeof bc[0]);
    nb--;
  }

  /* add_obstacles_to_map (map_generator.py:374-472) */
  if (c->obstacle_probability > 0) {
    for (int row = 0; row < H; row++)
      for (int col = 0; col < W; col++) {
        int t = row * W + col;
        double u = rng_double(r, PGTG_STREAM_MAP); /* always drawn (:415) */
        if (!(u < c->obstacle_probability) || e->exits[t] == 0) continue;
        int type = 1 + rng_choice_cdf(r, PGTG_STREAM_MAP, c->obstacle_cdf, 4); /* :418 */
        e->otype[t] = (unsigned char)type;
        if (type != 4) {
          e->omask[t] = (unsigned char)rng_index(r, PGTG_STREAM_MAP, 8); /* :430 */
        } else {
          int opts[6], n = 0, ex = e->exits[t];
          int cnt = (ex & 1) + ((ex >> 1) & 1) + ((ex >> 2) & 1) + ((ex >> 3) & 1);
          if (ex & 1) opts[n++] = 8;
          if (ex & 2) opts[n++] = 9;
          if (ex & 4) opts[n++] = 10;
          if (ex & 8) opts[n++] = 11;
          if ((ex & 1) && (ex & 4) && cnt >= 3) opts[n++] = 12;
          if ((ex & 2) && (ex & 8) && cnt >= 3) opts[n++] = 13;
          e->omask[t] = (unsigned char)opts[rng_index(r, PGTG_STREAM_MAP, n)]; /* :470 */
        }
      }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* parser.py + map.py */

static int find_direction(int ax, int ay, int bx, int by) {
  /* find_direction (parser.py:279-306): 0 north 1 east 2 south 3 west */
  if (ay == by) { if (ax - bx < 0) return 1; if (ax - bx > 0) return 3; }
  if (ax == bx) { if (ay - by < 0) return 2; if (ay - by > 0) return 0; }
  return -1;
}

static void replace_exit(unsigned char tile[9][9], uint16_t feat[9][9], int dir, int newf) {
  /* replace_features_in_tile(tile, "exit <dir>", new) (parser.py:169-190) */
  for (int x = 0; x < 9; x++)
    for (int y = 0; y < 9; y++)
      if (tile[x][y] & (2 << dir)) { tile[x][y] &= (unsigned char)~(2 << dir); feat[x][y] |= (uint16_t)newf; }
}

static void build_episode_map(ora_batch* b, ora_env* e) {
  /* parse_map_object (parser.py:13-166) then EpisodeMap.__init__ (map.py:11-42) */
  int W = e->W, H = e->H;
  /* parse_tile_map_to_graph (parser.py:244-276): directed edges from each tile's own exits,
   * added in N, E, S, W order; node id = y * W + x */
  static __thread ograph g;
  g_init(&g);
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      int a = i * W + j, ex = e->exits[a];
      g_add_node(&g, a);
      if ((ex & 1) && i > 0) g_add_edge1(&g, a, (i - 1) * W + j);
      if ((ex & 2) && j < W - 1) g_add_edge1(&g, a, i * W + j + 1);
      if ((ex & 4) && i < H - 1) g_add_edge1(&g, a, (i + 1) * W + j);
      if ((ex & 8) && j > 0) g_add_edge1(&g, a, i * W + j - 1);
    }
  /* shortest_path: Dijkstra, unit weights, (cost, push counter) keys == FIFO BFS with
   * first-discovered predecessor (graph-theory 2022.4.3, see header) */
  int path[GMAXN];
  int plen = g_bfs(&g, e->sy * W + e->sx, e->gy * W + e->gx, path);
  if (plen == 0) { e->error |= 8; plen = 1; path[0] = e->sy * W + e->sx; }
  for (int t = 0; t < W * H; t++) e->tile_dir[t] = -1;
  for (int i = 0; i + 1 < plen; i++)
    e->tile_dir[path[i]] = (signed char)find_direction(path[i] % W, path[i] / W, path[i + 1] % W, path[i + 1] / W);

  e->width = W * TW; e->height = H * TH;
  memset(e->grid, 0, sizeof(uint16_t) * (size_t)e->width * e->height);
  for (int tx = 0; tx < W; tx++)
    for (int ty = 0; ty < H; ty++) {
      int t = ty * W + tx, ex = e->exits[t];
      unsigned char tile[9][9];
      uint16_t feat[9][9];
      memcpy(tile, ORA_TILE[ex], sizeof tile); /* copy.deepcopy(TILES[exits]) (:51-53) */
      memset(feat, 0, sizeof feat);
      int on_path_not_last = 0;
      for (int i = 0; i + 1 < plen; i++) if (path[i] == t) on_path_not_last = 1;
      if (on_path_not_last) replace_exit(tile, feat, e->tile_dir[t], F_SUBGOAL);   /* :55-61 */
      if (t == path[0]) replace_exit(tile, feat, e->sdir, F_START);              /* :63-69 */
      if (t == path[plen - 1]) replace_exit(tile, feat, e->gdir, F_FINAL);       /* :71-77 */
      /* remaining exit markers are dropped (:79-99) */
      for (int x = 0; x < 9; x++)
        for (int y = 0; y < 9; y++) {
          if (tile[x][y] & 1) feat[x][y] |= F_WALL;
          if (e->otype[t] && ORA_MASK[e->omask[t]][x][y] && !(tile[x][y] & 1)) { /* :101-111, 193-209 */
            static const int of[5] = {0, F_ICE, F_BROKEN, F_SAND, F_LIGHT};
            feat[x][y] |= (uint16_t)of[e->otype[t]];
          }
          if (ex != 0) { /* add_traffic_lanes_to_tile (:113-118) */
            const ora_lane_sq* l = &ORA_LANES[ex][x][y];
            if (l->n || l->all) feat[x][y] |= F_LANE;
            if (l->spawner) feat[x][y] |= F_SPAWNER;
            /* border car spawners on 'car_lane all <inward>' squares (:120-148) */
            if (tx == 0 && l->all == 4) feat[x][y] |= F_SPAWNER;       /* all right */
            if (tx == W - 1 && l->all == 3) feat[x][y] |= F_SPAWNER;   /* all left */
            if (ty == 0 && l->all == 2) feat[x][y] |= F_SPAWNER;       /* all down */
            if (ty == H - 1 && l->all == 1) feat[x][y] |= F_SPAWNER;   /* all up */
          }
          e->grid[(tx * TW + x) * e->height + ty * TH + y] = feat[x][y];
          e->sq_type[(tx * TW + x) * e->height + ty * TH + y] = (unsigned char)ex;
        }
    }
  e->tile_dir[path[plen - 1]] = (signed char)e->gdir; /* :158 */
  e->num_subgoals = plen;                              /* len(dict) (:164) */

  /* EpisodeMap.__init__ scan, x outer / y inner (map.py:31-42) */
  e->n_starters = e->n_spawnable = e->n_spawners = 0;
  for (int x = 0; x < e->width; x++)
    for (int y = 0; y < e->height; y++) {
      int f = e->grid[x * e->height + y];
      if (f & F_START) { e->starters[e->n_starters][0] = (int16_t)x; e->starters[e->n_starters][1] = (int16_t)y; e->n_starters++; }
      if (f & F_LANE) { e->spawnable[e->n_spawnable][0] = (int16_t)x; e->spawnable[e->n_spawnable][1] = (int16_t)y; e->n_spawnable++; }
      if (f & F_SPAWNER) { e->spawners[e->n_spawners][0] = (int16_t)x; e->spawners[e->n_spawners][1] = (int16_t)y; e->n_spawners++; }
    }
  if (b->cfg.rng_mode == PGTG_RNG_PHILOX) {
    /* Philox specification (product, pgtg_logic.cuh): both index spaces of the car stream enumerate tile by tile
     * (t = ty * W + tx ascending) -- lane squares by local index lx * 9 + ly, car spawners by slot: the native
     * one, then the border spawners of the north, east, south and west map border. */
    e->n_spawnable = e->n_spawners = 0;
    for (int t = 0; t < W * H; t++) {
      int tx = t % W, ty = t / W, ex = e->exits[t];
      for (int x = 0; x < 9; x++)
        for (int y = 0; y < 9; y++)
          if (e->grid[(tx * TW + x) * e->height + ty * TH + y] & F_LANE) {
            e->spawnable[e->n_spawnable][0] = (int16_t)(tx * TW + x); e->spawnable[e->n_spawnable][1] = (int16_t)(ty * TH + y); e->n_spawnable++;
          }
      if (ex == 0) continue;
      for (int slot = 0; slot < 5; slot++)
        for (int x = 0; x < 9; x++)
          for (int y = 0; y < 9; y++) {
            const ora_lane_sq* l = &ORA_LANES[ex][x][y];
            int hit = slot == 0 ? l->spawner : slot == 1 ? (ty == 0 && l->all == 2) : slot == 2 ? (tx == W - 1 && l->all == 3)
                    : slot == 3 ? (ty == H - 1 && l->all == 1) : (tx == 0 && l->all == 4);
            if (hit) { e->spawners[e->n_spawners][0] = (int16_t)(tx * TW + x); e->spawners[e->n_spawners][1] = (int16_t)(ty * TH + y); e->n_spawners++; }
          }
    }
  }
}

static inline int inside_map(const ora_env* e, int x, int y) {
  return !(x < 0 || y < 0 || x >= e->width || y >= e->height); /* map.py:44-47 */
}
static inline int feat_at(const ora_env* e, int x, int y) { return e->grid[x * e->height + y]; }
static inline const ora_lane_sq* lanes_at(const ora_env* e, int x, int y) {
  return &ORA_LANES[e->sq_type[x * e->height + y]][x % TW][y % TH];
}

static void set_subgoals_to_used(ora_env* e, int x, int y) {
  /* map.py:143-171 (the reference would raise outside the map; subgoal lines never touch it) */
  e->grid[x * e->height + y] = (uint16_t)((e->grid[x * e->height + y] & ~F_SUBGOAL) | F_USED);
  if (inside_map(e, x, y + 1) && (feat_at(e, x, y + 1) & F_SUBGOAL)) set_subgoals_to_used(e, x, y + 1);
  if (inside_map(e, x, y - 1) && (feat_at(e, x, y - 1) & F_SUBGOAL)) set_subgoals_to_used(e, x, y - 1);
  if (inside_map(e, x + 1, y) && (feat_at(e, x + 1, y) & F_SUBGOAL)) set_subgoals_to_used(e, x + 1, y);
  if (inside_map(e, x - 1, y) && (feat_at(e, x - 1, y) & F_SUBGOAL)) set_subgoals_to_used(e, x - 1, y);
}

/* ------------------------------------------------------------------------------------------ */
/* environment.py: traffic */

static int light_phase(const pgtg_config* c, int counter) {
  /* get_traffic_light_phase (environment.py:1004-1015): 0 green 1 yellow 2 red */
  if (counter < c->light_green) return 0;
  if (counter < c->light_green + c->light_yellow) return 1;
  return 2;
}

static int select_profile(rctx* r, int slot, int pos) { /* _select_driver_profile (:658-662) */
  return rng_car_choice_cdf(r, slot, pos, r->b->cfg.profile_cdf, PGTG_NUM_PROFILES);
}

static int random_route_at(rctx* r, int x, int y, int slot, int pos) {
  /* sorted route names of the square's non-"all" lanes, then car_rng.choice (:854-874, 982-997) */
  const ora_lane_sq* l = lanes_at(r->e, x, y);
  if (l->n == 0) { r->e->error |= 16; return 0; }
  return l->route[rng_car_index(r, slot, pos, l->n)];
}

static void create_initial_traffic(rctx* r) {
  /* _create_initial_traffic (environment.py:830-879) */
  ora_env* e = r->e;
  int num_positions = e->n_spawnable;
  int num_cars = (int)((double)num_positions * r->b->cfg.traffic_density);
  if (num_cars > num_positions) num_cars = num_positions;
  if (num_cars <= 0 || num_positions <= 0) return;
  if (num_cars > e->cars_cap) { e->error |= 32; num_cars = e->cars_cap; }
  int* idx = (int*)malloc(sizeof(int) * (size_t)num_cars);
  rng_distinct(r, PGTG_STREAM_CAR, num_positions, num_cars, idx); /* :837-841 */
  for (int i = 0; i < num_cars; i++) {
    int x = e->spawnable[idx[i]][0], y = e->spawnable[idx[i]][1];
    ora_car* c = &e->cars[e->n_cars++];
    c->profile = select_profile(r, i, CW0_PROFILE);
    c->route = random_route_at(r, x, y, i, CW0_ROUTE);
    c->id = e->next_car_id++;
    c->x = x; c->y = y; c->patience = 0; c->delay = 0;
  }
  free(idx);
}

static void spawn_new_car(rctx* r, ora_car* out, int slot) {
  /* _spawn_new_car (environment.py:970-1002) */
  ora_env* e = r->e;
  int x = 0, y = 0;
  if (e->n_spawners > 0) {
    int i = rng_car_index(r, slot, CW_SPAWNER, e->n_spawners);
    x = e->spawners[i][0]; y = e->spawners[i][1];
  }
  /* argument order in the reference: routes are listed first, the profile is drawn (:992),
   * then the route (:997) */
  out->profile = select_profile(r, slot, CW_PROFILE);
  out->route = random_route_at(r, x, y, slot, CW_SPAWN_ROUTE);
  out->id = e->next_car_id++;
  out->x = x; out->y = y; out->patience = 0; out->delay = 0;
}

static int should_car_move(rctx* r, ora_car* c, int slot) {
  /* _should_car_move (environment.py:678-691) */
  const pgtg_config* cfg = &r->b->cfg;
  if (c->delay > 0) { c->delay--; return 0; }
  if (rng_car_double(r, slot, CW_DELAY) < cfg->drv_reaction_delay[c->profile]) {
    c->delay = 1 + rng_car_index(r, slot, CW_IDX, 3); /* integers(1, 4) */
    return 0;
  }
  return rng_car_double(r, slot, CW_SPEED) < cfg->drv_speed_multiplier[c->profile];
}

/* returns 0 = despawn (None), 1 = (position, route) written back into the car */
static int next_car_position_and_route(rctx* r, ora_car* c, int slot) {
  /* _get_next_car_position_and_route (environment.py:881-968) */
  ora_env* e = r->e;
  const pgtg_config* cfg = &r->b->cfg;
  if (!should_car_move(r, c, slot)) { c->patience++; return 1; }
  static const int DX[4] = {0, 0, -1, 1}, DY[4] = {-1, 1, 0, 0}; /* up, down, left, right */
  for (int d = 0; d < 4; d++) {
    int px = c->x + DX[d], py = c->y + DY[d];
    if (!inside_map(e, px, py)) continue;
    const ora_lane_sq* l = lanes_at(e, px, py);
    if (!(feat_at(e, px, py) & F_LANE)) continue;
    if (l->all && l->all - 1 == d) { /* :915-928 */
      c->patience = 0;
      int route = l->route[rng_car_index(r, slot, CW_IDX, l->n)];
      c->x = px; c->y = py; c->route = route;
      return 1;
    }
    for (int i = 0; i < l->n; i++) {
      if (l->route[i] != c->route || l->dir[i] != d) continue; /* :932 */
      if (feat_at(e, px, py) & F_LIGHT) { /* :934-942 */
        int phase = light_phase(cfg, e->light_counter);
        int stop;
        if (phase == 0) stop = 0;
        else if (phase == 1) stop = rng_car_double(r, slot, CW_LIGHT) < cfg->drv_yellow_stop[c->profile];
        else stop = rng_car_double(r, slot, CW_LIGHT) >= cfg->drv_red_violation[c->profile];
        if (stop) { c->patience++; return 1; }
      }
      int blocked = 0; /* :944-948 */
      for (int k = 0; k < e->n_cars; k++) if (e->cars[k].x == px && e->cars[k].y == py) { blocked = 1; break; }
      if (blocked) { /* :950-962 */
        if (cfg->drv_min_following[c->profile] == 0 || (double)c->patience > cfg->drv_patience_threshold[c->profile]) {
          if (rng_car_double(r, slot, CW_PUSH) < cfg->drv_push_probability[c->profile]) {
            c->patience = 0; c->x = px; c->y = py; return 1;
          }
        }
        c->patience++; return 1;
      }
      c->patience = 0; c->x = px; c->y = py; return 1; /* :964-965 */
    }
  }
  c->patience++; /* :967 */
  return 0;
}

static void advance_cars(rctx* r) {
  /* environment.py:1121-1127: iterate a snapshot; removed cars are deleted in place and the
   * replacement is appended (it does not move this tick). The car being processed is
   * updated inside self.cars, so position checks of later cars see it. */
  ora_env* e = r->e;
  int n0 = e->n_cars;
  int* ids = (int*)malloc(sizeof(int) * (size_t)(n0 > 0 ? n0 : 1));
  for (int i = 0; i < n0; i++) ids[i] = e->cars[i].id;
  for (int i = 0; i < n0; i++) {
    int k = -1;
    for (int j = 0; j < e->n_cars; j++) if (e->cars[j].id == ids[i]) { k = j; break; }
    if (k < 0) continue;
    if (!next_car_position_and_route(r, &e->cars[k], i)) {
      for (int j = k; j + 1 < e->n_cars; j++) e->cars[j] = e->cars[j + 1];
      e->n_cars--;
      ora_car nc;
      spawn_new_car(r, &nc, i);
      e->cars[e->n_cars++] = nc;
    }
  }
  free(ids);
}

/* ------------------------------------------------------------------------------------------ */
/* environment.py: rule engine */

static int nearest_goal_square(const ora_env* e, int px, int py, int* gx, int* gy) {
  /* full x-major scan, first strict minimum of the Manhattan distance (:1047-1053, 1474-1480) */
  int best = -1;
  for (int tx = 0; tx < e->width; tx++)
    for (int ty = 0; ty < e->height; ty++)
      if (feat_at(e, tx, ty) & (F_SUBGOAL | F_FINAL)) {
        int d = abs(tx - px) + abs(ty - py);
        if (best < 0 || d < best) { best = d; *gx = tx; *gy = ty; }
      }
  return best >= 0;
}

static int compass_octant(double dy, double dx) {
  /* _get_subgoal_compass_directions (environment.py:1069-1088); index into
   * [N, NE, E, SE, S, SW, W, NW] */
  double angle = atan2(dy, dx);
  const double PI_8 = M_PI / 8;
  if (-PI_8 <= angle && angle < PI_8) return 2;
  if (PI_8 <= angle && angle < 3 * PI_8) return 3;
  if (3 * PI_8 <= angle && angle < 5 * PI_8) return 4;
  if (5 * PI_8 <= angle && angle < 7 * PI_8) return 5;
  if (angle >= 7 * PI_8 || angle < -7 * PI_8) return 6;
  if (-7 * PI_8 <= angle && angle < -5 * PI_8) return 7;
  if (-5 * PI_8 <= angle && angle < -3 * PI_8) return 0;
  if (-3 * PI_8 <= angle && angle < -PI_8) return 1;
  return -1;
}

static int agent_direction(const ora_batch* b, const ora_env* e) {
  /* TrafficRuleEngine.get_agent_direction (environment.py:185-206) */
  int gx, gy;
  if (nearest_goal_square(e, e->x, e->y, &gx, &gy)) {
    int dx = gx - e->x, dy = gy - e->y;
    if (!(abs(dx) <= b->cfg.window_k && abs(dy) <= b->cfg.window_k)) { /* :1061 */
      int o = compass_octant((double)dy, (double)dx);
      if (o >= 0) return o / 2; /* 0,1 s2n; 2,3 w2e; 4,5 n2s; 6,7 e2w */
    }
  }
  double speed = sqrt((double)(e->vx * e->vx + e->vy * e->vy));
  return speed < 0.1 ? PGTG_AGENT_STATIONARY : PGTG_AGENT_NEAR_GOAL;
}

static int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

static int evaluate_rule(const ora_batch* b, const ora_env* e, const pgtg_rule* rule) {
  /* TrafficRuleEngine.evaluate_rule (environment.py:226-273) */
  int tx = floordiv(e->x, TW), ty = floordiv(e->y, TH);
  if (tx < 0) tx = 0; if (tx > e->W - 1) tx = e->W - 1;
  if (ty < 0) ty = 0; if (ty > e->H - 1) ty = e->H - 1;
  if ((int)e->exits[ty * e->W + tx] != rule->tile_type) return 0;
  double speed = sqrt((double)(e->vx * e->vx + e->vy * e->vy));
  if (!(rule->vel_lo <= speed && speed <= rule->vel_hi)) return 0;
  int in_tile = 0;
  for (int k = 0; k < e->n_cars; k++)
    if (e->cars[k].x / TW == tx && e->cars[k].y / TH == ty) in_tile++;
  if (in_tile < rule->min_traffic) return 0;
  int a = agent_direction(b, e);
  int matching = 0;
  for (int k = 0; k < e->n_cars; k++)
    if (e->cars[k].x / TW == tx && e->cars[k].y / TH == ty) matching += rule->weight[a][e->cars[k].route];
  return matching >= rule->min_matching_traffic;
}

/* ------------------------------------------------------------------------------------------ */
/* environment.py: observation */

static int plane_value(const ora_batch* b, const ora_env* e, int kind, int x, int y) {
  /* encode_map_with_hot_one on the cut-out (environment.py:1387-1445, 1508-1536) */
  if (!inside_map(e, x, y)) return (b->cfg.sliding && kind == PGTG_CH_WALLS) ? 1 : 0; /* fill {"wall"} (:1384) */
  int f = feat_at(e, x, y);
  int phase = light_phase(&b->cfg, e->light_counter);
  switch (kind) {
    case PGTG_CH_WALLS: return (f & F_WALL) != 0;
    case PGTG_CH_GOALS: return (f & (F_SUBGOAL | F_FINAL)) != 0;
    case PGTG_CH_TRAFFIC:
      for (int k = 0; k < e->n_cars; k++) if (e->cars[k].x == x && e->cars[k].y == y) return 1;
      return 0;
    case PGTG_CH_ICE: return (f & F_ICE) != 0;
    case PGTG_CH_BROKEN: return (f & F_BROKEN) != 0;
    case PGTG_CH_SAND: return (f & F_SAND) != 0;
    case PGTG_CH_LIGHT_GREEN: return phase == 0 && (f & F_LIGHT);
    case PGTG_CH_LIGHT_YELLOW: return phase == 1 && (f & F_LIGHT);
    case PGTG_CH_LIGHT_RED: return phase == 2 && (f & F_LIGHT);
    case PGTG_CH_SUBGOAL: return (f & F_SUBGOAL) != 0;
    case PGTG_CH_FINAL_GOAL: return (f & F_FINAL) != 0;
    case PGTG_CH_START: return (f & F_START) != 0;
    case PGTG_CH_USED_SUBGOAL: return (f & F_USED) != 0;
    case PGTG_CH_CAR_SPAWNER: return (f & F_SPAWNER) != 0;
    default: return 0;
  }
}

static void get_observation(const ora_batch* b, const ora_env* e, int8_t* obs_map, int32_t* pos,
                            int32_t* vel, int32_t* nsd) {
  /* get_observation (environment.py:1344-1506) */
  const pgtg_config* c = &b->cfg;
  int P = b->P;
  int pix = e->x < 0 ? 0 : e->x; if (pix > e->width - 1) pix = e->width - 1;
  int piy = e->y < 0 ? 0 : e->y; if (piy > e->height - 1) piy = e->height - 1;
  int tile_x = pix / TW, tile_y = piy / TH;
  int x0, y0;
  if (!c->sliding) { x0 = tile_x * TW; y0 = tile_y * TH; }
  else { x0 = e->x - c->window_k; y0 = e->y - c->window_k; }
  for (int ch = 0; ch < b->C; ch++)
    for (int ix = 0; ix < P; ix++)
      for (int iy = 0; iy < P; iy++)
        obs_map[(ch * P + ix) * P + iy] = (int8_t)plane_value(b, e, c->channel_kind[ch], x0 + ix, y0 + iy);
  pos[0] = c->sliding ? c->window_k : pix - x0; /* :1448-1461 */
  pos[1] = c->sliding ? c->window_k : piy - y0;
  vel[0] = e->vx; vel[1] = e->vy;
  int d = -1;
  if (c->use_next_subgoal_direction) { /* :1466-1504 */
    d = e->tile_dir[tile_y * e->W + tile_x]; /* map.py:120-141 */
    if (d == -1 || c->sliding) {
      int gx, gy;
      if (nearest_goal_square(e, pix, piy, &gx, &gy)) {
        int dx = gx - pix, dy = gy - piy;
        double angle = atan2((double)(-dy), (double)dx);
        int idx = (int)fmod((angle + M_PI) / (M_PI / 4), 8.0);
        static const int remap[8] = {2, 1, 0, 7, 6, 5, 4, 3};
        d = remap[idx];
      }
    }
  }
  *nsd = d;
}

/* ------------------------------------------------------------------------------------------ */
/* environment.py: reset and step */

static inline int vis_index(const ora_env* e, int x, int y) { return (x + 1) * (e->height + 2) + (y + 1); }

static void env_reset(ora_batch* b, ora_env* e) {
  /* PGTGEnv.reset (environment.py:581-656) */
  rctx r = {b, e};
  e->episode++;
  e->elapsed = 0;
  memset(e->draw_k, 0, sizeof e->draw_k); e->blk_stream = -1; e->cblk_slot = -2;
  if (b->cfg.rng_mode == PGTG_RNG_NUMPY) for (int s = 0; s < 5; s++) np_seed_child(e, s, 5u * (e->episode - 1u) + (uint32_t)s);
  if (b->cfg.fixed_map) {
    e->W = b->fw; e->H = b->fh;
    for (int t = 0; t < b->fw * b->fh; t++) {
      e->exits[t] = b->fixed_tiles[t].exits; e->otype[t] = b->fixed_tiles[t].obstacle_type; e->omask[t] = b->fixed_tiles[t].obstacle_mask;
    }
    e->sx = b->fsx; e->sy = b->fsy; e->sdir = b->fsdir; e->gx = b->fgx; e->gy = b->fgy; e->gdir = b->fgdir;
  } else {
    generate_map(&r, e);
  }
  build_episode_map(b, e);
  e->individual_subgoal_reward = b->cfg.sum_subgoals_reward / (double)e->num_subgoals; /* :631-633 */
  int si = rng_index(&r, PGTG_STREAM_MAP, e->n_starters);                             /* :635 */
  if (e->n_starters > 0) { e->x = e->starters[si][0]; e->y = e->starters[si][1]; } else { e->x = e->y = 0; e->error |= 64; }
  e->vx = e->vy = 0;
  e->terminated = e->truncated = e->flat_tire = 0;
  e->n_cars = 0; e->next_car_id = 0; e->light_counter = 0;
  e->braking_applied = 0;
  e->outcome = 0;
  e->ep_return = 0;
  if (e->visited) {
    memset(e->visited, 0, (size_t)(e->width + 2) * (e->height + 2));
    e->visited[vis_index(e, e->x, e->y)] = 1; /* positions_path = [position] (:643) */
  }
  if (b->cfg.traffic_density > 0) create_initial_traffic(&r); /* :652-653 */
}

static double env_step(ora_batch* b, ora_env* e, int action, double* cost_out) {
  /* PGTGEnv.step (environment.py:1092-1281) */
  const pgtg_config* c = &b->cfg;
  rctx r = {b, e};
  e->elapsed++;
  memset(e->draw_k, 0, sizeof e->draw_k); e->blk_stream = -1; e->cblk_slot = -2;
  int light_total = c->light_green + c->light_yellow + c->light_red;
  e->light_counter = (e->light_counter + 1) % light_total; /* :1113-1115 */
  int ax = action / 3 - 1, ay = action % 3 - 1;            /* constants.py:6-16 */
  advance_cars(&r);                                        /* :1121-1127 */
  double reward = 0, perf = 0, cost = 0;
  int cx = e->x, cy = e->y;
  e->vx += ax; e->vy += ay;                                /* :1139 */
  /* rule_engine.apply_braking (:1145, 285-294) */
  e->braking_applied = 0;
  for (int i = 0; i < c->num_rules; i++) if (evaluate_rule(b, e, &c->rules[i])) e->braking_applied = 1;
  if (e->braking_applied) { e->vx = 0; e->vy = 0; }

  /* _decompose_velocity (:693-748) evaluated lazily, one unit sub-step at a time */
  int dx = e->vx, dy = e->vy;
  int n_sub = abs(dx) > abs(dy) ? abs(dx) : abs(dy);
  int px_prev = 0, py_prev = 0;
  for (int i = 1; i <= n_sub + 1; i++) {
    int has_part = i <= n_sub, sx = 0, sy = 0;
    if (has_part) {
      int px, py;
      int sgx = (dx > 0) - (dx < 0), sgy = (dy > 0) - (dy < 0);
      if (dx == 0) { px = 0; py = i * sgy; }
      else if (dy == 0) { px = i * sgx; py = 0; }
      else if (abs(dx) >= abs(dy)) {
        double m_y = (double)dy / (double)abs(dx);
        px = i * sgx; py = (int)floor((double)i * m_y + 0.5); /* _round (:29-30) */
      } else {
        double m_x = (double)dx / (double)abs(dy);
        py = i * sgy; px = (int)floor((double)i * m_x + 0.5);
      }
      sx = px - px_prev; sy = py - py_prev; px_prev = px; py_prev = py;
    }
    /* crash: outside map, wall, or a car on the square (:1158-1171) */
    int crash = !inside_map(e, cx, cy) || (feat_at(e, cx, cy) & F_WALL);
    if (!crash && !c->ignore_traffic_collisions)
      for (int k = 0; k < e->n_cars; k++) if (e->cars[k].x == cx && e->cars[k].y == cy) { crash = 1; break; }
    if (crash) {
      if (c->separate_reward_cost) cost += c->crash_penalty; else reward -= c->crash_penalty;
      e->terminated = 1; e->outcome = 1; break;
    }
    int f = feat_at(e, cx, cy);
    if (f & F_FINAL) { /* :1174-1180 */
      if (c->separate_reward_cost) perf += e->individual_subgoal_reward + c->final_goal_bonus;
      else reward += e->individual_subgoal_reward + c->final_goal_bonus;
      e->terminated = 1; e->outcome = 2; break;
    }
    if (f & F_SUBGOAL) { /* :1183-1188 */
      if (c->separate_reward_cost) perf += e->individual_subgoal_reward; else reward += e->individual_subgoal_reward;
      set_subgoals_to_used(e, cx, cy);
    }
    if (!has_part) continue; /* :1191-1192 */
    int nx = cx + sx, ny = cy + sy; /* red light on the NEXT square, before ice (:1195-1202) */
    if (inside_map(e, nx, ny) && (feat_at(e, nx, ny) & F_LIGHT) && light_phase(c, e->light_counter) == 2) {
      if (c->separate_reward_cost) cost += c->traffic_light_violation_penalty; else reward -= c->traffic_light_violation_penalty;
    }
    if ((f & F_ICE) && rng_double(&r, PGTG_STREAM_ICE) < c->ice_probability) { /* :1205-1213 */
      int ia = rng_index(&r, PGTG_STREAM_ICE, 9);
      sx = ia / 3 - 1; sy = ia % 3 - 1;
    }
    if ((f & F_BROKEN) && rng_double(&r, PGTG_STREAM_BROKEN) < c->street_damage_probability) e->flat_tire = 1; /* :1216-1223 */
    if ((f & F_SAND) && rng_double(&r, PGTG_STREAM_SAND) < c->sand_probability) { /* :1226-1234 */
      cx += sx; cy += sy; e->vx = 0; e->vy = 0; break;
    }
    cx += sx; cy += sy; /* :1236 */
  }
  if (e->flat_tire) { e->vx = 0; e->vy = 0; } /* :1240-1241 */
  if (c->already_visited_position_penalty != 0 && !(ax == 0 && ay == 0) && e->visited &&
      e->visited[vis_index(e, cx, cy)]) { /* :1244-1250 */
    if (c->separate_reward_cost) cost += c->already_visited_position_penalty; else reward -= c->already_visited_position_penalty;
  }
  int ox = e->x, oy = e->y;
  e->x = cx; e->y = cy; /* :1253-1255 */
  if (e->visited) e->visited[vis_index(e, cx, cy)] = 1;
  if (c->standing_still_penalty != 0 && ax == 0 && ay == 0 && ox == cx && oy == cy) { /* :1257-1263 */
    if (c->separate_reward_cost) cost += c->standing_still_penalty; else reward -= c->standing_still_penalty;
  }
  *cost_out = cost;
  return c->separate_reward_cost ? perf : reward; /* :1271-1281 */
}

/* ------------------------------------------------------------------------------------------ */
/* batch API */

static int env_alloc(ora_batch* b, ora_env* e) {
  int W = b->cfg.fixed_map ? b->fw : b->cfg.map_w, H = b->cfg.fixed_map ? b->fh : b->cfg.map_h;
  size_t sq = (size_t)W * TW * H * TH;
  e->grid = (uint16_t*)calloc(sq, sizeof(uint16_t));
  e->sq_type = (unsigned char*)calloc(sq, 1);
  e->starters = calloc(8, sizeof *e->starters);
  e->spawnable = calloc(sq, sizeof *e->spawnable);
  e->spawners = calloc(sq, sizeof *e->spawners);
  e->cars_cap = b->cfg.max_cars > 0 ? b->cfg.max_cars : 1;
  e->cars = (ora_car*)calloc((size_t)e->cars_cap + 1, sizeof(ora_car));
  e->visited = b->cfg.already_visited_position_penalty != 0 ? (unsigned char*)calloc((size_t)(W * TW + 2) * (H * TH + 2), 1) : NULL;
  return e->grid && e->sq_type && e->spawnable && e->spawners && e->cars;
}

ora_batch* ora_create(const pgtg_config* cfg) {
  ora_batch* b = (ora_batch*)calloc(1, sizeof *b);
  b->cfg = *cfg;
  b->N = cfg->num_envs;
  b->C = cfg->num_channels;
  b->P = cfg->sliding ? 2 * cfg->window_k + 1 : 9;
  b->threads = 1;
  b->envs = NULL;
  return b;
}

int ora_load_fixed_map(ora_batch* b, const pgtg_tile* tiles, int w, int h, int sx, int sy, int sdir,
                       int gx, int gy, int gdir) {
  if (w * h > PGTG_MAX_TILES) return -1;
  memcpy(b->fixed_tiles, tiles, sizeof(pgtg_tile) * (size_t)(w * h));
  b->fw = w; b->fh = h; b->fsx = sx; b->fsy = sy; b->fsdir = sdir; b->fgx = gx; b->fgy = gy; b->fgdir = gdir;
  b->have_fixed = 1;
  return 0;
}

static int ensure_envs(ora_batch* b) {
  if (b->envs) return 0;
  if (b->cfg.fixed_map && !b->have_fixed) return -1;
  b->envs = (ora_env*)calloc((size_t)b->N, sizeof(ora_env));
  for (int i = 0; i < b->N; i++) {
    if (!env_alloc(b, &b->envs[i])) return -2;
    b->envs[i].seed = b->cfg.seed + (uint64_t)b->cfg.env_id_base + (uint64_t)i;
  }
  return 0;
}

int ora_load_draws(ora_batch* b, const double* values, const uint8_t* tags, const int64_t* offsets) {
  if (ensure_envs(b)) return -1;
  b->tape_values = values; b->tape_tags = tags; /* caller keeps the arrays alive */
  for (int i = 0; i < b->N; i++) { b->envs[i].cursor = offsets[i]; b->envs[i].tape_end = offsets[i + 1]; }
  return 0;
}

void ora_set_threads(ora_batch* b, int n) { b->threads = n < 1 ? 1 : n; }

typedef struct {
  ora_batch* b; int lo, hi; const int64_t* seeds; const uint8_t* mask; const int32_t* actions;
  int8_t* obs_map; int32_t* obs_pos; int32_t* obs_vel; int32_t* obs_nsd; double* reward; double* cost;
  uint8_t* terminated; uint8_t* truncated; int32_t* step_state; uint8_t* step_flags;
  int8_t* f_map; int32_t* f_pos; int32_t* f_vel; int32_t* f_nsd;
  double stats[8];
} job;

static void* reset_job(void* p) {
  job* j = (job*)p; ora_batch* b = j->b;
  size_t os = (size_t)b->C * b->P * b->P;
  for (int i = j->lo; i < j->hi; i++) {
    if (j->mask && !j->mask[i]) continue;
    ora_env* e = &b->envs[i];
    if (j->seeds) { e->seed = (uint64_t)j->seeds[i]; e->episode = 0; }
    env_reset(b, e);
    if (j->obs_map) get_observation(b, e, j->obs_map + os * i, j->obs_pos + 2 * i, j->obs_vel + 2 * i, j->obs_nsd + i);
  }
  return NULL;
}

static void* step_job(void* p) {
  job* j = (job*)p; ora_batch* b = j->b;
  size_t os = (size_t)b->C * b->P * b->P;
  for (int i = j->lo; i < j->hi; i++) {
    ora_env* e = &b->envs[i];
    double cost = 0;
    double rew = env_step(b, e, j->actions[i], &cost);
    e->ep_return += rew;
    int trunc = b->cfg.max_episode_steps > 0 && e->elapsed >= b->cfg.max_episode_steps;
    j->reward[i] = rew; j->cost[i] = cost;
    j->terminated[i] = (uint8_t)e->terminated; j->truncated[i] = (uint8_t)trunc;
    j->step_state[4 * i] = e->x; j->step_state[4 * i + 1] = e->y; j->step_state[4 * i + 2] = e->vx; j->step_state[4 * i + 3] = e->vy;
    j->step_flags[i] = (uint8_t)((e->flat_tire ? 1 : 0) | (e->braking_applied ? 2 : 0));
    if (e->terminated || trunc) {
      /* gymnasium 0.28.1 vector semantics: terminal observation kept aside, env reset in the
       * same step, returned observation is the reset observation */
      if (j->f_map) get_observation(b, e, j->f_map + os * i, j->f_pos + 2 * i, j->f_vel + 2 * i, j->f_nsd + i);
      j->stats[0] += 1; j->stats[1] += e->ep_return; j->stats[2] += e->elapsed;
      if (e->terminated) { if (e->outcome == 2) j->stats[3] += 1; else j->stats[4] += 1; }
      else j->stats[5] += 1;
      env_reset(b, e);
    }
    get_observation(b, e, j->obs_map + os * i, j->obs_pos + 2 * i, j->obs_vel + 2 * i, j->obs_nsd + i);
  }
  return NULL;
}

static void run_jobs(ora_batch* b, job* proto, void* (*fn)(void*)) {
  int T = b->threads; if (T > b->N) T = b->N; if (T < 1) T = 1;
  job* jobs = (job*)calloc((size_t)T, sizeof(job));
  pthread_t* th = (pthread_t*)calloc((size_t)T, sizeof(pthread_t));
  for (int t = 0; t < T; t++) {
    jobs[t] = *proto;
    jobs[t].lo = (int)((int64_t)b->N * t / T); jobs[t].hi = (int)((int64_t)b->N * (t + 1) / T);
    memset(jobs[t].stats, 0, sizeof jobs[t].stats);
    if (T == 1) fn(&jobs[t]); else pthread_create(&th[t], NULL, fn, &jobs[t]);
  }
  for (int t = 0; t < T; t++) {
    if (T > 1) pthread_join(th[t], NULL);
    for (int k = 0; k < 8; k++) b->stats[k] += jobs[t].stats[k];
  }
  free(jobs); free(th);
}

/* reset(seed): seeds == NULL keeps each env's stream (a later reset() without a seed) */
int ora_reset(ora_batch* b, const int64_t* seeds, const uint8_t* mask, int8_t* obs_map, int32_t* obs_pos,
              int32_t* obs_vel, int32_t* obs_nsd) {
  if (ensure_envs(b)) return -1;
  job j; memset(&j, 0, sizeof j);
  j.b = b; j.seeds = seeds; j.mask = mask; j.obs_map = obs_map; j.obs_pos = obs_pos; j.obs_vel = obs_vel; j.obs_nsd = obs_nsd;
  run_jobs(b, &j, reset_job);
  return 0;
}

int ora_step(ora_batch* b, const int32_t* actions, int8_t* obs_map, int32_t* obs_pos, int32_t* obs_vel,
             int32_t* obs_nsd, double* reward, double* cost, uint8_t* terminated, uint8_t* truncated,
             int32_t* step_state, uint8_t* step_flags, int8_t* f_map, int32_t* f_pos, int32_t* f_vel,
             int32_t* f_nsd) {
  if (!b->envs) return -1;
  job j; memset(&j, 0, sizeof j);
  j.b = b; j.actions = actions; j.obs_map = obs_map; j.obs_pos = obs_pos; j.obs_vel = obs_vel; j.obs_nsd = obs_nsd;
  j.reward = reward; j.cost = cost; j.terminated = terminated; j.truncated = truncated;
  j.step_state = step_state; j.step_flags = step_flags; j.f_map = f_map; j.f_pos = f_pos; j.f_vel = f_vel; j.f_nsd = f_nsd;
  run_jobs(b, &j, step_job);
  return 0;
}

int ora_max_cars(const ora_batch* b) { return b->cfg.max_cars > 0 ? b->cfg.max_cars : 1; }

/* same field layout as pgtg_get_state (include/pgtg_b200.h) */
int ora_get_state(ora_batch* b, pgtg_state* s) {
  if (!b->envs) return -1;
  int T = (b->cfg.fixed_map ? b->fw * b->fh : b->cfg.map_w * b->cfg.map_h), MC = ora_max_cars(b);
  for (int i = 0; i < b->N; i++) {
    ora_env* e = &b->envs[i];
    if (s->agent) { s->agent[4 * i] = e->x; s->agent[4 * i + 1] = e->y; s->agent[4 * i + 2] = e->vx; s->agent[4 * i + 3] = e->vy; }
    if (s->flat_tire) s->flat_tire[i] = (uint8_t)e->flat_tire;
    if (s->light_counter) s->light_counter[i] = e->light_counter;
    if (s->elapsed) s->elapsed[i] = e->elapsed;
    if (s->num_cars) s->num_cars[i] = e->n_cars;
    if (s->cars) {
      memset(s->cars + (size_t)i * MC * 7, 0, sizeof(int32_t) * (size_t)MC * 7);
      for (int k = 0; k < e->n_cars && k < MC; k++) {
        int32_t* o = s->cars + ((size_t)i * MC + k) * 7;
        ora_car* c = &e->cars[k];
        o[0] = c->id; o[1] = c->x; o[2] = c->y; o[3] = c->route; o[4] = c->profile; o[5] = c->patience; o[6] = c->delay;
      }
    }
    if (s->tiles) for (int t = 0; t < T; t++) {
      int sg = e->tile_dir[t] >= 0 ? e->tile_dir[t] + 1 : 0;
      s->tiles[(size_t)i * T + t] = (uint16_t)(e->exits[t] | e->otype[t] << 4 | e->omask[t] << 7 | sg << 11);
    }
    if (s->plan) { int32_t* o = s->plan + 8 * i; o[0] = e->sx; o[1] = e->sy; o[2] = e->sdir; o[3] = e->gx; o[4] = e->gy; o[5] = e->gdir; o[6] = e->num_subgoals; o[7] = 0; }
    if (s->used) for (int t = 0; t < T; t++) {
      /* a tile's subgoal is consumed iff its exit line carries "used subgoal" */
      int tx = t % e->W, ty = t / e->W, u = 0;
      for (int x = 0; x < 9; x++) for (int y = 0; y < 9; y++) if (feat_at(e, tx * 9 + x, ty * 9 + y) & F_USED) u = 1;
      s->used[(size_t)i * T + t] = (uint8_t)u;
    }
    if (s->draw_cursor) s->draw_cursor[i] = e->cursor;
    if (s->error) s->error[i] = e->error;
  }
  return 0;
}

/* set_to_state (environment.py:1301-1342): position, velocity, flat_tire, cars (id, x, y, route,
 * profile; patience and delay restart at 0); nothing else (quirk A.3-10) */
int ora_set_state(ora_batch* b, const pgtg_state* s) {
  if (!b->envs) return -1;
  int MC = ora_max_cars(b);
  for (int i = 0; i < b->N; i++) {
    ora_env* e = &b->envs[i];
    if (s->agent) { e->x = s->agent[4 * i]; e->y = s->agent[4 * i + 1]; e->vx = s->agent[4 * i + 2]; e->vy = s->agent[4 * i + 3]; }
    if (s->flat_tire) e->flat_tire = s->flat_tire[i];
    if (s->cars && s->num_cars) {
      e->n_cars = 0;
      for (int k = 0; k < s->num_cars[i] && k < e->cars_cap; k++) {
        const int32_t* o = s->cars + ((size_t)i * MC + k) * 7;
        ora_car* c = &e->cars[e->n_cars++];
        c->id = o[0]; c->x = o[1]; c->y = o[2]; c->route = o[3]; c->profile = o[4]; c->patience = 0; c->delay = 0;
      }
      if (e->n_cars > 0) e->next_car_id = e->cars[e->n_cars - 1].id + 1; /* :1340 */
    }
  }
  return 0;
}

int ora_observe(ora_batch* b, int8_t* obs_map, int32_t* obs_pos, int32_t* obs_vel, int32_t* obs_nsd) {
  if (!b->envs) return -1;
  size_t os = (size_t)b->C * b->P * b->P;
  for (int i = 0; i < b->N; i++) get_observation(b, &b->envs[i], obs_map + os * i, obs_pos + 2 * i, obs_vel + 2 * i, obs_nsd + i);
  return 0;
}

/* agent_direction string id per env (get_info()['traffic_rules']['agent_direction'], :1564) */
int ora_agent_direction(ora_batch* b, int32_t* out) {
  if (!b->envs) return -1;
  for (int i = 0; i < b->N; i++) out[i] = agent_direction(b, &b->envs[i]);
  return 0;
}

void ora_stats(ora_batch* b, double* out8, int reset_after) {
  memcpy(out8, b->stats, sizeof b->stats);
  if (reset_after) memset(b->stats, 0, sizeof b->stats);
}

void ora_destroy(ora_batch* b) {
  if (!b) return;
  if (b->envs) for (int i = 0; i < b->N; i++) {
    ora_env* e = &b->envs[i];
    free(e->grid); free(e->sq_type); free(e->starters); free(e->spawnable); free(e->spawners); free(e->cars); free(e->visited);
  }
  free(b->envs); free(b);
}

/* _decompose_velocity known answers (tests/test_environment.py:1127-1153): unit sub-steps */
int ora_decompose_velocity(int dx, int dy, int* out_xy) {
  int n_sub = abs(dx) > abs(dy) ? abs(dx) : abs(dy), pxp = 0, pyp = 0;
  for (int i = 1; i <= n_sub; i++) {
    int px, py, sgx = (dx > 0) - (dx < 0), sgy = (dy > 0) - (dy < 0);
    if (dx == 0) { px = 0; py = i * sgy; }
    else if (dy == 0) { px = i * sgx; py = 0; }
    else if (abs(dx) >= abs(dy)) { double m = (double)dy / (double)abs(dx); px = i * sgx; py = (int)floor((double)i * m + 0.5); }
    else { double m = (double)dx / (double)abs(dy); py = i * sgy; px = (int)floor((double)i * m + 0.5); }
    out_xy[2 * (i - 1)] = px - pxp; out_xy[2 * (i - 1) + 1] = py - pyp; pxp = px; pyp = py;
  }
  return n_sub;
}
