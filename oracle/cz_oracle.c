/* cz_oracle.c — CPU oracle in plain C: a restatement of CookingZoo's per-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).  The product
 * library never links or calls this file.  It exists beside oracle/cz_oracle.py because the Python
 * oracle (~10 k env-steps/s) cannot cover BASELINE's full batch sizes exhaustively; this one can.
 * Pinned the same way: tests/test_c_oracle.py replays every golden trace recorded from the
 * unmodified reference (tests/golden/*.npz) through it, bit for bit.
 *
 * Deliberately NOT shaped like the CUDA kernels: objects are structs, containers hold ordered
 * content lists, queries walk per-type lists in world insertion order — i.e. the reference's own
 * data model (paths relative to /root/reference/cooking_zoo/):
 *   czo_step            environment/cooking_env.py:243-269 (accumulated_step)
 *     agent_actions_3   cooking_world/cooking_action_util/action_scheme3.py:4-43
 *     agent_actions_1   cooking_world/cooking_action_util/action_scheme1.py:4-40
 *     checked_actions   cooking_world/cooking_world.py:192-221
 *     primary / merge   cooking_world/cooking_world.py:114-136, 243-261
 *     pickup_special    cooking_world/cooking_world.py:138-154
 *     execute           cooking_world/cooking_world.py:156-170, world_objects.py:250-269, 356-360, 738-745
 *     progress_world    cooking_world/cooking_world.py:77-88, world_objects.py:321-335
 *     linked            cooking_world/cooking_world.py:90-92
 *     agent_spawn       cooking_world/cooking_world.py:267-290, engine/parsing.py:154-167
 *     rewards           environment/cooking_env.py:290-350
 *   update_recipe       cooking_book/recipe.py:77-104
 *   czo_observe         environment/cooking_env.py:352-373
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { T_FLOOR, T_COUNTER, T_DELIVER, T_SWITCH, T_BLOCK, T_CUTBOARD, T_BLENDER, T_PLATE, T_ONION, T_TOMATO,
       T_LETTUCE, T_CARROT, T_CUCUMBER, T_BANANA, T_APPLE, T_WATERMELON, T_BREAD, T_AGENT, T_COUNT };
#define IS_STATIC(t) ((t) <= T_BLENDER)
#define IS_DYNAMIC(t) ((t) >= T_PLATE && (t) <= T_BREAD)
#define IS_FOOD(t) ((t) >= T_ONION && (t) <= T_BREAD)
#define IS_BLENDFOOD(t) ((t) == T_CARROT || (t) == T_BANANA)
static const int FV_LEN[T_COUNT] = {0, 3, 3, 4, 4, 3, 3, 3, 5, 5, 5, 6, 5, 6, 5, 5, 5, 7};

#define MAX_OBJ 96      /* <= 64 statics (8x8 level) + 32 dynamic slots */
#define MAX_CONTENT 32  /* a plate holds at most every food of the level */
#define MAX_AGENTS 4
#define MAX_RECIPES 8
#define MAX_NODES 8
#define MAX_META 32

typedef struct {
  int8_t type, x, y, exists;
  /* static */
  int8_t walkable, ready, toggle, active, pressed;
  /* dynamic */
  int8_t chopped, blend /* 0 fresh 1 in progress 2 mashed */, progress, free_flag;
  uint8_t content[MAX_CONTENT], n_content; /* indices into objs: Counter/Cutboard/... hold dynamics, Plate holds foods */
} Obj;

typedef struct { int x, y, orientation, holding /* obj index or -1 */, active, grace; } Agent;
typedef struct { int type, cond, kids, marked; uint64_t hits; } Node; /* hits: cells of satisfying objects */

struct CzoInit;
typedef struct CzoEnv {
  int W, H, A, R, max_steps, end_all, scheme, grace_period;
  double r_node, r_recipe, r_penalty, mtp, respawn, despawn;
  int n_meta, meta_kind[MAX_META], meta_type[MAX_META], meta_count[MAX_META];
  Obj objs[MAX_OBJ];
  struct CzoInit* init; /* cold copy of the initial world (czo_reset), kept out of the stepping footprint */
  int n_objs;
  int type_order[T_COUNT], n_types; /* world_objects insertion order */
  uint8_t by_type[T_COUNT][MAX_OBJ];
  int n_by_type[T_COUNT];
  uint8_t static_at[8][8];
  Agent agents[MAX_AGENTS];
  int spawn_nx[MAX_AGENTS], spawn_ny[MAX_AGENTS], spawn_x[MAX_AGENTS][8], spawn_y[MAX_AGENTS][8];
  int status_changed[MAX_AGENTS], relevant[MAX_AGENTS], n_live, t;
  Node nodes[MAX_RECIPES][MAX_NODES];
  int n_nodes[MAX_RECIPES];
  uint32_t error;
  uint64_t seed, env, episode;
  uint32_t draw;
} CzoEnv;

typedef struct CzoInit { Obj objs[MAX_OBJ]; Agent agents[MAX_AGENTS]; int n_objs, n_by_type[T_COUNT]; } CzoInit;

/* the shared counter-based stream (include/cz_b200.h: cz_spawn_uniform) */
static double spawn_uniform(uint64_t seed, uint64_t env, uint64_t episode, uint64_t t, uint64_t c) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (env + 1) + 0xD1B54A32D192ED03ull * (episode + 1) +
               0x8CB92BA72F3D8DD7ull * (t + 1) + 0xF1357AEA2E62A9C5ull * (c + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
static double draw_u(CzoEnv* e) { return spawn_uniform(e->seed, e->env, e->episode, (uint64_t)e->t, e->draw++); }

static int obj_done(const Obj* o) { return o->chopped || (IS_BLENDFOOD(o->type) && o->blend == 2); }

static void refresh_free(CzoEnv* e, Obj* holder) {
  for (int k = 0; k < holder->n_content; ++k) e->objs[holder->content[k]].free_flag = 0;
  if (holder->n_content) e->objs[holder->content[holder->n_content - 1]].free_flag = 1;
}
static void content_remove(Obj* holder, int idx) {
  int k = 0;
  while (k < holder->n_content && holder->content[k] != idx) ++k;
  for (; k + 1 < holder->n_content; ++k) holder->content[k] = holder->content[k + 1];
  holder->n_content--;
}
static int content_has(const Obj* holder, int idx) {
  for (int k = 0; k < holder->n_content; ++k) if (holder->content[k] == idx) return 1;
  return 0;
}
static void move_obj(CzoEnv* e, int idx, int x, int y) {
  Obj* o = &e->objs[idx];
  o->x = (int8_t)x; o->y = (int8_t)y;
  for (int k = 0; k < o->n_content; ++k) { e->objs[o->content[k]].x = (int8_t)x; e->objs[o->content[k]].y = (int8_t)y; }
}

static int add_object(CzoEnv* e, int type, int x, int y) {
  int idx = e->n_objs++;
  Obj* o = &e->objs[idx];
  memset(o, 0, sizeof(*o));
  o->type = (int8_t)type; o->x = (int8_t)x; o->y = (int8_t)y; o->exists = 1;
  o->walkable = (type == T_FLOOR || type == T_SWITCH);
  o->progress = 1; o->free_flag = 1;
  if (e->n_by_type[type] == 0) {
    int known = 0;
    for (int k = 0; k < e->n_types; ++k) known |= e->type_order[k] == type;
    if (!known) e->type_order[e->n_types++] = type;
  }
  e->by_type[type][e->n_by_type[type]++] = (uint8_t)idx;
  if (IS_STATIC(type)) e->static_at[y][x] = (uint8_t)idx;
  return idx;
}

/* get_objects_at(location, DynamicObject): type insertion order, then list order */
static int scan(const CzoEnv* e, int x, int y, int* out) {
  int n = 0;
  for (int k = 0; k < e->n_types; ++k) {
    int ty = e->type_order[k];
    if (!IS_DYNAMIC(ty)) continue;
    for (int j = 0; j < e->n_by_type[ty]; ++j) {
      const Obj* o = &e->objs[e->by_type[ty][j]];
      if (o->x == x && o->y == y) out[n++] = e->by_type[ty][j];
    }
  }
  return n;
}
static int agent_on(const CzoEnv* e, int x, int y) {
  for (int i = 0; i < e->A; ++i) if (e->agents[i].x == x && e->agents[i].y == y) return 1;
  return 0;
}
static void target(const Agent* a, int action, int* tx, int* ty) {
  *tx = a->x + (action == 2) - (action == 1);
  *ty = a->y + (action == 3) - (action == 4);
}
static int in_grid(const CzoEnv* e, int x, int y) { return x >= 0 && y >= 0 && x < e->W && y < e->H; }
static int walkable(const CzoEnv* e, int x, int y) { return e->objs[e->static_at[y][x]].walkable; }

/* ---- recipes ------------------------------------------------------------------------- */
static void update_recipe(CzoEnv* e, int r) {
  Node* nd = e->nodes[r];
  for (int k = e->n_nodes[r] - 1; k >= 0; --k) {
    nd[k].marked = 0; nd[k].hits = 0;
    int ok = 1;
    for (int j = 0; j < e->n_nodes[r]; ++j) if ((nd[k].kids >> j & 1) && !nd[j].marked) ok = 0;
    if (!ok) continue;
    int ty = nd[k].type;
    for (int j = 0; j < e->n_by_type[ty]; ++j) {
      const Obj* o = &e->objs[e->by_type[ty][j]];
      if (nd[k].cond == 1 && !o->chopped) continue;
      if (nd[k].cond == 2 && o->blend != 2) continue;
      uint64_t cell = 1ull << (o->y * 8 + o->x);
      int all = 1;
      for (int c = 0; c < e->n_nodes[r]; ++c) if ((nd[k].kids >> c & 1) && !(nd[c].hits & cell)) all = 0;
      if (all) { nd[k].hits |= cell; nd[k].marked = 1; }
    }
  }
}

/* ---- interactions -------------------------------------------------------------------- */
static int st_accepts(const Obj* st, const Obj* o) {
  switch (st->type) {
    case T_COUNTER: case T_DELIVER: return st->n_content < 1;
    case T_CUTBOARD: return IS_FOOD(o->type) && st->n_content < 1 && !o->chopped;
    case T_BLENDER: return IS_BLENDFOOD(o->type) && !st->toggle && st->n_content + 1 <= 1 && o->blend == 0;
    default: return 0;
  }
}
static int plate_accepts(const Obj* p, const Obj* o) { return IS_FOOD(o->type) && obj_done(o) && p->n_content < 64 && p->n_content < MAX_CONTENT; }
static void add_content(CzoEnv* e, int holder, int idx) {
  Obj* h = &e->objs[holder];
  if (h->type == T_CUTBOARD || h->type == T_BLENDER) h->ready = 1;
  h->content[h->n_content++] = (uint8_t)idx;
  refresh_free(e, h);
}

static void execute(CzoEnv* e, int x, int y) {
  if (agent_on(e, x, y)) return;
  Obj* st = &e->objs[e->static_at[y][x]];
  if (st->type == T_CUTBOARD) {
    if (!st->ready) return;
    int n = st->n_content;
    for (int k = 0; k < n; ++k) {
      Obj* o = &e->objs[st->content[k]];
      if (!IS_FOOD(o->type)) return;
      if (o->chopped) continue;
      o->chopped = 1;
      st->ready = 0;
      if (o->type == T_BREAD) { /* Bread.chop spawns a chopped twin appended to content and world */
        int twin = add_object(e, T_BREAD, o->x, o->y);
        e->objs[twin].chopped = 1;
        /* more Breads than the meta file has slots: the reference's observation silently grows (cooking_env.py:371) */
        for (int m = 0; m < e->n_meta; ++m)
          if (e->meta_type[m] == T_BREAD && e->n_by_type[T_BREAD] > e->meta_count[m]) e->error |= 32u;
        st = &e->objs[e->static_at[y][x]];
        st->content[st->n_content++] = (uint8_t)twin;
      }
      return;
    }
    e->error |= 1u;
  } else if (st->type == T_BLENDER) {
    if (st->ready) st->toggle = !st->toggle;
  }
}

static void primary(CzoEnv* e, Agent* ag, int x, int y) {
  if (agent_on(e, x, y)) return;
  int sidx = e->static_at[y][x];
  Obj* st = &e->objs[sidx];
  int dyn[MAX_OBJ];
  int n = scan(e, x, y, dyn);
  if (ag->holding < 0) {
    if (!n) return;
    int rel = 1;
    if (st->type == T_DELIVER) rel = 0;
    else if (st->type == T_CUTBOARD) { if (st->n_content == 1) st->ready = 0; }
    else if (st->type == T_BLENDER) { if (st->toggle) rel = 0; else if (st->n_content - 1 == 0) st->ready = 0; }
    if (!rel) return;
    int grab = dyn[n - 1];
    for (int k = 0; k < n; ++k) if (e->objs[dyn[k]].free_flag) { grab = dyn[k]; break; }
    if (content_has(st, grab)) {
      ag->holding = grab;
      move_obj(e, grab, ag->x, ag->y);
      content_remove(st, grab);
    }
    return;
  }
  int h = ag->holding, plates[MAX_OBJ], np = 0;
  for (int k = 0; k < n; ++k) if (e->objs[dyn[k]].type == T_PLATE) plates[np++] = dyn[k];
  if (np == 1) {
    if (plate_accepts(&e->objs[plates[0]], &e->objs[h])) {
      add_content(e, plates[0], h);
      move_obj(e, h, x, y);
      ag->holding = -1;
    }
  } else if (e->objs[h].type == T_PLATE && n) {
    int p = dyn[n - 1];
    if (plate_accepts(&e->objs[h], &e->objs[p])) {
      add_content(e, h, p);
      move_obj(e, p, ag->x, ag->y);
      if (content_has(st, p)) content_remove(st, p); else e->error |= 2u;
    }
  } else if (st_accepts(st, &e->objs[h])) {
    add_content(e, sidx, h);
    move_obj(e, h, x, y);
    ag->holding = -1;
  }
}

static void pickup_special(CzoEnv* e, Agent* ag, int x, int y) {
  if (agent_on(e, x, y)) return;
  int dyn[MAX_OBJ];
  int n = scan(e, x, y, dyn);
  if (ag->holding >= 0 || !n) return;
  int plate = -1, np = 0;
  for (int k = 0; k < n; ++k) if (e->objs[dyn[k]].type == T_PLATE) { plate = dyn[k]; ++np; }
  if (np != 1 || e->objs[plate].n_content == 0) return;
  Obj* p = &e->objs[plate];
  int o = p->content[--p->n_content];
  ag->holding = o;
  move_obj(e, o, ag->x, ag->y);
}

static void walk_to(CzoEnv* e, Agent* ag, int tx, int ty) {
  ag->x = tx; ag->y = ty;
  if (ag->holding >= 0) move_obj(e, ag->holding, tx, ty);
  Obj* st = &e->objs[e->static_at[ty][tx]];
  if (st->type == T_SWITCH) { st->active = !st->active; st->pressed = 1; }
}

/* check_inbounds + check_collisions */
static void checked_actions(CzoEnv* e, const int* idx, int n, int* acts) {
  int ex[MAX_AGENTS], ey[MAX_AGENTS], w[MAX_AGENTS], fin[MAX_AGENTS];
  for (int k = 0; k < n; ++k) {
    int a = acts[k], tx, ty;
    if (a == 0 || a == 5) continue;
    target(&e->agents[idx[k]], a, &tx, &ty);
    if (tx > e->W - 1 || tx < 0 || ty > e->H - 1 || ty < 0) acts[k] = 0;
  }
  for (int k = 0; k < n; ++k) {
    const Agent* ag = &e->agents[idx[k]];
    int tx, ty;
    target(ag, acts[k], &tx, &ty);
    w[k] = walkable(e, tx, ty);
    ex[k] = w[k] ? tx : ag->x; ey[k] = w[k] ? ty : ag->y;
  }
  for (int k = 0; k < n; ++k) {
    int hit = 0;
    for (int j = 0; j < n; ++j) if (j != k && ex[j] == ex[k] && ey[j] == ey[k]) hit = 1;
    fin[k] = (hit && w[k]) ? 0 : acts[k];
  }
  memcpy(acts, fin, sizeof(int) * n);
}

static void agent_actions_3(CzoEnv* e, const int* idx, int n, int* acts) {
  int fx[MAX_AGENTS], fy[MAX_AGENTS];
  for (int k = 0; k < n; ++k) {
    Agent* ag = &e->agents[idx[k]];
    if (acts[k] >= 1 && acts[k] <= 4) { target(ag, acts[k], &fx[k], &fy[k]); ag->orientation = acts[k]; }
    else { fx[k] = ag->x; fy[k] = ag->y; }
  }
  checked_actions(e, idx, n, acts);
  for (int k = 0; k < n; ++k) {
    Agent* ag = &e->agents[idx[k]];
    int tx, ty;
    target(ag, acts[k], &tx, &ty);
    if (walkable(e, tx, ty)) { walk_to(e, ag, tx, ty); continue; }
    if (acts[k] == 0) continue;
    Obj* st = &e->objs[e->static_at[fy[k]][fx[k]]];
    int dyn[MAX_OBJ], nd = scan(e, fx[k], fy[k], dyn), unfinished = 0;
    for (int j = 0; j < nd; ++j) if (e->objs[dyn[j]].type != T_PLATE && !obj_done(&e->objs[dyn[j]])) unfinished = 1;
    if ((st->type == T_CUTBOARD || st->type == T_BLENDER) && unfinished) execute(e, fx[k], fy[k]);
    else primary(e, ag, fx[k], fy[k]);
  }
}

static void agent_actions_1(CzoEnv* e, const int* idx, int n, int* acts) {
  for (int k = 0; k < n; ++k) if (acts[k] >= 1 && acts[k] <= 4) e->agents[idx[k]].orientation = acts[k];
  checked_actions(e, idx, n, acts);
  for (int k = 0; k < n; ++k) {
    Agent* ag = &e->agents[idx[k]];
    int a = acts[k], tx, ty;
    if (a >= 1 && a <= 4) {
      target(ag, a, &tx, &ty);
      if (walkable(e, tx, ty)) walk_to(e, ag, tx, ty);
    } else if (a >= 5 && a <= 7) {
      target(ag, ag->orientation, &tx, &ty);
      /* off the grid: get_objects_at(...)[0] raises IndexError (cooking_world.py:119, :160); pick-up-special (:138-154)
       * never looks the static object up and simply finds nothing */
      if (!in_grid(e, tx, ty)) { if (a != 6) e->error |= 64u; continue; }
      if (a == 5) primary(e, ag, tx, ty);
      else if (a == 6) pickup_special(e, ag, tx, ty);
      else execute(e, tx, ty);
    }
  }
}

static void progress_world(CzoEnv* e) {
  for (int j = 0; j < e->n_by_type[T_BLENDER]; ++j) {
    Obj* b = &e->objs[e->by_type[T_BLENDER][j]];
    if (!b->n_content || !b->toggle) continue;
    int all = 1;
    for (int k = 0; k < b->n_content; ++k) {
      Obj* c = &e->objs[b->content[k]];
      if (!obj_done(c) && (c->blend == 0 || c->blend == 1)) { c->progress -= 1; c->blend = (int8_t)(c->progress > 0 ? 1 : 2); }
      if (c->blend != 2) all = 0;
    }
    if (all) {
      b->toggle = 0; b->ready = 0;
      for (int k = 0; k < b->n_content; ++k) e->objs[b->content[k]].progress = 1;
    }
  }
  for (int i = 0; i < e->n_objs; ++i) if (e->objs[i].exists && e->objs[i].n_content) refresh_free(e, &e->objs[i]);
}

static void linked(CzoEnv* e) {
  for (int j = 0; j < e->n_by_type[T_SWITCH]; ++j) {
    Obj* s = &e->objs[e->by_type[T_SWITCH][j]];
    if (s->pressed) {
      if (e->n_by_type[T_SWITCH] > 1) e->error |= 4u;
      for (int k = 0; k < e->n_by_type[T_BLOCK]; ++k) { Obj* b = &e->objs[e->by_type[T_BLOCK][k]]; b->walkable = !b->walkable; }
    }
    s->pressed = 0;
  }
}

static void agent_spawn(CzoEnv* e) {
  for (int i = 0; i < e->A; ++i) {
    Agent* ag = &e->agents[i];
    if (ag->grace > 0) { ag->grace--; continue; }
    if (e->respawn <= 0.0 && e->despawn <= 0.0) continue; /* the reference's draws change nothing at rate 0 */
    int n_active = 0;
    for (int j = 0; j < e->A; ++j) n_active += e->agents[j].active;
    if (n_active > 1 && ag->active) {
      if (draw_u(e) < e->despawn && ag->holding < 0) { ag->active = 0; e->status_changed[i] = 1; }
    } else if (!ag->active) {
      if (draw_u(e) < e->respawn) {
        ag->active = 1; e->status_changed[i] = 1; ag->grace = e->grace_period;
        int found = 0;
        for (int tries = 0; tries < 1002 && !found; ++tries) {
          int kx = (int)(draw_u(e) * e->spawn_nx[i]); if (kx > e->spawn_nx[i] - 1) kx = e->spawn_nx[i] - 1;
          int ky = (int)(draw_u(e) * e->spawn_ny[i]); if (ky > e->spawn_ny[i] - 1) ky = e->spawn_ny[i] - 1;
          int x = e->spawn_x[i][kx], y = e->spawn_y[i][ky];
          if (in_grid(e, x, y) && e->objs[e->static_at[y][x]].type == T_FLOOR && !agent_on(e, x, y)) { ag->x = x; ag->y = y; found = 1; }
        }
        if (!found) e->error |= 8u;
      }
    }
  }
}

/* ---- public API ---------------------------------------------------------------------- */
CzoEnv* czo_create(const int32_t* cfg, const double* rw, const int32_t* meta, int n_meta, const int32_t* world, int n_world,
                   const int32_t* agents, const int32_t* spawn, const int32_t* recipes) {
  CzoEnv* e = (CzoEnv*)calloc(1, sizeof(CzoEnv));
  if (!e) return 0;
  e->W = cfg[0]; e->H = cfg[1]; e->A = cfg[2]; e->R = cfg[3]; e->max_steps = cfg[4]; e->end_all = cfg[5];
  e->scheme = cfg[6]; e->grace_period = cfg[7];
  e->r_node = rw[0]; e->r_recipe = rw[1]; e->r_penalty = rw[2]; e->mtp = rw[3]; e->respawn = rw[4]; e->despawn = rw[5];
  e->n_meta = n_meta;
  for (int k = 0; k < n_meta; ++k) { e->meta_kind[k] = meta[3 * k]; e->meta_type[k] = meta[3 * k + 1]; e->meta_count[k] = meta[3 * k + 2]; }
  for (int k = 0; k < n_world; ++k) add_object(e, world[3 * k], world[3 * k + 1], world[3 * k + 2]);
  /* dynamic objects start as the content of the Counter under them (parsing.py:107-108) */
  for (int i = 0; i < e->n_objs; ++i)
    if (IS_DYNAMIC(e->objs[i].type)) {
      Obj* holder = &e->objs[e->static_at[e->objs[i].y][e->objs[i].x]];
      holder->content[holder->n_content++] = (uint8_t)i;
      refresh_free(e, holder);
    }
  for (int i = 0; i < e->A; ++i) {
    Agent* ag = &e->agents[i];
    ag->x = agents[2 * i]; ag->y = agents[2 * i + 1]; ag->orientation = 1; ag->holding = -1; ag->active = 1; ag->grace = e->grace_period;
    const int32_t* sp = spawn + i * 18;
    e->spawn_nx[i] = sp[0]; e->spawn_ny[i] = sp[9];
    for (int k = 0; k < 8; ++k) { e->spawn_x[i][k] = sp[1 + k]; e->spawn_y[i][k] = sp[10 + k]; }
  }
  for (int r = 0; r < e->R; ++r) {
    const int32_t* rc = recipes + r * (1 + 3 * MAX_NODES);
    e->n_nodes[r] = rc[0];
    for (int k = 0; k < rc[0]; ++k) { e->nodes[r][k].type = rc[1 + 3 * k]; e->nodes[r][k].cond = rc[2 + 3 * k]; e->nodes[r][k].kids = rc[3 + 3 * k]; }
  }
  e->init = (CzoInit*)malloc(sizeof(CzoInit));
  if (!e->init) { free(e); return 0; }
  memcpy(e->init->objs, e->objs, sizeof(e->objs));
  memcpy(e->init->agents, e->agents, sizeof(e->agents));
  e->init->n_objs = e->n_objs;
  memcpy(e->init->n_by_type, e->n_by_type, sizeof(e->n_by_type));
  for (int r = 0; r < e->R; ++r) update_recipe(e, r);
  e->n_live = e->A;
  for (int i = 0; i < e->A; ++i) e->relevant[i] = 1;
  return e;
}
void czo_destroy(CzoEnv* e) { if (e) { free(e->init); free(e); } }

void czo_reset(CzoEnv* e) {
  memcpy(e->objs, e->init->objs, sizeof(e->objs));
  memcpy(e->agents, e->init->agents, sizeof(e->agents));
  e->n_objs = e->init->n_objs;
  memcpy(e->n_by_type, e->init->n_by_type, sizeof(e->n_by_type));
  e->t = 0; e->error = 0; e->n_live = e->A;
  for (int r = 0; r < e->R; ++r) update_recipe(e, r);
  for (int i = 0; i < e->A; ++i) { e->relevant[i] = 1; e->status_changed[i] = 0; }
}

void czo_set_stream(CzoEnv* e, uint64_t seed, uint64_t env, uint64_t episode) { e->seed = seed; e->env = env; e->episode = episode; }

void czo_step(CzoEnv* e, const uint8_t* actions, double* reward, uint8_t* term, uint8_t* trunc, uint8_t* rel_out) {
  const int A = e->A;
  e->t += 1;
  e->draw = 0;
  int idx[MAX_AGENTS], acts[MAX_AGENTS], n = 0;
  for (int i = 0; i < A; ++i) { e->status_changed[i] = 0; if (e->agents[i].active) { idx[n] = i; acts[n++] = actions[i]; } }
  if (e->scheme == 1) agent_actions_1(e, idx, n, acts); else agent_actions_3(e, idx, n, acts);
  progress_world(e);
  linked(e);
  agent_spawn(e);
  int relevant[MAX_AGENTS], n_rel = 0;
  for (int i = 0; i < A; ++i) { relevant[i] = e->agents[i].active || e->status_changed[i]; n_rel += relevant[i]; }
  /* compute_truncated */
  int tr[MAX_AGENTS] = {0};
  if (e->t >= e->max_steps) {
    if (e->n_live < A) e->error |= 16u;
    for (int k = 0; k < n_rel; ++k) tr[k] = 1;
    for (int i = 0; i < A; ++i) { e->agents[i].active = 0; e->status_changed[i] = relevant[i]; }
  }
  for (int i = 0, k = 0; i < A; ++i) {
    if (!relevant[i]) continue;
    if (e->status_changed[i] && !e->agents[i].active) tr[k] = 1;
    ++k;
  }
  /* compute_rewards: left-to-right adds starting from the int 0 */
  double rr[MAX_RECIPES];
  int all_done = 1, any_done = 0;
  for (int r = 0; r < e->R; ++r) {
    int before = 0, after = 0, was = e->nodes[r][0].marked;
    for (int k = 0; k < e->n_nodes[r]; ++k) before += !e->nodes[r][k].marked;
    update_recipe(e, r);
    for (int k = 0; k < e->n_nodes[r]; ++k) after += !e->nodes[r][k].marked;
    int now = e->nodes[r][0].marked;
    volatile double v = 0.0;
    v = v + (double)(before - after) * e->r_node;
    v = v + ((now && !was) ? e->r_recipe : 0.0);
    v = v + ((!now && was) ? e->r_penalty : 0.0);
    v = v + e->mtp;
    rr[r] = v;
    all_done &= now; any_done |= now;
  }
  int done = e->end_all ? all_done : any_done;
  e->n_live = 0;
  for (int i = 0, k = 0; i < A; ++i) {
    int rel = e->agents[i].active || e->status_changed[i];
    reward[i] = 0.0; term[i] = 0; trunc[i] = 0; rel_out[i] = (uint8_t)rel;
    e->relevant[i] = rel;
    if (!rel) continue;
    reward[i] = k < e->R ? rr[k] : 0.0;
    term[i] = (uint8_t)done; trunc[i] = (uint8_t)tr[k];
    ++k; e->n_live++;
  }
}

int czo_obs_len(const CzoEnv* e) {
  int n = 0;
  for (int k = 0; k < e->n_meta; ++k) n += FV_LEN[e->meta_type[k]] * e->meta_count[k];
  return n;
}

void czo_observe(const CzoEnv* e, int agent, double* out) {
  const Agent* me = &e->agents[agent];
  const double W = e->W, H = e->H;
  int p = 0;
  for (int m = 0; m < e->n_meta; ++m) {
    int ty = e->meta_type[m], n = 0, len = FV_LEN[ty];
    if (ty == T_AGENT) {
      for (int i = 0; i < e->A; ++i, ++n) {
        const Agent* ag = &e->agents[i];
        out[p++] = i == agent ? ag->x / W : (ag->x - me->x) / W;
        out[p++] = i == agent ? ag->y / H : (ag->y - me->y) / H;
        for (int o = 1; o <= 4; ++o) out[p++] = ag->orientation == o;
        out[p++] = 1.0;
      }
    } else {
      for (int j = 0; j < e->n_by_type[ty] && j < e->meta_count[m]; ++j, ++n) {  /* beyond the slots: flagged (error 32) at creation */
        const Obj* o = &e->objs[e->by_type[ty][j]];
        if (!len) continue;
        out[p++] = (o->x - me->x) / W;
        out[p++] = (o->y - me->y) / H;
        if (ty == T_SWITCH) out[p++] = o->active;
        else if (ty == T_BLOCK) out[p++] = o->walkable;
        else if (IS_FOOD(ty)) {
          out[p++] = !obj_done(o);
          out[p++] = o->chopped;
          if (IS_BLENDFOOD(ty)) out[p++] = o->blend == 2;
        }
        out[p++] = 1.0;
      }
    }
    for (int k = 0; k < (e->meta_count[m] - n) * len; ++k) out[p++] = 0.0;
  }
}

/* canonical arrays, same convention as oracle/ref_dump.dump_state */
void czo_export(const CzoEnv* e, int16_t* agents, int16_t* objs, int16_t* statics, int32_t* marks, int32_t* t) {
  int dyn_base[T_COUNT], sta_base[T_COUNT], nd = 0, ns = 0, slot_of[MAX_OBJ];
  memset(dyn_base, -1, sizeof(dyn_base)); memset(sta_base, -1, sizeof(sta_base));
  for (int m = 0; m < e->n_meta; ++m) {
    int ty = e->meta_type[m];
    if (IS_DYNAMIC(ty)) { dyn_base[ty] = nd; nd += e->meta_count[m]; }
    else if (IS_STATIC(ty)) { sta_base[ty] = ns; ns += e->meta_count[m]; }
  }
  memset(objs, 0, sizeof(int16_t) * 9 * nd); memset(statics, 0, sizeof(int16_t) * 4 * ns);
  for (int i = 0; i < MAX_OBJ; ++i) slot_of[i] = -1;
  for (int ty = 0; ty < T_COUNT; ++ty)
    if (IS_DYNAMIC(ty) && dyn_base[ty] >= 0)
      for (int j = 0; j < e->n_by_type[ty]; ++j) slot_of[e->by_type[ty][j]] = dyn_base[ty] + j;
  for (int i = 0; i < e->n_objs; ++i) {
    const Obj* o = &e->objs[i];
    if (IS_DYNAMIC(o->type) && slot_of[i] >= 0) {
      int16_t* r = objs + 9 * slot_of[i];
      r[0] = 1; r[1] = (int16_t)o->x; r[2] = (int16_t)o->y; r[3] = (int16_t)o->chopped; r[4] = (int16_t)o->blend; r[5] = (int16_t)o->free_flag;
    }
  }
  for (int i = 0; i < e->A; ++i) {
    const Agent* ag = &e->agents[i];
    int16_t* r = agents + 6 * i;
    r[0] = (int16_t)ag->x; r[1] = (int16_t)ag->y; r[2] = (int16_t)ag->orientation;
    r[3] = (int16_t)(ag->holding >= 0 ? slot_of[ag->holding] : -1); r[4] = (int16_t)ag->active; r[5] = (int16_t)ag->grace;
    if (ag->holding >= 0) { int16_t* q = objs + 9 * slot_of[ag->holding]; q[6] = 0; q[7] = (int16_t)i; q[8] = 0; }
  }
  for (int i = 0; i < e->n_objs; ++i) {
    const Obj* o = &e->objs[i];
    for (int k = 0; k < o->n_content; ++k) {
      int16_t* q = objs + 9 * slot_of[o->content[k]];
      if (IS_STATIC(o->type)) { q[6] = 1; q[7] = (int16_t)(o->y * e->W + o->x); q[8] = (int16_t)k; }
      else { q[6] = 2; q[7] = (int16_t)slot_of[i]; q[8] = (int16_t)k; }
    }
  }
  for (int ty = 0; ty < T_COUNT; ++ty)
    if (IS_STATIC(ty) && sta_base[ty] >= 0)
      for (int j = 0; j < e->n_by_type[ty]; ++j) {
        const Obj* o = &e->objs[e->by_type[ty][j]];
        int16_t* r = statics + 4 * (sta_base[ty] + j);
        r[0] = 1; r[1] = (int16_t)o->x; r[2] = (int16_t)o->y;
        r[3] = (int16_t)((o->ready ? 1 : 0) | (o->toggle ? 2 : 0) | (o->active ? 4 : 0) | (o->walkable ? 8 : 0) | (o->pressed ? 16 : 0));
      }
  for (int r = 0; r < e->R; ++r) {
    marks[r] = 0;
    for (int k = 0; k < e->n_nodes[r]; ++k) if (e->nodes[r][k].marked) marks[r] |= 1 << k;
  }
  *t = e->t;
}

uint32_t czo_error(const CzoEnv* e) { return e->error; }

/* Agent.move_to (world_objects.py:794-797) without the Floor bookkeeping of a walk: test helper for the directed
 * scenarios of SURVEY.md Appendix E, the counterpart of RefEnv.teleport / OracleEnv.teleport */
void czo_teleport(CzoEnv* e, int agent, int x, int y) {
  Agent* ag = &e->agents[agent];
  ag->x = x; ag->y = y;
  if (ag->holding >= 0) move_obj(e, ag->holding, x, y);
}
int czo_sizeof(void) { return (int)sizeof(CzoEnv); }

/* one step of n independent environments + every agent's observation (the CPU baseline's unit of work) */
void czo_batch_step(CzoEnv** envs, int n, const uint8_t* actions, double* obs, double* reward, uint8_t* term, uint8_t* trunc) {
  uint8_t rel[MAX_AGENTS];
  for (int i = 0; i < n; ++i) {
    CzoEnv* e = envs[i];
    const int A = e->A, L = czo_obs_len(e);
    czo_step(e, actions + (size_t)i * A, reward + (size_t)i * A, term + (size_t)i * A, trunc + (size_t)i * A, rel);
    for (int a = 0; a < A; ++a) czo_observe(e, a, obs + ((size_t)i * A + a) * L);
  }
}
