// cz_kernels.cu — kernels and the C ABI of libcz_b200.so (sm_100a only).
//
// Execution model ("tile" = 32 consecutive environments owned by one warp):
//   1. the warp loads the tile's packed state, structure-of-arrays in HBM, into shared-memory
//      columns with coalesced 128 B loads (lane = environment);
//   2. every lane advances its environment one step (cz_device.cuh) — scalar table-driven
//      code, no cross-lane traffic;
//   3. state, rewards and flags go back with coalesced stores;
//   4. the warp walks its 32 x A observation rows: the lanes fill one row (one lane per
//      observed slot) in a shared-memory staging buffer and a single elected lane hands the
//      2224-byte row to the TMA engine (cp.async.bulk shared->global, SASS UBLKCP), double
//      buffered so row r+1 is built while row r drains.
// The path is HBM-write bound (observations are 92 % of the bytes); tensor cores are unused.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <new>

#include "cz_device.cuh"

#define CZ_WARPS_PER_BLOCK 4
#define CZ_THREADS (32 * CZ_WARPS_PER_BLOCK)

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_OBSERVE = 2 };
enum { OBS_TMA = 0, OBS_STG = 1 };

// ---- TMA bulk store helpers (PTX ISA: cp.async.bulk, sm_90+) ------------------------------
__device__ __forceinline__ void cz_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cz_bulk_store_nocommit(void* gdst, const void* ssrc, uint32_t bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cz_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cz_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// One observation row (get_feature_vector, cooking_env.py:352-373) = table segments + computed slots.
//
// Computed slots (dynamic objects, agents, live Switch/Block): one lane per slot writes
// [x, y, flags..., 1] (or zeros when the slot is empty) as doubles into the staging buffer.
// The stores are fully unrolled and predicated so lanes with different feature counts do not
// serialise.  `row` points at staging element 0 == row element T.stage_lo.
__device__ __forceinline__ void cz_fill_computed(const CzDev& T, const uint32_t* sobj, const uint32_t* sag,
                                                 uint32_t sbits, uint32_t variant, int e, int agent, int lane,
                                                 double* row) {
  const uint32_t me = sag[agent * OSTRIDE + e];
  const int ax = me & 7u, ay = (me >> 3) & 7u;
  for (int q = lane; q < T.n_comp; q += 32) {
    const uint32_t d = __ldg(T.comp_slots + q);
    const uint32_t off = d & 0xFFFu, flen = (d >> 12) & 7u, kind = (d >> 15) & 3u, idx = (d >> 17) & 255u;
    uint32_t rec, fb4;
    bool present, self = false;
    if (kind == 1) {  // dynamic object: [!done, chopped, mashed] (world_objects.py:447,555,...)
      rec = sobj[idx * OSTRIDE + e];
      present = rec & O_PRESENT;
      uint32_t c = (rec >> 7) & 1u, m = (rec >> 8) & 1u;
      fb4 = ((c | m) ^ 1u) | c << 1 | m << 2;
    } else if (kind == 2) {  // agent: one-hot orientation; every agent, active or not (cooking_env.py:356)
      present = (int)idx < T.A;
      rec = present ? sag[idx * OSTRIDE + e] : 0u;
      fb4 = (1u << A_ORI(rec)) >> 1;
      self = (int)idx == agent;
    } else {  // live Switch / Block: [switch_active] / [walkable] (world_objects.py:174,221)
      uint32_t cell = __ldg(T.static_cells + variant * T.S + idx);
      present = cell != 0xFFu;
      rec = present ? cell : 0u;
      uint32_t g = __ldg(T.grid + variant * 64 + rec);
      fb4 = ((g & 15u) == ST_SWITCH ? (sbits >> (12 + (g >> 4))) : (sbits >> (16 + (g >> 4)))) & 1u;
    }
    const int x = rec & 7u, y = (rec >> 3) & 7u;
    // (x - ax) / W from a table of host-divided doubles; the observer's own entry is x / W (:364-368)
    double X = __ldg(T.xlut + (x - (self ? 0 : ax) + T.W - 1));
    double Y = __ldg(T.ylut + (y - (self ? 0 : ay) + T.H - 1));
    const uint32_t one = 1u << (flen - 1);  // the trailing 1 of every feature vector
    uint32_t fb = (fb4 & (one - 1u)) | one;
    if (!present) { X = 0.0; Y = 0.0; fb = 0u; }
    double* out = row + ((int)off - T.stage_lo);
    out[0] = X;
    out[1] = Y;
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (k < (int)flen) out[2 + k] = (fb >> k & 1u) ? 1.0 : 0.0;
  }
}

// Table segments: runs of static slots depend only on (layout variant, observer cell); copy them
// from the L1/L2-resident table straight to the row with 128-bit loads and stores.
__device__ __forceinline__ void cz_copy_table_segments(const CzDev& T, uint32_t variant, uint32_t cell, int lane,
                                                       double* __restrict__ gdst) {
  const double* src = T.obs_table + ((size_t)variant * 64 + cell) * T.tab_len;
#pragma unroll
  for (int sgi = 0; sgi < 2; ++sgi) {
    if (sgi < T.n_segs) {
      const double2* s2 = reinterpret_cast<const double2*>(src + T.segs[sgi][2]);
      double2* g2 = reinterpret_cast<double2*>(gdst + T.segs[sgi][0]);
      const int n2 = T.segs[sgi][1] >> 1;
      for (int k = lane; k < n2; k += 32) g2[k] = __ldg(s2 + k);
    }
  }
}

// shared-memory layout of one warp
struct WarpSmem {
  uint32_t obj[CZ_MAX_DYN * OSTRIDE];
  uint32_t ag[CZ_MAX_AGENTS * OSTRIDE];
  uint32_t sbits[32];
  uint32_t variant[32];
  uint32_t wobs[32];
};

template <int MODE, int OBS>
__global__ void __launch_bounds__(CZ_THREADS)
cz_env_kernel(const CzDev T, uint32_t* __restrict__ state, const uint8_t* __restrict__ actions,
              const int32_t* __restrict__ layout_ids, const uint8_t* __restrict__ recipe_ids,
              const uint8_t* __restrict__ mask, double* __restrict__ obs, double* __restrict__ reward,
              uint8_t* __restrict__ term, uint8_t* __restrict__ trunc, uint32_t* __restrict__ errflags,
              int n_envs, uint32_t flags, uint64_t seed, int64_t env_offset) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem* ws = reinterpret_cast<WarpSmem*>(smem_raw) + warp;
  // staging rows (the computed span of a row) live after the per-warp structs, 2 per warp, 16-byte aligned
  const size_t row_bytes = ((size_t)T.stage_len * 8 + 15) & ~(size_t)15;
  unsigned char* stage_base = smem_raw + ((sizeof(WarpSmem) * CZ_WARPS_PER_BLOCK + 15) & ~(size_t)15);
  double* stage0 = reinterpret_cast<double*>(stage_base + (size_t)(2 * warp) * row_bytes);
  double* stage1 = reinterpret_cast<double*>(stage_base + (size_t)(2 * warp + 1) * row_bytes);

  const int n_tiles = (n_envs + 31) >> 5;
  const int warps_total = gridDim.x * CZ_WARPS_PER_BLOCK;
  const int D = T.D, A = T.A;
  const size_t N = (size_t)n_envs;
  uint32_t* misc = state + (size_t)(D + A) * N;

  for (int tile = blockIdx.x * CZ_WARPS_PER_BLOCK + warp; tile < n_tiles; tile += warps_total) {
    const int env = tile * 32 + lane;
    const bool valid = env < n_envs;
    EnvRegs e;
    e.o = ws->obj + lane;
    e.ag = ws->ag + lane;
    e.err = 0;
    bool write_obs = valid;
    if (valid) {
      // ---- phase 1: state -> shared columns (coalesced: consecutive lanes, consecutive words)
      for (int s = 0; s < D; ++s) e.o[s * OSTRIDE] = state[(size_t)s * N + env];
      for (int i = 0; i < A; ++i) e.ag[i * OSTRIDE] = state[(size_t)(D + i) * N + env];
      e.sbits = misc[(size_t)CZ_ROW_SBITS * N + env];
      e.tinfo = misc[(size_t)CZ_ROW_TINFO * N + env];
      e.marks = misc[(size_t)CZ_ROW_MARKS * N + env];
      e.variant = misc[(size_t)CZ_ROW_VARIANT * N + env];
      e.rids = misc[(size_t)CZ_ROW_RECIPES * N + env];
      e.episode = misc[(size_t)CZ_ROW_EPISODE * N + env];

      bool do_reset = false;
      int layout = 0;
      if (MODE == MODE_RESET) {
        do_reset = mask == nullptr || mask[env] != 0;
        write_obs = do_reset;
        if (do_reset) {
          layout = layout_ids[env];
          uint32_t rids = 0;
          for (int r = 0; r < T.R; ++r)
            rids |= (uint32_t)(recipe_ids ? recipe_ids[(size_t)env * T.R + r] : __ldg(T.default_recipes + r)) << (8 * r);
          e.rids = rids;
        }
      } else if (MODE == MODE_STEP) {
        if ((flags & CZ_STEP_AUTO_RESET) && (e.tinfo & TI_DONE)) {
          do_reset = true;
          layout = (int)(cz_mix(seed, (uint64_t)(env_offset + env), (uint64_t)e.episode) % (uint64_t)T.P);
        }
      }

      if (do_reset) {
        // ---- CookingEnvironment.reset (cooking_env.py:178-210): pooled layout -> state
        const uint32_t* src = T.pool + (size_t)layout * T.rows;
        for (int s = 0; s < D; ++s) e.o[s * OSTRIDE] = __ldg(src + s);
        for (int i = 0; i < A; ++i) e.ag[i * OSTRIDE] = __ldg(src + D + i);
        e.sbits = __ldg(src + D + A + CZ_ROW_SBITS);
        e.tinfo = __ldg(src + D + A + CZ_ROW_TINFO);
        e.variant = __ldg(src + D + A + CZ_ROW_VARIANT);
        e.episode += 1;
        uint32_t marks = 0;
        for (int r = 0; r < T.R; ++r) marks |= cz_recipe_marks(T, e, (e.rids >> (8 * r)) & 255u) << (8 * r);
        e.marks = marks;
        if (MODE == MODE_STEP) {
          for (int i = 0; i < A; ++i) {
            reward[(size_t)env * A + i] = 0.0;
            term[(size_t)env * A + i] = 0;
            trunc[(size_t)env * A + i] = 0;
          }
        }
      } else if (MODE == MODE_STEP) {
        // ---- phase 2: one accumulated_step per lane
        uint32_t act = 0;
        for (int i = 0; i < A; ++i) act |= (uint32_t)actions[(size_t)env * A + i] << (8 * i);
        cz_step_env(T, e, act, reward + (size_t)env * A, term + (size_t)env * A, trunc + (size_t)env * A);
      }

      // ---- phase 3: shared columns -> state
      if (MODE != MODE_OBSERVE && (MODE == MODE_STEP || do_reset)) {
        for (int s = 0; s < D; ++s) state[(size_t)s * N + env] = e.o[s * OSTRIDE];
        for (int i = 0; i < A; ++i) state[(size_t)(D + i) * N + env] = e.ag[i * OSTRIDE];
        misc[(size_t)CZ_ROW_SBITS * N + env] = e.sbits;
        misc[(size_t)CZ_ROW_TINFO * N + env] = e.tinfo;
        misc[(size_t)CZ_ROW_MARKS * N + env] = e.marks;
        misc[(size_t)CZ_ROW_VARIANT * N + env] = e.variant;
        misc[(size_t)CZ_ROW_RECIPES * N + env] = e.rids;
        misc[(size_t)CZ_ROW_EPISODE * N + env] = e.episode;
        if (errflags && e.err) errflags[env] |= e.err;
      }
    }
    ws->sbits[lane] = e.sbits;
    ws->variant[lane] = e.variant;
    ws->wobs[lane] = write_obs ? 1u : 0u;
    __syncwarp();

    // ---- phase 4: observation rows, one (environment, agent) at a time, whole warp
    int buf = 0;
    const int n_here = min(32, n_envs - tile * 32);
    for (int le = 0; le < n_here; ++le) {
      if (!ws->wobs[le]) continue;
      const uint32_t sb = ws->sbits[le], var = ws->variant[le];
      for (int a = 0; a < A; ++a) {
        double* row = buf ? stage1 : stage0;
        double* gdst = obs + ((size_t)(tile * 32 + le) * A + a) * T.L;
        if (OBS == OBS_TMA) {
          // the bulk stores that last read this buffer (two rows ago) must have drained
          if (lane == 0) cz_bulk_wait_read<1>();
          __syncwarp();
          cz_fill_computed(T, ws->obj, ws->ag, sb, var, le, a, lane, row);
          cz_fence_async_smem();  // generic-proxy writes -> visible to the async proxy
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
              if (r < T.n_ranges)
                cz_bulk_store_nocommit(gdst + T.ranges[r][0], row + (T.ranges[r][0] - T.stage_lo), (uint32_t)(T.ranges[r][1] * 8));
            cz_bulk_commit();
          }
          cz_copy_table_segments(T, var, A_XY(ws->ag[a * OSTRIDE + le]), lane, gdst);
        } else {
          cz_fill_computed(T, ws->obj, ws->ag, sb, var, le, a, lane, row);
          __syncwarp();
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            if (r < T.n_ranges) {
              const double* srow = row + (T.ranges[r][0] - T.stage_lo);
              double* g = gdst + T.ranges[r][0];
              if ((T.L & 1) == 0) {  // rows and ranges are 16-byte aligned: 128-bit coalesced stores
                for (int k = lane; k < (T.ranges[r][1] >> 1); k += 32)
                  reinterpret_cast<double2*>(g)[k] = reinterpret_cast<const double2*>(srow)[k];
              } else {
                for (int k = lane; k < T.ranges[r][1]; k += 32) g[k] = srow[k];
              }
            }
          }
          if ((T.L & 1) == 0) cz_copy_table_segments(T, var, A_XY(ws->ag[a * OSTRIDE + le]), lane, gdst);
          __syncwarp();
        }
        buf ^= 1;
      }
    }
    if (OBS == OBS_TMA) {
      if (lane == 0) cz_bulk_wait_read<0>();  // staging and columns are reused by the next tile
    }
    __syncwarp();
  }
}

// =========================================================================================
// Host side: tables object and the C ABI
// =========================================================================================
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int cz_fail(int code, const char* fmt, const char* detail) {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
#define CZ_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) return cz_fail(CZ_ECUDA, #call ": %s", cudaGetErrorString(_e)); \
  } while (0)

struct cz_tables {
  CzDev dev;
  int device;
  int obs_path;
  int num_sms;
  void* allocs[32];
  int n_allocs;
  // scratch for cz_step_host
  uint8_t* d_actions; double* d_obs; double* d_reward; uint8_t* d_term; uint8_t* d_trunc;
  int scratch_envs;
};

template <typename Tp>
static int upload(cz_tables* t, const Tp* host, size_t count, const Tp** out) {
  *out = nullptr;
  if (count == 0) count = 1;
  void* d = nullptr;
  CZ_CUDA(cudaMalloc(&d, count * sizeof(Tp)));
  t->allocs[t->n_allocs++] = d;
  if (host) CZ_CUDA(cudaMemcpy(d, host, count * sizeof(Tp), cudaMemcpyHostToDevice));
  else CZ_CUDA(cudaMemset(d, 0, count * sizeof(Tp)));
  *out = (const Tp*)d;
  return CZ_OK;
}

static size_t cz_smem_bytes(const CzDev& T) {
  size_t row_bytes = ((size_t)T.stage_len * 8 + 15) & ~(size_t)15;
  return ((sizeof(WarpSmem) * CZ_WARPS_PER_BLOCK + 15) & ~(size_t)15) + 2 * CZ_WARPS_PER_BLOCK * row_bytes;
}

extern "C" int cz_abi_version(void) { return CZ_ABI_VERSION; }
extern "C" const char* cz_last_error(void) { return g_err; }
extern "C" uint64_t cz_launch_count(void) { return g_launches.load(); }

extern "C" uint64_t cz_layout_draw(uint64_t seed, uint64_t env, uint64_t episode) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (env + 1) + 0xD1B54A32D192ED03ull * (episode + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

extern "C" int cz_tables_create(const cz_table_desc* d, int device, cz_tables** out) {
  if (!d || !out) return cz_fail(CZ_EINVAL, "%s", "null argument");
  *out = nullptr;
  if (d->abi_version != CZ_ABI_VERSION) return cz_fail(CZ_EINVAL, "%s", "cz_table_desc.abi_version mismatch");
  if (d->width < 1 || d->width > 8 || d->height < 1 || d->height > 8) return cz_fail(CZ_ELIMIT, "%s", "level larger than 8x8");
  if (d->num_agents < 1 || d->num_agents > CZ_MAX_AGENTS) return cz_fail(CZ_ELIMIT, "%s", "num_agents out of range");
  if (d->num_recipes < 1 || d->num_recipes > CZ_MAX_RECIPES) return cz_fail(CZ_ELIMIT, "%s", "num_recipes out of range");
  if (d->num_dyn_slots < 1 || d->num_dyn_slots > CZ_MAX_DYN) return cz_fail(CZ_ELIMIT, "%s", "num_dyn_slots out of range");
  if (d->num_types < 1 || d->num_types > CZ_MAX_TYPES) return cz_fail(CZ_ELIMIT, "%s", "num_types out of range");
  if (d->num_static_slots < 0 || d->num_static_slots > CZ_MAX_STATIC_SLOTS) return cz_fail(CZ_ELIMIT, "%s", "num_static_slots out of range");
  if (d->obs_len < 1 || d->obs_len > 4095) return cz_fail(CZ_ELIMIT, "%s", "obs_len out of range");
  if (d->num_variants < 1 || d->num_layouts < 1 || d->num_book < 1 || d->num_book > 255) return cz_fail(CZ_EINVAL, "%s", "empty tables");
  if (d->max_steps < 1 || d->max_steps >= (1 << 20)) return cz_fail(CZ_ELIMIT, "%s", "max_steps out of range");
  if (d->grace_period < 0 || d->grace_period > 65535) return cz_fail(CZ_ELIMIT, "%s", "grace_period out of range");
  CZ_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CZ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return cz_fail(CZ_ECUDA, "%s", "libcz_b200 is built for sm_100a (B200) only");
  cz_tables* t = new (std::nothrow) cz_tables();
  if (!t) return cz_fail(CZ_EINVAL, "%s", "out of host memory");
  memset(t, 0, sizeof(*t));
  t->device = device;
  t->num_sms = prop.multiProcessorCount;
  const char* p = getenv("CZ_OBS_PATH");
  t->obs_path = (p && !strcmp(p, "stg")) ? OBS_STG : OBS_TMA;
  if (d->obs_len & 1) t->obs_path = OBS_STG;  // bulk copies need 16-byte rows
  CzDev& T = t->dev;
  T.W = d->width; T.H = d->height; T.A = d->num_agents; T.R = d->num_recipes; T.D = d->num_dyn_slots;
  T.S = d->num_static_slots; T.T = d->num_types; T.L = d->obs_len;
  T.n_comp = d->num_comp_slots; T.n_segs = d->num_obs_segs; T.n_ranges = d->num_obs_ranges; T.tab_len = d->obs_table_len;
  if (T.n_segs < 0 || T.n_segs > 2 || T.n_ranges < 0 || T.n_ranges > 3 || T.n_comp < 0 || T.tab_len < 0)
    { delete t; return cz_fail(CZ_EINVAL, "%s", "bad observation plan"); }
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) T.segs[i][j] = d->obs_segs[i][j];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) T.ranges[i][j] = d->obs_ranges[i][j];
  T.stage_lo = 0; T.stage_len = 2;
  if (T.n_ranges > 0) {
    T.stage_lo = T.ranges[0][0];
    T.stage_len = T.ranges[T.n_ranges - 1][0] + T.ranges[T.n_ranges - 1][1] - T.stage_lo;
  }
  if ((T.L & 1) && T.n_segs > 0) { delete t; return cz_fail(CZ_EINVAL, "%s", "table segments need an even obs_len"); }
  T.V = d->num_variants; T.P = d->num_layouts; T.B = d->num_book; T.max_steps = d->max_steps;
  T.end_all = d->end_all; T.grace = d->grace_period; T.n_switches = d->num_switches; T.n_blocks = d->num_blocks;
  T.rows = T.D + T.A + CZ_NUM_MISC_ROWS;
  T.r_node = d->reward_node; T.r_recipe = d->reward_recipe; T.r_penalty = d->reward_penalty; T.r_time = d->reward_time;
  T.respawn = d->respawn_rate; T.despawn = d->despawn_rate;
  int rc = CZ_OK;
#define UP(field, src, count) if (rc == CZ_OK) rc = upload(t, src, (size_t)(count), &T.field)
  UP(xlut, d->xlut, 2 * T.W - 1);
  UP(ylut, d->ylut, 2 * T.H - 1);
  UP(grid, d->grid, (size_t)T.V * 64);
  UP(static_cells, d->static_cells, (size_t)T.V * (T.S > 0 ? T.S : 1));
  UP(scan_order, d->scan_order, (size_t)T.V * T.D);
  UP(special_cells, d->special_cells, (size_t)T.V * 4 * CZ_MAX_SPECIAL);
  UP(static_masks, d->static_masks, (size_t)T.V * 8);
  UP(slot_type, d->slot_type, T.D);
  UP(type_flags, d->type_flags, T.T);
  UP(type_base, d->type_base, T.T);
  UP(type_count, d->type_count, T.T);
  UP(comp_slots, d->comp_slots, T.n_comp);
  UP(obs_table, d->obs_table, (size_t)T.V * 64 * T.tab_len);
  UP(recipe_nodes, d->recipe_nodes, (size_t)T.B * CZ_MAX_NODES);
  UP(recipe_len, d->recipe_len, T.B);
  UP(pool, d->pool, (size_t)T.P * T.rows);
  UP(default_recipes, d->default_recipes, T.R);
#undef UP
  if (rc != CZ_OK) { cz_tables_destroy(t); return rc; }
  size_t smem = cz_smem_bytes(T);
  if (smem > (size_t)prop.sharedMemPerBlockOptin) { cz_tables_destroy(t); return cz_fail(CZ_ELIMIT, "%s", "obs_len too large for shared memory staging"); }
#define SET_SMEM(M, O) CZ_CUDA(cudaFuncSetAttribute(cz_env_kernel<M, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
  SET_SMEM(MODE_STEP, OBS_TMA); SET_SMEM(MODE_STEP, OBS_STG);
  SET_SMEM(MODE_RESET, OBS_TMA); SET_SMEM(MODE_RESET, OBS_STG);
  SET_SMEM(MODE_OBSERVE, OBS_TMA); SET_SMEM(MODE_OBSERVE, OBS_STG);
#undef SET_SMEM
  *out = t;
  return CZ_OK;
}

extern "C" int cz_tables_destroy(cz_tables* t) {
  if (!t) return CZ_OK;
  cudaSetDevice(t->device);
  for (int i = 0; i < t->n_allocs; ++i) cudaFree(t->allocs[i]);
  if (t->d_actions) cudaFree(t->d_actions);
  if (t->d_obs) cudaFree(t->d_obs);
  if (t->d_reward) cudaFree(t->d_reward);
  if (t->d_term) cudaFree(t->d_term);
  if (t->d_trunc) cudaFree(t->d_trunc);
  delete t;
  return CZ_OK;
}

extern "C" int cz_state_rows(const cz_tables* t) { return t ? t->dev.rows : CZ_EINVAL; }

static int cz_grid(const cz_tables* t, int n_envs) {
  int tiles = (n_envs + 31) / 32;
  int blocks = (tiles + CZ_WARPS_PER_BLOCK - 1) / CZ_WARPS_PER_BLOCK;
  int cap = t->num_sms * 8;  // persistent tile loop beyond this many blocks
  return blocks < cap ? blocks : cap;
}

template <int MODE>
static int cz_launch(const cz_tables* t, uint32_t* state, const uint8_t* actions, const int32_t* layout_ids,
                     const uint8_t* recipe_ids, const uint8_t* mask, double* obs, double* reward, uint8_t* term,
                     uint8_t* trunc, uint32_t* err, int n_envs, uint32_t flags, uint64_t seed, int64_t env_offset,
                     void* stream) {
  if (!t || !state || !obs) return cz_fail(CZ_EINVAL, "%s", "null argument");
  if (n_envs <= 0) return CZ_OK;
  if (((uintptr_t)obs & 15) != 0) return cz_fail(CZ_EINVAL, "%s", "obs must be 16-byte aligned");
  size_t smem = cz_smem_bytes(t->dev);
  int grid = cz_grid(t, n_envs);
  cudaStream_t s = (cudaStream_t)stream;
  if (t->obs_path == OBS_TMA)
    cz_env_kernel<MODE, OBS_TMA><<<grid, CZ_THREADS, smem, s>>>(t->dev, state, actions, layout_ids, recipe_ids, mask, obs,
                                                               reward, term, trunc, err, n_envs, flags, seed, env_offset);
  else
    cz_env_kernel<MODE, OBS_STG><<<grid, CZ_THREADS, smem, s>>>(t->dev, state, actions, layout_ids, recipe_ids, mask, obs,
                                                               reward, term, trunc, err, n_envs, flags, seed, env_offset);
  g_launches.fetch_add(1);
  CZ_CUDA(cudaGetLastError());
  return CZ_OK;
}

extern "C" int cz_reset(const cz_tables* t, uint32_t* state, const int32_t* layout_ids, const uint8_t* recipe_ids,
                        const uint8_t* mask, double* obs, int n_envs, void* stream) {
  if (!layout_ids) return cz_fail(CZ_EINVAL, "%s", "layout_ids is required");
  return cz_launch<MODE_RESET>(t, state, nullptr, layout_ids, recipe_ids, mask, obs, nullptr, nullptr, nullptr, nullptr,
                               n_envs, 0, 0, 0, stream);
}

extern "C" int cz_step(const cz_tables* t, uint32_t* state, const uint8_t* actions, double* obs, double* reward,
                       uint8_t* terminated, uint8_t* truncated, uint32_t* error_flags, int n_envs, uint32_t flags,
                       uint64_t seed, int64_t env_offset, void* stream) {
  if (!actions || !reward || !terminated || !truncated) return cz_fail(CZ_EINVAL, "%s", "null argument");
  return cz_launch<MODE_STEP>(t, state, actions, nullptr, nullptr, nullptr, obs, reward, terminated, truncated,
                              error_flags, n_envs, flags, seed, env_offset, stream);
}

extern "C" int cz_observe(const cz_tables* t, const uint32_t* state, double* obs, int n_envs, void* stream) {
  return cz_launch<MODE_OBSERVE>(t, const_cast<uint32_t*>(state), nullptr, nullptr, nullptr, nullptr, obs, nullptr,
                                 nullptr, nullptr, nullptr, n_envs, 0, 0, 0, stream);
}

extern "C" int cz_step_host(cz_tables* t, uint32_t* state_dev, const uint8_t* actions_host, double* obs_host,
                            double* reward_host, uint8_t* terminated_host, uint8_t* truncated_host, int n_envs,
                            uint32_t flags, uint64_t seed, int64_t env_offset, void* stream) {
  if (!t || !actions_host || !obs_host || !reward_host || !terminated_host || !truncated_host)
    return cz_fail(CZ_EINVAL, "%s", "null argument");
  const CzDev& T = t->dev;
  if (n_envs > t->scratch_envs) {
    if (t->d_actions) { cudaFree(t->d_actions); cudaFree(t->d_obs); cudaFree(t->d_reward); cudaFree(t->d_term); cudaFree(t->d_trunc); }
    t->scratch_envs = 0;
    size_t na = (size_t)n_envs * T.A;
    CZ_CUDA(cudaMalloc((void**)&t->d_actions, na));
    CZ_CUDA(cudaMalloc((void**)&t->d_obs, na * T.L * sizeof(double)));
    CZ_CUDA(cudaMalloc((void**)&t->d_reward, na * sizeof(double)));
    CZ_CUDA(cudaMalloc((void**)&t->d_term, na));
    CZ_CUDA(cudaMalloc((void**)&t->d_trunc, na));
    t->scratch_envs = n_envs;
  }
  cudaStream_t s = (cudaStream_t)stream;
  size_t na = (size_t)n_envs * T.A;
  CZ_CUDA(cudaMemcpyAsync(t->d_actions, actions_host, na, cudaMemcpyHostToDevice, s));
  int rc = cz_step(t, state_dev, t->d_actions, t->d_obs, t->d_reward, t->d_term, t->d_trunc, nullptr, n_envs, flags, seed,
                   env_offset, stream);
  if (rc != CZ_OK) return rc;
  CZ_CUDA(cudaMemcpyAsync(obs_host, t->d_obs, na * T.L * sizeof(double), cudaMemcpyDeviceToHost, s));
  CZ_CUDA(cudaMemcpyAsync(reward_host, t->d_reward, na * sizeof(double), cudaMemcpyDeviceToHost, s));
  CZ_CUDA(cudaMemcpyAsync(terminated_host, t->d_term, na, cudaMemcpyDeviceToHost, s));
  CZ_CUDA(cudaMemcpyAsync(truncated_host, t->d_trunc, na, cudaMemcpyDeviceToHost, s));
  CZ_CUDA(cudaStreamSynchronize(s));
  return CZ_OK;
}
