// cz_obs32.cuh — float32 observation rows (reported separately from the f64 headline).
//
// obs32[e][a][k] == (float) get_feature_vector(a)[k] of environment e (cooking_env.py:352-373): every
// element of a row is either a host-divided table entry ((x - ax) / W, (y - ay) / H), 0 or 1, so the
// float32 row is the element-wise rounding of the float64 row and the tables are rounded once on the
// host (cz_tables_create).  One warp builds the A rows of one environment in shared memory (table
// segments, zeroed computed ranges, then one lane per computed slot) and streams them out with the
// widest store the environment's byte offset allows.  Works for every observation plan (no "simple"
// restriction).  Included by cz_kernels.cu.
#pragma once

#define CZ_OBS32_MAX_WARPS 8
#ifndef CZ_OBS32_TMA
#define CZ_OBS32_TMA 1  // the packed float32 writers hand their staging block to the TMA engine: one cp.async.bulk per warp
                        // (two-environment writer alone: 52.4 against 61.7 us per launch with lane copies; 0: A/B build)
#endif
#ifndef CZ_OBS32_UNROLL
#define CZ_OBS32_UNROLL 4  // zero-fill and copy-out loops of the float32 writers: the loop overhead is most of their instructions
#endif
constexpr int kObs32Unroll = CZ_OBS32_UNROLL;  // (#pragma unroll takes a constant expression, not a macro)

__global__ void __launch_bounds__(32 * CZ_OBS32_MAX_WARPS)
cz_obs32_kernel(const __grid_constant__ CzDev T, const uint32_t* __restrict__ state, float* __restrict__ obs, int n_envs,
                int warps_per_block) {
  extern __shared__ __align__(16) unsigned char smem_f32[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * warps_per_block + warp;
  if (warp >= warps_per_block || env >= n_envs) return;
  const int A = T.A, L = T.L, D = T.D;
  const size_t N = (size_t)n_envs;
  const int env_floats = A * L;
  float* stage = reinterpret_cast<float*>(smem_f32) + (size_t)warp * ((env_floats + 3) & ~3);
  const uint32_t* misc = state + (size_t)(D + A) * N;
  const uint32_t var = __ldg(misc + (size_t)CZ_ROW_VARIANT * N + env);
  const uint32_t sbits = __ldg(misc + (size_t)CZ_ROW_SBITS * N + env);

  // table segments and zeroed computed ranges of every row.  The plan is read into registers once, with
  // constant indices (a run-time index into the kernel parameters would copy the struct to local memory).
  const bool even = (L & 1) == 0;  // table runs exist only for even L and sit on even boundaries: 8-byte moves
  const int sh = even ? 1 : 0;
  const int s0o = T.segs[0][0] >> sh, s0n = T.n_segs > 0 ? T.segs[0][1] >> sh : 0, s0t = T.segs[0][2] >> sh;
  const int s1o = T.segs[1][0] >> sh, s1n = T.n_segs > 1 ? T.segs[1][1] >> sh : 0, s1t = T.segs[1][2] >> sh;
  const int r0o = T.ranges[0][0] >> sh, r0n = T.n_ranges > 0 ? T.ranges[0][1] >> sh : 0;
  const int r1o = T.ranges[1][0] >> sh, r1n = T.n_ranges > 1 ? T.ranges[1][1] >> sh : 0;
  const int r2o = T.ranges[2][0] >> sh, r2n = T.n_ranges > 2 ? T.ranges[2][1] >> sh : 0;
  for (int a = 0; a < A; ++a) {
    const uint32_t cell = __ldg(state + (size_t)(D + a) * N + env) & 63u;
    const float* tab = T.obs_table32 + ((size_t)var * 64 + cell) * T.tab_len;
    float* row = stage + a * L;
    if (even) {
      float2* row2 = reinterpret_cast<float2*>(row);
      const float2* tab2 = reinterpret_cast<const float2*>(tab);
      for (int k = lane; k < s0n; k += 32) row2[s0o + k] = __ldg(tab2 + s0t + k);
      for (int k = lane; k < s1n; k += 32) row2[s1o + k] = __ldg(tab2 + s1t + k);
      for (int k = lane; k < r0n; k += 32) row2[r0o + k] = make_float2(0.0f, 0.0f);
      for (int k = lane; k < r1n; k += 32) row2[r1o + k] = make_float2(0.0f, 0.0f);
      for (int k = lane; k < r2n; k += 32) row2[r2o + k] = make_float2(0.0f, 0.0f);
    } else {
      for (int k = lane; k < s0n; k += 32) row[s0o + k] = __ldg(tab + s0t + k);
      for (int k = lane; k < s1n; k += 32) row[s1o + k] = __ldg(tab + s1t + k);
      for (int k = lane; k < r0n; k += 32) row[r0o + k] = 0.0f;
      for (int k = lane; k < r1n; k += 32) row[r1o + k] = 0.0f;
      for (int k = lane; k < r2n; k += 32) row[r2o + k] = 0.0f;
    }
  }
  __syncwarp();

  // computed slots: one lane per slot, every observer
  for (int q = lane; q < T.n_comp; q += 32) {
    const uint32_t d = __ldg(T.comp_slots + q);
    const int off = d & 0xFFFu;
    const uint32_t flen = (d >> 12) & 7u, kind = (d >> 15) & 3u, idx = (d >> 17) & 255u;
    uint32_t rec = 0, fb4 = 0;
    bool present;
    if (kind != 0u) {
      const bool is_agent = kind == 2u;
      const bool exists = !is_agent || (int)idx < A;
      rec = exists ? __ldg(state + (size_t)(is_agent ? D + idx : idx) * N + env) : 0u;
      present = is_agent ? exists : (rec & O_PRESENT) != 0;
      const uint32_t c = (rec >> 7) & 1u, m = (rec >> 8) & 1u;
      fb4 = is_agent ? ((1u << A_ORI(rec)) >> 1) : (((c | m) ^ 1u) | c << 1 | m << 2);
    } else {  // static slot outside the table runs (live Switch / Block, or a plain one between computed slots)
      const uint32_t cell = __ldg(T.static_cells + var * T.S + idx);
      present = cell != 0xFFu;
      rec = present ? cell : 0u;
      const uint32_t g = __ldg(T.grid + var * 64 + rec);
      fb4 = ((g & 15u) == ST_SWITCH ? (sbits >> (12 + (g >> 4))) : (sbits >> (16 + (g >> 4)))) & 1u;
    }
    const uint32_t one = 1u << (flen - 1);
    const uint32_t fb = present ? ((fb4 & (one - 1u)) | one) : 0u;
    const int x = rec & 7u, y = (rec >> 3) & 7u;
    for (int a = 0; a < A; ++a) {
      const uint32_t me = __ldg(state + (size_t)(D + a) * N + env);
      const bool self = kind == 2u && (int)idx == a;  // the observer's own entry is x / W, y / H (cooking_env.py:364-368)
      float X = __ldg(T.xlut32 + (x - (self ? 0 : (int)(me & 7u)) + T.W - 1));
      float Y = __ldg(T.ylut32 + (y - (self ? 0 : (int)((me >> 3) & 7u)) + T.H - 1));
      if (!present) { X = 0.0f; Y = 0.0f; }
      float* out = stage + a * L + off;
      out[0] = X;
      out[1] = Y;
#pragma unroll
      for (uint32_t k = 0; k < 5; ++k)
        if (k < flen) out[2 + k] = (fb >> k & 1u) ? 1.0f : 0.0f;
    }
  }
  __syncwarp();

  float* g = obs + (size_t)env * env_floats;
  if ((env_floats & 3) == 0) {  // every environment starts 16-byte aligned
    for (int k = lane; k < env_floats >> 2; k += 32) reinterpret_cast<float4*>(g)[k] = reinterpret_cast<const float4*>(stage)[k];
  } else if ((env_floats & 1) == 0) {
    for (int k = lane; k < env_floats >> 1; k += 32) reinterpret_cast<float2*>(g)[k] = reinterpret_cast<const float2*>(stage)[k];
  } else {
    for (int k = lane; k < env_floats; k += 32) g[k] = stage[k];
  }
}

// Specialised writer for the packed (observer, slot) plans of the specialised step kernels (`simple` tables, NA * L a
// even): a lane owns one (observer, slot) pair and two table float2 per row, exactly the lane map of
// cz_obs_envs_kernel; the A rows are staged as one 16-byte aligned block and leave as float4.
// [x, y, flags..., 1] of one (observer, slot) pair as floats into its staging row (L2 float2 per row)
__device__ __forceinline__ void cz_pair_store32(const CzDev& T, const LaneSlot& ls, const PairRegs& p, float2* stage2, int L2) {
  if (ls.off < 0) return;
  const bool is_agent = ls.kind == 2, is_static = ls.kind == 0;
  const uint32_t rec = p.rec, me = p.me;
  const bool present = is_agent || (rec & O_PRESENT);
  const uint32_t c = (rec >> 7) & 1u, m = (rec >> 8) & 1u;
  const uint32_t fb4 = is_static ? p.static_fb : (is_agent ? ((1u << A_ORI(rec)) >> 1) : (((c | m) ^ 1u) | c << 1 | m << 2));
  const uint32_t one = 1u << (ls.flen - 1);
  const uint32_t fb = present ? ((fb4 & (one - 1u)) | one) : 0u;
  const bool self = is_agent && (int)ls.idx == ls.agent;
  const int x = rec & 7u, y = (rec >> 3) & 7u;
  float X = __ldg(T.xlut32 + (x - (self ? 0 : (int)(me & 7u)) + T.W - 1));
  float Y = __ldg(T.ylut32 + (y - (self ? 0 : (int)((me >> 3) & 7u)) + T.H - 1));
  if (!present) { X = 0.0f; Y = 0.0f; }
  float* out = reinterpret_cast<float*>(stage2 + ls.agent * L2) + T.stage_lo + ls.off;  // ls.off is relative to stage_lo
  out[0] = X;
  out[1] = Y;
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (k < (int)ls.flen) out[2 + k] = (fb >> k & 1u) ? 1.0f : 0.0f;
}

template <int NA, bool TWO>
__global__ void __launch_bounds__(32 * CZ_OBS32_MAX_WARPS, (TWO || NA >= 3) ? 5 : 8)
cz_obs32_fast_kernel(const __grid_constant__ CzDev T, const uint32_t* __restrict__ state, float* __restrict__ obs, int n_envs) {
  extern __shared__ __align__(16) unsigned char smem_f32[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int env = blockIdx.x * CZ_OBS32_MAX_WARPS + warp;
  if (env >= n_envs) return;
  const int D = T.D, tab2 = T.tab_len >> 1, L2 = T.L >> 1;
  const size_t N = (size_t)n_envs;
  float2* stage2 = reinterpret_cast<float2*>(smem_f32) + (size_t)warp * NA * L2;  // NA * L floats, 16-byte aligned

  const uint32_t var = __ldg(state + (size_t)(D + NA + CZ_ROW_VARIANT) * N + env);
  const LaneSlot ls = cz_lane_slot_packed(T, lane);
  const PairRegs p = cz_pair_load<NA>(T, state, N, env, var, ls);
  LaneSlot ls1;
  PairRegs p1;
  if constexpr (TWO) {
    ls1 = cz_lane_slot_packed(T, lane + 32);
    p1 = cz_pair_load<NA>(T, state, N, env, var, ls1);
  }
  const float2* tab = reinterpret_cast<const float2*>(T.obs_table32) + (size_t)var * 64 * tab2 + lane;
  // table segments of every row: loads first
  float2 v0[NA], v1[NA];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    uint32_t cell;
    if constexpr (TWO) cell = __ldg(state + (size_t)(D + a) * N + env) & 63u;
    else cell = __shfl_sync(0xffffffffu, p.me, a * T.n_comp) & 63u;  // lane a*n_comp observes for agent a
    if (ls.t0 >= 0) v0[a] = __ldg(tab + cell * tab2);
    if (ls.t1 >= 0) v1[a] = __ldg(tab + cell * tab2 + 32);
  }
  // the computed range of every row starts as zeros (never-occupied slots stay zero)
  {
    const int o2 = T.ranges[0][0] >> 1, n2 = T.ranges[0][1] >> 1;
#pragma unroll
    for (int a = 0; a < NA; ++a)
      for (int k = lane; k < n2; k += 32) stage2[a * L2 + o2 + k] = make_float2(0.0f, 0.0f);
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    if (ls.t0 >= 0) stage2[a * L2 + ls.t0] = v0[a];
    if (ls.t1 >= 0) stage2[a * L2 + ls.t1] = v1[a];
  }
  __syncwarp();
  cz_pair_store32(T, ls, p, stage2, L2);
  if constexpr (TWO) cz_pair_store32(T, ls1, p1, stage2, L2);
#if CZ_OBS32_TMA
  if (((NA * T.L) & 3) == 0) {  // 16-byte aligned, a multiple of 16 bytes: one bulk store for the environment's rows
    cz_fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      cz_bulk_store_stream(obs + (size_t)env * NA * T.L, stage2, (uint32_t)(NA * T.L) * 4u);
      cz_bulk_commit();
      cz_bulk_wait_read<0>();  // the staging block must outlive the read
    }
    return;
  }
#endif
  __syncwarp();
  if (((NA * T.L) & 3) == 0) {  // every environment (and every warp's staging block) starts 16-byte aligned
    float4* g4 = reinterpret_cast<float4*>(obs + (size_t)env * NA * T.L);
    const float4* s4 = reinterpret_cast<const float4*>(stage2);
    for (int k = lane; k < (NA * T.L) >> 2; k += 32) g4[k] = s4[k];
  } else {
    float2* g2 = reinterpret_cast<float2*>(obs + (size_t)env * NA * T.L);
    for (int k = lane; k < NA * L2; k += 32) g2[k] = stage2[k];
  }
}


// Two environments per warp (packed plans with at most 32 pairs): the float32 rows of one environment are only 4.3
// 16-byte elements per lane, so the one-environment writer spends its life waiting for its two levels of loads.  Here a
// warp issues the state loads of two neighbouring environments together, the table runs go global -> shared with
// cp.async (no registers), and the 2 * NA rows leave as one contiguous float4 stream.
__device__ __forceinline__ void cz_cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

#ifndef CZ_OBS32_PAIR_BLOCKS
#define CZ_OBS32_PAIR_BLOCKS 6
#endif
template <int NA>
__global__ void __launch_bounds__(32 * CZ_OBS32_MAX_WARPS, CZ_OBS32_PAIR_BLOCKS)
cz_obs32_pair_kernel(const __grid_constant__ CzDev T, const uint32_t* __restrict__ state, float* __restrict__ obs, int n_envs) {
  extern __shared__ __align__(16) unsigned char smem_f32[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int env0 = (blockIdx.x * CZ_OBS32_MAX_WARPS + warp) * 2;
  if (env0 >= n_envs) return;
  const bool two = env0 + 1 < n_envs;
  const int env1 = two ? env0 + 1 : env0;  // a lone last environment is built twice and written once
  const int D = T.D, tab2 = T.tab_len >> 1, L2 = T.L >> 1;
  const size_t N = (size_t)n_envs;
  float2* stage2 = reinterpret_cast<float2*>(smem_f32) + (size_t)warp * 2 * NA * L2;

  const uint32_t var0 = __ldg(state + (size_t)(D + NA + CZ_ROW_VARIANT) * N + env0);
  const uint32_t var1 = __ldg(state + (size_t)(D + NA + CZ_ROW_VARIANT) * N + env1);
  const LaneSlot ls = cz_lane_slot_packed(T, lane);
  const PairRegs p0 = cz_pair_load<NA>(T, state, N, env0, var0, ls);
  const PairRegs p1 = cz_pair_load<NA>(T, state, N, env1, var1, ls);
  {
    const float2* tab0 = reinterpret_cast<const float2*>(T.obs_table32) + (size_t)var0 * 64 * tab2 + lane;
    const float2* tab1 = reinterpret_cast<const float2*>(T.obs_table32) + (size_t)var1 * 64 * tab2 + lane;
#pragma unroll
    for (int a = 0; a < NA; ++a) {  // lane a * n_comp observes for agent a
      const uint32_t c0 = __shfl_sync(0xffffffffu, p0.me, a * T.n_comp) & 63u;
      const uint32_t c1 = __shfl_sync(0xffffffffu, p1.me, a * T.n_comp) & 63u;
      if (ls.t0 >= 0) {
        cz_cp_async8(stage2 + a * L2 + ls.t0, tab0 + c0 * tab2);
        cz_cp_async8(stage2 + (NA + a) * L2 + ls.t0, tab1 + c1 * tab2);
      }
      if (ls.t1 >= 0) {
        cz_cp_async8(stage2 + a * L2 + ls.t1, tab0 + c0 * tab2 + 32);
        cz_cp_async8(stage2 + (NA + a) * L2 + ls.t1, tab1 + c1 * tab2 + 32);
      }
    }
  }
  {  // the computed range of every row starts as zeros (never-occupied slots stay zero)
    const int o2 = T.ranges[0][0] >> 1, n2 = T.ranges[0][1] >> 1;
#pragma unroll
    for (int a = 0; a < 2 * NA; ++a) {
#pragma unroll kObs32Unroll
      for (int k = lane; k < n2; k += 32) stage2[a * L2 + o2 + k] = make_float2(0.0f, 0.0f);
    }
  }
  __syncwarp();
  cz_pair_store32(T, ls, p0, stage2, L2);
  cz_pair_store32(T, ls, p1, stage2 + NA * L2, L2);
  asm volatile("cp.async.wait_all;" ::: "memory");
#if CZ_OBS32_TMA
  cz_fence_async_smem();  // generic-proxy and cp.async writes -> visible to the async proxy
  __syncwarp();
  float* g = obs + (size_t)env0 * NA * T.L;  // env0 is even and NA * L is even: 16-byte aligned
  if (two) {
    if (lane == 0) {  // both environments' rows are one contiguous block: a single bulk store
      cz_bulk_store_stream(g, stage2, (uint32_t)(2 * NA * T.L) * 4u);
      cz_bulk_commit();
      cz_bulk_wait_read<0>();  // the staging block must outlive the read
    }
    return;
  }
#else
  __syncwarp();
  float* g = obs + (size_t)env0 * NA * T.L;  // env0 is even and NA * L is even: 16-byte aligned
#endif
  if (two) {
    float4* g4 = reinterpret_cast<float4*>(g);
    const float4* s4 = reinterpret_cast<const float4*>(stage2);
#pragma unroll kObs32Unroll
    for (int k = lane; k < (NA * T.L) >> 1; k += 32) g4[k] = s4[k];
  } else {
    float2* g2 = reinterpret_cast<float2*>(g);
    for (int k = lane; k < NA * L2; k += 32) g2[k] = stage2[k];
  }
}

// `alone`: nothing else is meant to share the SMs with this launch (in-place step, cz_observe_f32).  The two-environments-per-warp
// kernel fills the shared memory (6 blocks x 35.6 KB); beside the dynamics kernel of the pipelined mode it measured
// 1.47 G env-steps/s against 1.68 G for the one-environment kernel, so the pipelined step keeps the latter.
static int cz_launch_obs32(const cz_tables* t, const uint32_t* state, float* obs, int n_envs, cudaStream_t s, bool alone = true) {
  if (!t || !state || !obs) return cz_fail(CZ_EINVAL, "%s", "null argument");
  if (n_envs <= 0) return CZ_OK;
  if (((uintptr_t)obs & 15) != 0) return cz_fail(CZ_EINVAL, "%s", "obs must be 16-byte aligned");
  if ((alone || t->obs32_pair == 2) && t->simple && t->obs32_pair && n_envs >= 2 * CZ_OBS32_MAX_WARPS * 148) {  // two environments per warp
    const int blocks = (n_envs + 2 * CZ_OBS32_MAX_WARPS - 1) / (2 * CZ_OBS32_MAX_WARPS);
    const size_t smem = (size_t)CZ_OBS32_MAX_WARPS * 2 * t->dev.A * t->dev.L * 4;
    if (smem <= 48 * 1024) {
      switch (t->dev.A) {
        case 1: cz_obs32_pair_kernel<1><<<blocks, 32 * CZ_OBS32_MAX_WARPS, smem, s>>>(t->dev, state, obs, n_envs); break;
        case 2: cz_obs32_pair_kernel<2><<<blocks, 32 * CZ_OBS32_MAX_WARPS, smem, s>>>(t->dev, state, obs, n_envs); break;
        default: goto one_env_per_warp;  // 3-4 agents: the staging block of two environments is too large to keep occupancy
      }
      g_launches.fetch_add(1);
      CZ_CUDA(cudaGetLastError());
      return CZ_OK;
    }
  }
one_env_per_warp:
  if (t->simple || t->simple2) {
    const int blocks = (n_envs + CZ_OBS32_MAX_WARPS - 1) / CZ_OBS32_MAX_WARPS;
    const size_t smem = (size_t)CZ_OBS32_MAX_WARPS * t->dev.A * t->dev.L * 4;
    if (smem <= 48 * 1024) {
#define CZ_O32_GO(NA)                                                                                                       \
  if (t->simple2) cz_obs32_fast_kernel<NA, true><<<blocks, 32 * CZ_OBS32_MAX_WARPS, smem, s>>>(t->dev, state, obs, n_envs);  \
  else cz_obs32_fast_kernel<NA, false><<<blocks, 32 * CZ_OBS32_MAX_WARPS, smem, s>>>(t->dev, state, obs, n_envs)
      switch (t->dev.A) {
        case 1: CZ_O32_GO(1); break;
        case 2: CZ_O32_GO(2); break;
        case 3: CZ_O32_GO(3); break;
        default: CZ_O32_GO(4); break;
      }
#undef CZ_O32_GO
      g_launches.fetch_add(1);
      CZ_CUDA(cudaGetLastError());
      return CZ_OK;
    }
  }
  const size_t per_warp = (size_t)((t->dev.A * t->dev.L + 3) & ~3) * 4;
  int warps = (int)(((size_t)96 * 1024) / per_warp);  // keep at least two blocks per SM resident
  if (warps > CZ_OBS32_MAX_WARPS) warps = CZ_OBS32_MAX_WARPS;
  if (warps < 1) warps = 1;
  const size_t smem = per_warp * warps;
  if (smem > t->smem_optin) return cz_fail(CZ_ELIMIT, "%s", "observation rows too long for the float32 writer");
  const int blocks = (n_envs + warps - 1) / warps;
  cz_obs32_kernel<<<blocks, 32 * warps, smem, s>>>(t->dev, state, obs, n_envs, warps);
  g_launches.fetch_add(1);
  CZ_CUDA(cudaGetLastError());
  return CZ_OK;
}
