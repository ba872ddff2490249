// pgtg_kernels.cu -- sm_100a kernels and the CUDA backend of the C ABI (include/pgtg_b200.h).
//
// One fused launch per tick: every CTA owns a contiguous slice of B envs and runs
//   stage -> step (1 env / thread) -> ballot+scan compaction of done envs -> on-device reset /
//   procedural map regeneration by the first n_done threads -> observation bit assembly in shared
//   memory -> vectorised expansion to the int8 observation planes (contiguous 16-byte stores).
// Nothing here is a dense contraction, so no tensor cores: the kernel is bounded by HBM traffic
// (observation write + SoA state scan), see DESIGN.md for the byte accounting.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "pgtg_phases.cuh"

struct pgtg_env;
static thread_local cudaError_t g_cuda_err = cudaSuccess;
static const char* bk_error() { return cudaGetErrorString(g_cuda_err); }
static int ck(cudaError_t e) { if (e != cudaSuccess) { g_cuda_err = e; return -1; } return 0; }
static void* bk_alloc(size_t n) { void* p = nullptr; if (ck(cudaMalloc(&p, n))) return nullptr; return p; }
static void bk_free(void* p) { cudaFree(p); }
static int bk_set_device(int d) { return ck(cudaSetDevice(d)); }
static int bk_h2d(void* d, const void* s, size_t n, void* st) { return ck(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, (cudaStream_t)st)); }
static int bk_d2h(void* d, const void* s, size_t n, void* st) { return ck(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, (cudaStream_t)st)); }
static int bk_memset(void* d, int v, size_t n) { return ck(cudaMemset(d, v, n)); }
static int bk_memset_async(void* d, int v, size_t n, void* st) { return ck(cudaMemsetAsync(d, v, n, (cudaStream_t)st)); }
static int bk_sync(void* st) { return ck(st ? cudaStreamSynchronize((cudaStream_t)st) : cudaDeviceSynchronize()); }
// side stream (highest priority) for the persistent map-generation kernel, and its events
static int bk_side_create(void** stream, void** ev_tick, void** ev_map0, void** ev_map1, int* sm_count) {
  int lo = 0, hi = 0, dev = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStream_t s; cudaEvent_t a, b, c;
  if (ck(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi))) return -1;  // priority made no measurable difference
  if (ck(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)) || ck(cudaEventCreateWithFlags(&b, cudaEventDisableTiming)) ||
      ck(cudaEventCreateWithFlags(&c, cudaEventDisableTiming))) return -1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
  *stream = s; *ev_tick = a; *ev_map0 = b; *ev_map1 = c;
  return 0;
}
static void bk_side_destroy(void* stream, void* a, void* b, void* c) {
  if (stream) { cudaStreamSynchronize((cudaStream_t)stream); cudaStreamDestroy((cudaStream_t)stream); }
  if (a) cudaEventDestroy((cudaEvent_t)a);
  if (b) cudaEventDestroy((cudaEvent_t)b);
  if (c) cudaEventDestroy((cudaEvent_t)c);
}
static int bk_stream_wait(void* st, void* ev) { return ck(cudaStreamWaitEvent((cudaStream_t)st, (cudaEvent_t)ev, 0)); }
static int bk_dl_device_type() { return 2; }  // kDLCUDA
static void* bk_event_create() { cudaEvent_t ev; if (ck(cudaEventCreate(&ev))) return nullptr; return ev; }
static void bk_event_destroy(void* ev) { cudaEventDestroy((cudaEvent_t)ev); }
static int bk_event_record(void* ev, void* st) { return ck(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)st)); }
static double bk_event_elapsed(void* a, void* b) { float ms = 0; cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b); return ms; }
static int bk_pick_block(const pgtg::DevCfg& c, int* block, size_t* smem);
static int bk_launch(pgtg_env*, int mode, const uint8_t* mask, const int64_t* seeds, const void* actions, int action_bytes, void* stream);
static int bk_stats_reduce(pgtg_env*, void* stream);
static int bk_stats_reset(pgtg_env*, void* stream);
static int bk_flatten(pgtg_env*, void* stream);
static int bk_conn_table_max_bits() { return 24; }
static int bk_build_conn_table(pgtg_env*, uint32_t* table_dev);
static int bk_build_path_table(pgtg_env*, uint64_t* table_dev);

#include "pgtg_api_impl.hpp"

namespace pgtg {

constexpr int STATS_STRIDE = 8;
#ifndef PGTG_MIN_BLOCKS
#define PGTG_MIN_BLOCKS 8
#endif
#ifndef PGTG_MAPGEN_TABLED_MIN_BLOCKS
#define PGTG_MAPGEN_TABLED_MIN_BLOCKS 16  /* 32 registers, no spills: the tabled generator has no flood fill / BFS state */
#endif
#ifndef PGTG_LEAN_MIN_BLOCKS
#define PGTG_LEAN_MIN_BLOCKS 8
#endif
#ifndef PGTG_MAPGEN_MIN_BLOCKS
#define PGTG_MAPGEN_MIN_BLOCKS 12
#endif

// per-CTA episode statistics row (no cross-CTA atomics on the hot path)
struct StatsArgs {
  double* rows;  // [gridDim.x][8]
};

// -DPGTG_PHASE_CLOCKS: profiling build (never the shipped one; load it through PGTG_B200_LIB):
// per-warp SM-clock time of each phase of the ring-fed tick, summed into g_phase_clk and printed
// by the statistics reduction. The macros expand to nothing in the normal build.
#ifdef PGTG_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[16];
#define PG_CLK_INIT long long clk_prev = clock64();
#define PG_CLK(i) { if ((threadIdx.x & 31) == 0) { long long t_ = clock64(); atomicAdd(&g_phase_clk[i], (unsigned long long)(t_ - clk_prev)); clk_prev = t_; } }
#else
#define PG_CLK_INIT
#define PG_CLK(i)
#endif

template <int RNG, int MODE, int TMAX, bool PREGEN, bool LEAN = false>
__global__ void __launch_bounds__(128, LEAN ? PGTG_LEAN_MIN_BLOCKS : PGTG_MIN_BLOCKS) pgtg_tick_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ DevPtrs p,
                                                        const uint8_t* __restrict__ mask, const int64_t* __restrict__ seeds,
                                                        const void* __restrict__ actions, int action_bytes, StatsArgs sa,
                                                        const __grid_constant__ SharedLayout layout) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int B = blockDim.x, tid = threadIdx.x;
  const int env0 = blockIdx.x * B;
  const int nvalid = min(B, c.N - env0);
  const int env = env0 + tid;
  const bool valid = tid < nvalid;
  const int lane = tid & 31, warp = tid >> 5, nwarps = B >> 5;
  BlockShared sh = carve_layout(smem, layout);

  PG_CLK_INIT
  phase_stage(c, p, sh, tid, B, env0, nvalid, true, !(PREGEN && MODE == MODE_STEP));
  if (tid < 8) { sh.counters[8 + tid] = 0; sh.dsum[tid] = 0.0; }
  PG_CLK(0)
  __syncthreads();
  PG_CLK(1)

  bool done = false;
  if (MODE == MODE_STEP) {
    StepResult r;
    r.outcome = 0; r.ep_return = 0;
    int len = 0;
    EnvRegs er;  // lean instantiation: the env's registers stay in registers from step to emit
    if (valid) {
      int a = action_bytes == 8 ? (int)((const long long*)actions)[env] : ((const int*)actions)[env];
      if (LEAN) { r = phase_step_regs<RNG, true>(c, p, sh, tid, env, a, er); len = (int)er.elapsed; }
      else { r = phase_step<RNG, false>(c, p, sh, tid, env, a); len = (int)sh.regs[tid].elapsed; }
      done = r.outcome != 0;
      len = done ? len : 0;
    }
    PG_CLK(2)
    // episode statistics: ballots for the counters, warp reductions for the sums, one row per CTA
    unsigned any = __ballot_sync(0xffffffffu, done);
    if (any) {
      unsigned g = __ballot_sync(0xffffffffu, r.outcome == 2), cr = __ballot_sync(0xffffffffu, r.outcome == 1),
               tr = __ballot_sync(0xffffffffu, r.outcome == 3);
      int lsum = __reduce_add_sync(0xffffffffu, len);
      double rs = r.ep_return;
      for (int o = 16; o > 0; o >>= 1) rs += __shfl_down_sync(0xffffffffu, rs, o);
      if (lane == 0) {
        atomicAdd(&sh.counters[8], __popc(g)); atomicAdd(&sh.counters[9], __popc(cr)); atomicAdd(&sh.counters[10], __popc(tr));
        atomicAdd(&sh.counters[11], lsum); atomicAdd(&sh.counters[12], __popc(any));
        atomicAdd(&sh.dsum[0], rs);
      }
    }
    if (LEAN || (PREGEN && !c.write_final_obs)) {
      // Hot configuration (next maps come from the ring, no terminal-observation output): the
      // reset is a cheap swap, so every finished env is reset by its own thread and the CTA
      // needs one barrier only, the one in front of the byte expansion. Map requests are queued
      // per warp; the atomic's round trip hides behind the observation emit.
      uint32_t k = 0, qbase = 0;
      if (any && lane == 0) qbase = atomicAdd(p.regen_count + p.parity, (uint32_t)__popc(any));
      PG_CLK(3)
      if (done) {
        if (LEAN) { k = er.episode + 1u; phase_reset_regs<RNG, TMAX, true, true>(c, p, sh, tid, env, er); }
        else { k = sh.regs[tid].episode + 1u; phase_reset<RNG, TMAX, true, false>(c, p, sh, tid, env); }  // k: the episode this env is about to start
      }
      PG_CLK(4)
      if (valid) { if (LEAN) phase_emit_regs<true>(c, p, sh, tid, env, false, er); else phase_emit<false>(c, p, sh, tid, env, false); }
      PG_CLK(5)
      if (any) {
        qbase = __shfl_sync(0xffffffffu, qbase, 0);
        if (done) {
          uint2 q; q.x = (uint32_t)env; q.y = k + 2u;
          p.regen_list[(size_t)p.parity * 2 * c.N + qbase + __popc(any & ((1u << lane) - 1u))] = q;
        }
      }
      PG_CLK(6)
      __syncthreads();
      PG_CLK(7)
      if (tid == 0 && sh.counters[12]) {
        double* row = sa.rows + (size_t)blockIdx.x * STATS_STRIDE;
        row[0] += sh.counters[12]; row[1] += sh.dsum[0]; row[2] += sh.counters[11];
        row[3] += sh.counters[8]; row[4] += sh.counters[9]; row[5] += sh.counters[10];
      }
      phase_expand(c, p.obs_map, sh, tid, B, env0, nvalid);
      PG_CLK(8)
      return;
    }
  } else if (MODE == MODE_RESET) {
    if (valid) {
      EnvRegs e = load_regs(c, p, env);
      if (!mask || mask[env]) {
        if (seeds) { p.key[env] = (uint64_t)seeds[env]; e.episode = 0; }
        p.ep_return[env] = 0.0;
        done = true;
      }
      sh.regs[tid] = e;
    }
  } else {
    if (valid) sh.regs[tid] = load_regs(c, p, env);
  }

  if (LEAN) return;  // (the lean instantiation is step-mode only and has returned above)

  // compaction of the done envs: warp ballot + CTA scan -> dense list in shared memory
  unsigned ballot = __ballot_sync(0xffffffffu, done);
  if (lane == 0) sh.counters[1 + warp] = __popc(ballot);
  __syncthreads();
  int base = 0, n_done = 0;
  for (int w = 0; w < nwarps; w++) { int v = sh.counters[1 + w]; if (w < warp) base += v; n_done += v; }
  if (done) sh.done_list[base + __popc(ballot & ((1u << lane) - 1u))] = tid;
  if (MODE == MODE_STEP && tid == 0 && n_done) {
    double* row = sa.rows + (size_t)blockIdx.x * STATS_STRIDE;
    row[0] += n_done; row[1] += sh.dsum[0]; row[2] += sh.counters[11];
    row[3] += sh.counters[8]; row[4] += sh.counters[9]; row[5] += sh.counters[10];
  }
  __syncthreads();

  if (MODE != MODE_OBSERVE && c.pregen && n_done) {
    // CTA-uniform: queue map requests for the map-generation kernel. An env that starts episode k
    // frees ring slot (k & 1): ask for the map of episode k + 2 (a full reset also needs k + 1).
    const int per = MODE == MODE_RESET ? 2 : 1;
    if (tid == 0) sh.counters[20] = (int)atomicAdd(p.regen_count + p.parity, (uint32_t)(n_done * per));
    __syncthreads();
    if (tid < n_done) {
      int local = sh.done_list[tid];
      uint32_t k = sh.regs[local].episode + 1u;  // the episode this env is about to start
      uint2* q = p.regen_list + (size_t)p.parity * 2 * c.N + sh.counters[20] + tid * per;
      uint2 r; r.x = (uint32_t)(env0 + local);
      if (MODE == MODE_RESET) { r.y = k + 1u; q[0] = r; r.y = k + 2u; q[1] = r; }
      else { r.y = k + 2u; q[0] = r; }
    }
  }

  if (MODE == MODE_STEP && c.write_final_obs && n_done) {  // CTA-uniform condition
    if (done) phase_emit(c, p, sh, tid, env, true);
    __syncthreads();
    phase_expand_final(c, p.f_obs_map, sh, tid, B, env0, n_done);
    __syncthreads();
    for (int i = tid; i < sh.bits_words; i += B) sh.bits[i] = 0;
    __syncthreads();
  }

  if (MODE != MODE_OBSERVE) {
    if (tid < n_done) {
      int local = sh.done_list[tid];
      phase_reset<RNG, TMAX, PREGEN>(c, p, sh, local, env0 + local);
    }
    __syncthreads();
  }
  if (valid) phase_emit(c, p, sh, tid, env, false);
  __syncthreads();
  phase_expand(c, p.obs_map, sh, tid, B, env0, nvalid);
}

// Map generation ahead of time: dense over the envs queued by the tick that just ran (every lane
// busy, no CTA barrier after the staging, tiny shared-memory footprint -> high occupancy). The
// loop is grid-stride so that the launch code may also run it as a small persistent grid
// (PGTG_MAPGEN_CTAS_PER_SM); the default is one request per thread. TABLED: see generate_map.
template <int RNG, int TMAX, bool TABLED = false>
__global__ void __launch_bounds__(128, TABLED ? PGTG_MAPGEN_TABLED_MIN_BLOCKS : PGTG_MAPGEN_MIN_BLOCKS) pgtg_mapgen_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ DevPtrs p, int parity,
                                                                                const __grid_constant__ SharedLayout layout) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t count = p.regen_count[parity];
  if (blockIdx.x * blockDim.x >= count) return;
  BlockShared sh = carve_layout(smem, layout);
  stage_tables(c, p, sh, threadIdx.x, blockDim.x);
  __syncthreads();
  const uint2* list = p.regen_list + (size_t)parity * 2 * c.N;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    uint2 r = list[i];
    phase_pregenerate<RNG, TMAX, TABLED>(c, p, sh, threadIdx.x, (int)r.x, r.y);
  }
}

// one thread per 32-bit word of the start-goal connectivity table (32 subgraphs each)
__global__ void pgtg_build_conn_table_kernel(const __grid_constant__ DevCfg c, uint32_t* __restrict__ table, int s, int g) {
  const uint32_t words = (1u << c.conn_bits) / 32u + ((1u << c.conn_bits) < 32u ? 1u : 0u);
  const uint32_t rowmask = (1u << (c.W - 1)) - 1u, emask = (1u << c.conn_ne) - 1u;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
    uint32_t bits = 0;
    for (uint32_t b = 0; b < 32; b++) {
      uint32_t idx = w * 32u + b, ec = idx & emask, so = idx >> c.conn_ne, e = 0;
      for (int r = 0; r < c.H; r++) e |= ((ec >> (r * (c.W - 1))) & rowmask) << (r * c.W);
      if (flood_connected32(c.W, e, so, s, g)) bits |= 1u << b;
    }
    table[w] = bits;
  }
}

struct FlatOrder { int plane[PGTG_MAX_CHANNELS]; };

// one thread per entry of the subgoal-path table
__global__ void __launch_bounds__(128) pgtg_build_path_table_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ DevPtrs p, uint64_t* __restrict__ table) {
  extern __shared__ __align__(16) unsigned char smem[];
  BlockShared sh = carve_mapgen(smem, c, blockDim.x);
  stage_tables(c, p, sh, threadIdx.x, blockDim.x);
  __syncthreads();
  const uint32_t total = 1u << c.conn_bits;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x)
    table[g] = path_table_entry(c, *sh.lut, g, sh.tiles + threadIdx.x * c.tile_stride);
}

// FlattenObservation view: one thread per output float, coalesced float32 stores; reads the int8
// planes through L2 (they were just written by the tick).
__global__ void pgtg_flatten_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ DevPtrs p, FlatOrder order,
                                    float* __restrict__ out, int dim) {
  const int PP = c.P * c.P, map_dim = c.C * PP, nsd_dim = c.use_nsd ? 9 : 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)c.N * dim; i += (size_t)gridDim.x * blockDim.x) {
    int env = (int)(i / dim), j = (int)(i - (size_t)env * dim);
    float v;
    if (j < map_dim) {
      int k = j / PP, cell = j - k * PP;
      v = (float)p.obs_map[((size_t)env * c.C + order.plane[k]) * PP + cell];
    } else if (j < map_dim + nsd_dim) {
      v = (p.obs_nsd[env] + 1 == j - map_dim) ? 1.0f : 0.0f;  // Discrete(9, start=-1) one-hot
    } else if (j < map_dim + nsd_dim + 18) {
      int q = j - map_dim - nsd_dim;  // MultiDiscrete([9, 9]) -> two one-hots
      v = (p.obs_position[2 * env + (q >= 9)] == (q >= 9 ? q - 9 : q)) ? 1.0f : 0.0f;
    } else {
      v = (float)p.obs_velocity[2 * env + (j - map_dim - nsd_dim - 18)];
    }
    out[i] = v;
  }
}

__global__ void pgtg_reduce_stats_kernel(const double* __restrict__ rows, int nrows, double* __restrict__ out) {
  // out[k] = sum over CTAs of rows[.][k]; one warp per statistic
  int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0;
  for (int r = lane; r < nrows; r += 32) s += rows[(size_t)r * STATS_STRIDE + k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[k] = s;
}

}  // namespace pgtg

// ---- CUDA backend: launch ----------------------------------------------------------------------
static int bk_pick_block(const pgtg::DevCfg& c, int* block, size_t* smem) {
  const char* forced = getenv("PGTG_BLOCK");  // experiment knob: CTA size of the tick kernel (32 / 64 / 128)
  int fb = forced ? atoi(forced) : 0;
  if (fb == 32 || fb == 64 || fb == 128) {
    size_t s = pgtg::block_shared_bytes(c, fb);
    if (s <= 200 * 1024) { *block = fb; *smem = s; return 0; }
  }
  for (int B : {128, 64, 32}) {
    size_t s = pgtg::block_shared_bytes(c, B);
    if (s <= 200 * 1024) { *block = B; *smem = s; return 0; }
  }
  return -1;
}

template <int RNG, int MODE, int TMAX, bool PREGEN, bool LEAN = false>
static int launch_one(pgtg_env* e, const uint8_t* mask, const int64_t* seeds, const void* actions, int action_bytes, cudaStream_t st) {
  auto kern = pgtg::pgtg_tick_kernel<RNG, MODE, TMAX, PREGEN, LEAN>;
  // same L1/shared carveout as the map-generation kernel: an SM cannot host CTAs of two kernels with
  // different carveouts, which would serialise the two (measured: no overlap at all without this)
  static bool carve_set = false;
  if (!carve_set) { cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); carve_set = true; }
  const size_t smem = LEAN ? pgtg::block_shared_bytes(e->dc, e->block, true) : e->smem;
  if (smem > 48 * 1024) {
    if (ck(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return -1;
  }
  pgtg::StatsArgs sa = {e->stats_rows};
  unsigned char* const origin = (unsigned char*)4096;  // any 16-byte-aligned address: only differences are used
  const pgtg::SharedLayout layout = pgtg::layout_of(pgtg::carve_shared(origin, e->dc, e->block, LEAN), origin);
  kern<<<e->nblk, e->block, smem, st>>>(e->dc, e->dp, mask, seeds, actions, action_bytes, sa, layout);
  return ck(cudaGetLastError());
}

// TMAX = compile-time bound on the tile count (register-resident boards for the default 4x4 map)
template <int RNG, int MODE, bool PREGEN>
static int launch_sized(pgtg_env* e, const uint8_t* mask, const int64_t* seeds, const void* actions, int action_bytes, cudaStream_t st) {
  if (e->dc.T <= 16) return launch_one<RNG, MODE, 16, PREGEN>(e, mask, seeds, actions, action_bytes, st);
  if (e->dc.T <= 64) return launch_one<RNG, MODE, 64, PREGEN>(e, mask, seeds, actions, action_bytes, st);
  return launch_one<RNG, MODE, 256, PREGEN>(e, mask, seeds, actions, action_bytes, st);
}

template <int RNG, int TMAX, bool TABLED = false>
static int launch_mapgen(pgtg_env* e, cudaStream_t st) {
  const int B = 128;
  size_t smem = pgtg::mapgen_shared_bytes(e->dc, B);
  auto kern = pgtg::pgtg_mapgen_kernel<RNG, TMAX, TABLED>;
  static bool carve_set = false;
  if (!carve_set) { cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); carve_set = true; }
  if (smem > 48 * 1024 && ck(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return -1;
  int full = (2 * e->dc.N + B - 1) / B;
  int grid = e->mapgen_grid > 0 && e->mapgen_grid < full ? e->mapgen_grid : full;
  unsigned char* const origin = (unsigned char*)4096;
  const pgtg::SharedLayout layout = pgtg::layout_of(pgtg::carve_mapgen(origin, e->dc, B), origin);
  kern<<<grid, B, smem, st>>>(e->dc, e->dp, e->dp.parity, layout);
  return ck(cudaGetLastError());
}

template <int RNG>
static int launch_mode(pgtg_env* e, int mode, const uint8_t* mask, const int64_t* seeds, const void* actions, int action_bytes, cudaStream_t st) {
  switch (mode) {
    case MODE_STEP:
      // plain configuration: the lean instantiation (the ring-fed reset does not depend on the board size)
      if (RNG != PGTG_RNG_TAPE && e->dc.pregen && e->dc.lean && !getenv("PGTG_NO_LEAN"))
        return launch_one<RNG == PGTG_RNG_TAPE ? PGTG_RNG_PHILOX : RNG, MODE_STEP, 16, RNG != PGTG_RNG_TAPE, RNG != PGTG_RNG_TAPE>(e, mask, seeds, actions, action_bytes, st);
      if (RNG != PGTG_RNG_TAPE && e->dc.pregen) return launch_sized<RNG, MODE_STEP, RNG != PGTG_RNG_TAPE>(e, mask, seeds, actions, action_bytes, st);
      return launch_sized<RNG, MODE_STEP, false>(e, mask, seeds, actions, action_bytes, st);
    case MODE_RESET:
      return launch_sized<RNG, MODE_RESET, false>(e, mask, seeds, actions, action_bytes, st);
    case MODE_MAPGEN:
      if (RNG == PGTG_RNG_TAPE) return -1;
      if (e->dc.conn_bits && e->dc.path_tab && !getenv("PGTG_NO_TABLED")) return launch_mapgen<RNG == PGTG_RNG_TAPE ? PGTG_RNG_PHILOX : RNG, 16, true>(e, st);
      if (e->dc.T <= 16) return launch_mapgen<RNG == PGTG_RNG_TAPE ? PGTG_RNG_PHILOX : RNG, 16>(e, st);
      if (e->dc.T <= 64) return launch_mapgen<RNG == PGTG_RNG_TAPE ? PGTG_RNG_PHILOX : RNG, 64>(e, st);
      return launch_mapgen<RNG == PGTG_RNG_TAPE ? PGTG_RNG_PHILOX : RNG, 256>(e, st);
    default:
      return launch_one<PGTG_RNG_PHILOX, MODE_OBSERVE, 16, false>(e, mask, seeds, actions, action_bytes, st);
  }
}

static int bk_launch(pgtg_env* e, int mode, const uint8_t* mask, const int64_t* seeds, const void* actions, int action_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (e->cfg.rng_mode) {
    case PGTG_RNG_TAPE: return launch_mode<PGTG_RNG_TAPE>(e, mode, mask, seeds, actions, action_bytes, st);
    case PGTG_RNG_NUMPY: return launch_mode<PGTG_RNG_NUMPY>(e, mode, mask, seeds, actions, action_bytes, st);
    default: return launch_mode<PGTG_RNG_PHILOX>(e, mode, mask, seeds, actions, action_bytes, st);
  }
}

extern "C" int pgtg_observe(pgtg_env* e, void* stream) {
  if (!e || !e->did_reset) return fail(PGTG_ERR_STATE, "observe before reset");
  bk_set_device(e->device);
  if (bk_launch(e, MODE_OBSERVE, nullptr, nullptr, nullptr, 0, stream)) return fail(PGTG_ERR_CUDA, std::string("observe launch failed: ") + bk_error());
  e->launches++;
  return PGTG_OK;
}

// Sum the per-CTA statistic rows into the 8-double `stats` buffer on the device (the buffer the
// host all-reduces with NCCL), on `stream`, without synchronising.
static int bk_stats_reduce(pgtg_env* e, void* stream) {
#ifdef PGTG_PHASE_CLOCKS
  {
    unsigned long long h[16];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, pgtg::g_phase_clk, sizeof(h));
    const char* names[9] = {"stage", "barrier", "step", "statistics, queue atomic", "reset", "emit", "queue write", "barrier", "expand"};
    double tot = 0;
    for (int i = 0; i < 9; i++) tot += (double)h[i];
    for (int i = 0; i < 9 && tot > 0; i++) fprintf(stderr, "[phase clocks] %-26s %6.2f %%\n", names[i], 100.0 * (double)h[i] / tot);
    memset(h, 0, sizeof(h));
    cudaMemcpyToSymbol(pgtg::g_phase_clk, h, sizeof(h));
  }
#endif
  pgtg::pgtg_reduce_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(e->stats_rows, e->nblk, e->dp.stats);
  e->launches++;
  return ck(cudaGetLastError());
}
static int bk_build_conn_table(pgtg_env* e, uint32_t* table_dev) {
  int s = e->dc.start_y * e->dc.W + e->dc.start_x, g = e->dc.goal_y * e->dc.W + e->dc.goal_x;
  pgtg::pgtg_build_conn_table_kernel<<<148 * 8, 256>>>(e->dc, table_dev, s, g);
  return ck(cudaGetLastError());
}
static int bk_build_path_table(pgtg_env* e, uint64_t* table_dev) {
  const int B = 128;
  size_t smem = pgtg::mapgen_shared_bytes(e->dc, B);
  uint32_t total = 1u << e->dc.conn_bits;
  int blocks = (int)((total + B - 1) / B < 148u * 12u ? (total + B - 1) / B : 148u * 12u);
  pgtg::pgtg_build_path_table_kernel<<<blocks, B, smem>>>(e->dc, e->dp, table_dev);
  return ck(cudaGetLastError());
}
static int bk_flatten(pgtg_env* e, void* stream) {
  pgtg::FlatOrder order;
  for (int i = 0; i < PGTG_MAX_CHANNELS; i++) order.plane[i] = e->flat_order[i];
  size_t total = (size_t)e->dc.N * e->flat_dim;
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  pgtg::pgtg_flatten_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(e->dc, e->dp, order, e->flat, e->flat_dim);
  return ck(cudaGetLastError());
}
static int bk_stats_reset(pgtg_env* e, void* stream) {
  if (ck(cudaMemsetAsync(e->stats_rows, 0, sizeof(double) * pgtg::STATS_STRIDE * (size_t)e->nblk, (cudaStream_t)stream))) return -1;
  return ck(cudaMemsetAsync(e->dp.stats, 0, 64, (cudaStream_t)stream));
}
