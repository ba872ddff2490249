// Scratch microbenchmark: how fast can the observation write pattern of the tick kernel go on its own?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart static -o tools/scratch/write_pattern tools/scratch/write_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_cs(uint4* p, uint4 v) { asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void st_def(uint4* p, uint4 v) { *p = v; }

template <int MODE>  // 0: default store, 1: .cs
__global__ void fill_stride(uint4* out, size_t n16) {
  uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) { if (MODE) st_cs(out + i, v); else st_def(out + i, v); }
}

// one CTA = one contiguous chunk (like a batch of 128 envs x 729 bytes)
template <int MODE, int LUT>
__global__ void __launch_bounds__(128, 8) fill_chunks(uint4* out, int chunk16) {
  extern __shared__ uint32_t sm[];
  uint2* lut = (uint2*)sm;               // 256 entries
  uint16_t* half = (uint16_t*)(sm + 512);  // chunk16 entries
  if (LUT) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = make_uint2(i * 0x01010101u & 0x01010101u, i);
    for (int i = threadIdx.x; i < chunk16; i += blockDim.x) half[i] = (uint16_t)(i * 2654435761u >> 20);
    __syncthreads();
  }
  uint4* o = out + (size_t)blockIdx.x * chunk16;
  const int B = blockDim.x;
  int j = threadIdx.x;
  for (; j + 3 * B < chunk16; j += 4 * B) {
    uint4 v0, v1, v2, v3;
    if (LUT) {
      uint32_t h0 = half[j], h1 = half[j + B], h2 = half[j + 2 * B], h3 = half[j + 3 * B];
      uint2 a, b;
      a = lut[h0 & 255]; b = lut[h0 >> 8]; v0 = make_uint4(a.x, a.y, b.x, b.y);
      a = lut[h1 & 255]; b = lut[h1 >> 8]; v1 = make_uint4(a.x, a.y, b.x, b.y);
      a = lut[h2 & 255]; b = lut[h2 >> 8]; v2 = make_uint4(a.x, a.y, b.x, b.y);
      a = lut[h3 & 255]; b = lut[h3 >> 8]; v3 = make_uint4(a.x, a.y, b.x, b.y);
    } else { v0 = v1 = v2 = v3 = make_uint4(j, 1, 2, 3); }
    if (MODE) { st_cs(o + j, v0); st_cs(o + j + B, v1); st_cs(o + j + 2 * B, v2); st_cs(o + j + 3 * B, v3); }
    else { st_def(o + j, v0); st_def(o + j + B, v1); st_def(o + j + 2 * B, v2); st_def(o + j + 3 * B, v3); }
  }
  for (; j < chunk16; j += B) { uint4 v = make_uint4(j, 1, 2, 3); if (MODE) st_cs(o + j, v); else st_def(o + j, v); }
}

// chunk kernel + a latency phase in front (dependent global loads + ALU chain), to mimic step/emit
template <int CHAIN>
__global__ void __launch_bounds__(128, 8) chunks_with_work(uint4* out, int chunk16, const uint32_t* __restrict__ state, uint32_t* sink) {
  extern __shared__ uint32_t sm[];
  uint2* lut = (uint2*)sm;
  uint16_t* half = (uint16_t*)(sm + 512);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = make_uint2(i * 0x01010101u & 0x01010101u, i);
  size_t env = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t x = state[env];
  for (int i = 0; i < CHAIN; i++) x = x * 1664525u + 1013904223u + (x >> 7);  // dependent ALU chain
  for (int i = threadIdx.x; i < chunk16; i += blockDim.x) half[i] = (uint16_t)((i + x) * 2654435761u >> 20);
  __syncthreads();
  uint4* o = out + (size_t)blockIdx.x * chunk16;
  const int B = blockDim.x;
  int j = threadIdx.x;
  for (; j + 3 * B < chunk16; j += 4 * B) {
    uint32_t h0 = half[j], h1 = half[j + B], h2 = half[j + 2 * B], h3 = half[j + 3 * B];
    uint2 a, b; uint4 v0, v1, v2, v3;
    a = lut[h0 & 255]; b = lut[h0 >> 8]; v0 = make_uint4(a.x, a.y, b.x, b.y);
    a = lut[h1 & 255]; b = lut[h1 >> 8]; v1 = make_uint4(a.x, a.y, b.x, b.y);
    a = lut[h2 & 255]; b = lut[h2 >> 8]; v2 = make_uint4(a.x, a.y, b.x, b.y);
    a = lut[h3 & 255]; b = lut[h3 >> 8]; v3 = make_uint4(a.x, a.y, b.x, b.y);
    st_cs(o + j, v0); st_cs(o + j + B, v1); st_cs(o + j + 2 * B, v2); st_cs(o + j + 3 * B, v3);
  }
  for (; j < chunk16; j += B) { uint4 v = make_uint4(j, 1, 2, 3); st_cs(o + j, v); }
  if (x == 12345u) sink[0] = x;
}

template <class F> float timeit(F f, int reps = 20) {
  for (int i = 0; i < 3; i++) f();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const int N = 2097152, B = 128, chunk16 = B * 729 / 16;  // 5832
  const size_t bytes = (size_t)N * 729, n16 = bytes / 16;
  uint4* out; cudaMalloc(&out, bytes);
  uint32_t* state; cudaMalloc(&state, (size_t)N * 4); cudaMemset(state, 1, (size_t)N * 4);
  uint32_t* sink; cudaMalloc(&sink, 4);
  const int nblk = N / B;
  size_t smem = 2048 + chunk16 * 2 + 64;
  size_t smem27 = 27136;
  cudaFuncSetAttribute(fill_chunks<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  float ms;
  auto rep = [&](const char* name, float ms) { printf("%-46s %.3f ms  %.0f GB/s\n", name, ms, bytes / ms / 1e6); };
  ms = timeit([&] { fill_stride<0><<<148 * 16, 256>>>(out, n16); }); rep("grid-stride fill, default store", ms);
  ms = timeit([&] { fill_stride<1><<<148 * 16, 256>>>(out, n16); }); rep("grid-stride fill, st.cs", ms);
  ms = timeit([&] { fill_chunks<0, 0><<<nblk, B, smem>>>(out, chunk16); }); rep("CTA chunks (93 KB), default store", ms);
  ms = timeit([&] { fill_chunks<1, 0><<<nblk, B, smem>>>(out, chunk16); }); rep("CTA chunks, st.cs", ms);
  ms = timeit([&] { fill_chunks<1, 1><<<nblk, B, smem>>>(out, chunk16); }); rep("CTA chunks, st.cs, smem LUT expand", ms);
  ms = timeit([&] { fill_chunks<1, 1><<<nblk, B, smem27>>>(out, chunk16); }); rep("  same with 27 KB smem/CTA (8 CTAs/SM)", ms);
  ms = timeit([&] { chunks_with_work<0><<<nblk, B, smem27>>>(out, chunk16, state, sink); }); rep("  + state load, chain 0", ms);
  ms = timeit([&] { chunks_with_work<500><<<nblk, B, smem27>>>(out, chunk16, state, sink); }); rep("  + state load, chain 500 (x3 instr)", ms);
  ms = timeit([&] { chunks_with_work<1000><<<nblk, B, smem27>>>(out, chunk16, state, sink); }); rep("  + state load, chain 1000", ms);
  ms = timeit([&] { chunks_with_work<2000><<<nblk, B, smem27>>>(out, chunk16, state, sink); }); rep("  + state load, chain 2000", ms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
