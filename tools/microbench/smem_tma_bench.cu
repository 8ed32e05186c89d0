// Micro-benchmarks behind the tile-stream kernels' design choices (B200, sm_100a):
//   1. shared-memory scatter-add: native integer ATOMS.ADD vs the float compare-and-swap loop vs plain LDS/STS,
//      random slots in an 8 KB ring, at the occupancies the lattice kernels run at;
//   2. per-warp TMA bulk-copy pipelines (cp.async.bulk + mbarrier): chip-wide GB/s as a function of copy size,
//      pipeline depth and warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_tma_bench smem_tma_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

// ---------------- 1. scatter-add variants ----------------
// every warp: iters rounds of 8 independent updates (like 8 arc columns) at pseudo-random slots
template <int MODE>  // 0 = ATOMS.ADD u32 (F2I first), 1 = float atomicAdd (CAS loop), 2 = LDS + STS (no atomic), 3 = LDS gather only
__global__ void scatter_kernel(int iters, int W, float* out) {
  extern __shared__ unsigned char sm[];
  unsigned* ru = reinterpret_cast<unsigned*>(sm);
  float* rf = reinterpret_cast<float*>(sm);
  for (int i = threadIdx.x; i < W; i += blockDim.x) ru[i] = 0;
  __syncthreads();
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  const float v = 1e-3f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x = x * 1664525u + 1013904223u;
      const int slot = (x >> 8) % W;
      if (MODE == 0) atomicAdd(&ru[slot], __float2uint_rn(v * 2147483648.0f * (1.0f / 1024.f)));
      else if (MODE == 1) atomicAdd(&rf[slot], v);
      else if (MODE == 2) { float o = rf[slot]; rf[slot] = o + v; }
      else acc += rf[slot];
    }
  }
  __syncthreads();
  if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = acc + rf[threadIdx.x % W];
}

// ---------------- 2. per-warp TMA bulk pipelines ----------------
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// each warp streams its own contiguous region of `src` in copies of `bytes`, DEPTH copies in flight, and reads
// one word per lane from each landed stage (so that the data is actually consumed)
template <int DEPTH>
__global__ void tma_stream_kernel(const unsigned char* __restrict__ src, size_t per_warp_bytes, int bytes, int n_arrays, float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t gw = static_cast<size_t>(blockIdx.x) * nw + warp;
  unsigned char* stage0 = sm + static_cast<size_t>(warp) * DEPTH * bytes * n_arrays;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + static_cast<size_t>(nw) * DEPTH * bytes * n_arrays) + warp * DEPTH;
  if (lane == 0)
    for (int d = 0; d < DEPTH; ++d) mbar_init(smem_u32(bars + d), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const unsigned char* base = src + gw * per_warp_bytes * n_arrays;
  const int n = static_cast<int>(per_warp_bytes / bytes);
  auto issue = [&](int i) {
    const int d = i % DEPTH;
    const unsigned bar = smem_u32(bars + d);
    mbar_expect_tx(bar, bytes * n_arrays);
    for (int a = 0; a < n_arrays; ++a)
      bulk_g2s(smem_u32(stage0 + (d * n_arrays + a) * bytes), base + a * per_warp_bytes + static_cast<size_t>(i) * bytes, bytes, bar);
  };
  if (lane == 0)
    for (int i = 0; i < DEPTH && i < n; ++i) issue(i);
  float acc = 0.f;
  for (int i = 0; i < n; ++i) {
    const int d = i % DEPTH;
    mbar_wait(smem_u32(bars + d), (i / DEPTH) & 1);
    for (int a = 0; a < n_arrays; ++a) acc += reinterpret_cast<const float*>(stage0 + (d * n_arrays + a) * bytes)[lane];
    __syncwarp();
    if (lane == 0 && i + DEPTH < n) issue(i + DEPTH);
  }
  if (out) out[gw * 32 + lane] = acc;
}

// the same stream with plain coalesced loads (LDG.128 per lane, UNROLL loads in flight)
__global__ void ldg_stream_kernel(const float4* __restrict__ src, size_t per_warp_vec, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t gw = static_cast<size_t>(blockIdx.x) * nw + warp;
  const float4* p = src + gw * per_warp_vec;
  float acc = 0.f;
  for (size_t i = lane; i + 32 * 7 < per_warp_vec; i += 32 * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = p[i + 32 * u];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (out) out[gw * 32 + lane] = acc;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const double ghz = prop.clockRate * 1e-6;
  printf("device %s, %d SMs, %.3f GHz\n", prop.name, sms, ghz);
  float* out; CK(cudaMalloc(&out, sizeof(float) * 148 * 32 * 1024 * 4));

  // ---- 1. scatter-add ----
  const char* names[4] = {"ATOMS.ADD.u32 (F2I)", "float atomicAdd (CAS loop)", "LDS+STS non-atomic", "LDS gather only"};
  for (int W : {2048, 16384}) {
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int threads = cfg == 0 ? 32 : cfg == 1 ? 64 : 128;
      const int blocks = sms * 7;
      const int iters = 2000;
      for (int mode = 0; mode < 4; ++mode) {
        auto run = [&]() {
          const size_t smem = W * 4;
          if (mode == 0) scatter_kernel<0><<<blocks, threads, smem>>>(iters, W, out);
          if (mode == 1) scatter_kernel<1><<<blocks, threads, smem>>>(iters, W, out);
          if (mode == 2) scatter_kernel<2><<<blocks, threads, smem>>>(iters, W, out);
          if (mode == 3) scatter_kernel<3><<<blocks, threads, smem>>>(iters, W, out);
        };
        if (W * 4 > 48 * 1024) continue;
        const float ms = time_ms(run);
        const double ops = static_cast<double>(blocks) * threads * iters * 8;
        printf("scatter W=%5d threads/block=%3d (7 blocks/SM) %-28s %8.3f ms  %7.2f G lane-updates/s  %.3f cyc/lane-update/SM\n", W,
               threads, names[mode], ms, ops / ms * 1e-6, ms * 1e-3 * ghz * 1e9 * sms / ops);
      }
    }
  }

  // ---- 2. TMA bulk pipelines ----
  const size_t total = static_cast<size_t>(3) << 30;  // 3 GiB source (>> L2)
  unsigned char* src; CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  for (int warps_per_block : {1, 2, 4}) {
    for (int bytes : {512, 1024, 2048, 4096}) {
      for (int depth : {2, 3, 4}) {
        for (int n_arrays : {1, 2}) {
          const int blocks = sms * 7;
          const size_t nwarps = static_cast<size_t>(blocks) * warps_per_block;
          size_t per_warp = (total / nwarps / n_arrays) / bytes * bytes;
          if (per_warp > (static_cast<size_t>(1) << 21)) per_warp = static_cast<size_t>(1) << 21;  // <= 2 MiB per warp per array
          const size_t smem = static_cast<size_t>(warps_per_block) * depth * bytes * n_arrays + warps_per_block * depth * 8;
          if (smem > 32 * 1024) continue;
          auto run = [&]() {
            if (depth == 2) tma_stream_kernel<2><<<blocks, warps_per_block * 32, smem>>>(src, per_warp, bytes, n_arrays, out);
            if (depth == 3) tma_stream_kernel<3><<<blocks, warps_per_block * 32, smem>>>(src, per_warp, bytes, n_arrays, out);
            if (depth == 4) tma_stream_kernel<4><<<blocks, warps_per_block * 32, smem>>>(src, per_warp, bytes, n_arrays, out);
          };
          const float ms = time_ms(run, 3);
          const double gb = static_cast<double>(per_warp) * n_arrays * nwarps * 1e-9;
          printf("tma  warps/block=%d (7 blocks/SM) copy=%4d B x%d arrays depth=%d smem/block=%5zu  %8.3f ms  %8.1f GB/s\n", warps_per_block,
                 bytes, n_arrays, depth, smem, ms, gb / (ms * 1e-3));
        }
      }
    }
  }
  for (int warps_per_block : {1, 2, 4, 8}) {
    const int blocks = sms * 7;
    const size_t nwarps = static_cast<size_t>(blocks) * warps_per_block;
    size_t per_warp_vec = total / nwarps / 16;
    if (per_warp_vec > (1u << 17)) per_warp_vec = 1u << 17;
    per_warp_vec = per_warp_vec / 256 * 256;
    auto run = [&]() { ldg_stream_kernel<<<blocks, warps_per_block * 32>>>(reinterpret_cast<const float4*>(src), per_warp_vec, out); };
    const float ms = time_ms(run, 3);
    printf("ldg  warps/block=%d (7 blocks/SM) 8 x LDG.128 in flight per lane  %8.3f ms  %8.1f GB/s\n", warps_per_block, ms,
           static_cast<double>(per_warp_vec) * 16 * nwarps * 1e-9 / (ms * 1e-3));
  }
  return 0;
}
