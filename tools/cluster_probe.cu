// Micro-benchmarks that size the round-2 "code predictor on one cluster" design.  Build and run
// on a B200:   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/cluster_probe tools/cluster_probe.cu && tools/cluster_probe
// Reports: (1) L2->SM read bandwidth of 1 / 2 / 4 / 8 / 16 / 148 CTAs streaming an L2-resident buffer with the same
// 16-byte ld.global.nc pattern the GEMV phases use; (2) latency of a 16-CTA cluster barrier (barrier.cluster
// arrive.release + wait.acquire) vs the software grid barrier over 16 and 148 CTAs; (3) round-trip latency of a
// distributed-shared-memory load from a peer CTA.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint4 ldnc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// every CTA streams `bytes_per_cta` starting at its own offset, `reps` times (first rep warms L2)
__global__ void __launch_bounds__(256) stream_kernel(const char* buf, size_t bytes_per_cta, int reps, unsigned* sink, long long* cyc) {
  const char* base = buf + (size_t)blockIdx.x * bytes_per_cta;
  unsigned acc = 0;
  long long t0 = 0;
  for (int r = 0; r < reps; ++r) {
    if (r == 1) { __syncthreads(); t0 = clock64(); }
    for (size_t off = (size_t)threadIdx.x * 16; off < bytes_per_cta; off += 256 * 16 * 8) {
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const size_t o = off + (size_t)i * 256 * 16;
        v[i] = o < bytes_per_cta ? ldnc16(base + o) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

__global__ void __cluster_dims__(16, 1, 1) __launch_bounds__(256) cluster_barrier_kernel(int iters, long long* out) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (clock64() - t0) / iters;
}

__global__ void __cluster_dims__(16, 1, 1) __launch_bounds__(256) dsmem_kernel(int iters, long long* out, unsigned* sink) {
  __shared__ unsigned cell[64];
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x < 64) cell[threadIdx.x] = threadIdx.x * 7u + blockIdx.x;
  cl.sync();
  unsigned* peer = cl.map_shared_rank(cell, (cl.block_rank() + 1) % cl.num_blocks());
  unsigned idx = threadIdx.x & 63, acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { idx = peer[idx] & 63; acc += idx; }   // dependent chain of remote loads
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
  cl.sync();
}

__global__ void __launch_bounds__(256) grid_barrier_kernel(unsigned* ctr, int iters, long long* out) {
  unsigned epoch = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      unsigned v;
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while ((int)(v - epoch) < 0);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (clock64() - t0) / iters;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const double ghz = prop.clockRate / 1e6;
  printf("%s, %d SMs, %.3f GHz nominal\n", prop.name, prop.multiProcessorCount, ghz);
  const size_t total = (size_t)64 << 20;  // 64 MB: fits the 126 MB L2
  char* buf; cudaMalloc(&buf, total); cudaMemset(buf, 1, total);
  unsigned* sink; cudaMalloc(&sink, 148 * 256 * 4 * 2);
  long long* out; cudaMallocManaged(&out, 256 * 8);
  const int grids[] = {1, 2, 4, 8, 16, 148};
  for (int g : grids) {
    const size_t per = (total / g) & ~(size_t)(256 * 16 * 8 - 1);
    stream_kernel<<<g, 256>>>(buf, per, 4, sink, out);
    cudaDeviceSynchronize();
    long long mx = 0;
    for (int i = 0; i < g; ++i) mx = out[i] > mx ? out[i] : mx;
    const double sec = mx / (ghz * 1e9);
    printf("L2-resident stream, %3d CTAs x %6.2f MB x 3 reps: %8.1f GB/s total, %7.1f GB/s per CTA\n", g, per / 1048576.0,
           3.0 * per * g / sec / 1e9, 3.0 * per / sec / 1e9);
  }
  // 16-CTA clusters exceed the portable limit of 8: opt in
  cudaFuncSetAttribute(cluster_barrier_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cluster_barrier_kernel<<<16, 256>>>(2000, out); cudaDeviceSynchronize();
  printf("cluster barrier (16 CTAs, arrive.release + wait.acquire): %lld cycles = %.0f ns  [%s]\n", out[0], out[0] / ghz,
         cudaGetErrorString(cudaGetLastError()));
  dsmem_kernel<<<16, 256>>>(2000, out, sink); cudaDeviceSynchronize();
  printf("DSMEM dependent remote load: %lld cycles = %.0f ns  [%s]\n", out[0], out[0] / ghz, cudaGetErrorString(cudaGetLastError()));
  unsigned* ctr; cudaMalloc(&ctr, 4);
  for (int g : {16, 148}) {
    cudaMemset(ctr, 0, 4);
    void* args[] = {&ctr, nullptr, &out};
    int iters = 2000; args[1] = &iters;
    cudaLaunchCooperativeKernel((const void*)grid_barrier_kernel, dim3(g), dim3(256), args, 0, 0);
    cudaDeviceSynchronize();
    printf("software grid barrier, %3d CTAs: %lld cycles = %.0f ns  [%s]\n", g, out[0], out[0] / ghz, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
