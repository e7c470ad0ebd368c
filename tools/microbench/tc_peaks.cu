// tc_peaks.cu -- measured per-SM ceilings of the two resources that bound the Hamming match kernel on sm_100a:
//   (1) tcgen05.mma.kind::i8 issue rate (M128 N128 K32 from shared memory, no epilogue)         -> dense int8 TOP/s
//   (2) tcgen05.ld (32x32b.x32) read bandwidth of TMEM (what the arg-max epilogue must sustain) -> bytes / clk / SM
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tc_peaks tools/microbench/tc_peaks.cu && /tmp/tc_peaks
// Prints one JSON line.  Not part of the product library.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W2;\nbra W1;\nW2:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) | (1ull << 46);
}
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);  // s32 += s8 x s8, M128 N128

__global__ void __launch_bounds__(160, 1) k_mma_rate(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 1);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 32) {
    const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem) + 32768);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const uint32_t d = tm + (uint32_t)((it & 3) * 128);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(da + (uint64_t)(k * 16)), "l"(db + (uint64_t)(k * 16)), "r"(kIdesc), "r"(k > 0 ? 1u : 0u)
                     : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

__global__ void __launch_bounds__(288, 1) k_tmem_read(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  if (warp >= 1) {  // 8 reader warps: lane quadrant = warp & 3, two warps per quadrant on different column halves
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp - 1) >> 2) * 256);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int c = 0; c < 8; c++) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
            "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(base + (uint32_t)(c * 32))
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= v[0] ^ v[31];
      }
    }
    const long long dt = clock64() - t0;
    if ((threadIdx.x & 31) == 0) {
      cycles[blockIdx.x * 8 + (warp - 1)] = dt;
      sink[blockIdx.x * 8 + (warp - 1)] = acc;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  long long* d_cyc;
  uint32_t* d_sink;
  cudaMalloc(&d_cyc, sizeof(long long) * sms * 8);
  cudaMalloc(&d_sink, 4 * sms * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  // ---- (1) int8 MMA rate, all SMs busy
  cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int it1 = 4000;
  k_mma_rate<<<sms, 160, 65536>>>(it1, d_cyc);
  cudaEventRecord(e0);
  k_mma_rate<<<sms, 160, 65536>>>(it1, d_cyc);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms1 = 0;
  cudaEventElapsedTime(&ms1, e0, e1);
  long long* h = new long long[sms * 8];
  cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double cyc1 = 0;
  for (int i = 0; i < sms; i++) cyc1 += (double)h[i] / sms;
  const double mma = (double)it1 * 8;
  const double ops1 = mma * 2.0 * 128 * 128 * 32 * sms;
  // ---- (2) TMEM read bandwidth, all SMs busy
  const int it2 = 20000;
  k_tmem_read<<<sms, 288>>>(it2, d_cyc, d_sink);
  cudaEventRecord(e0);
  k_tmem_read<<<sms, 288>>>(it2, d_cyc, d_sink);
  cudaEventRecord(e1);
  cudaError_t err2 = cudaDeviceSynchronize();
  float ms2 = 0;
  cudaEventElapsedTime(&ms2, e0, e1);
  cudaMemcpy(h, d_cyc, sizeof(long long) * sms * 8, cudaMemcpyDeviceToHost);
  double cyc2 = 0;
  for (int i = 0; i < sms * 8; i++) cyc2 += (double)h[i] / (sms * 8);
  const double bytes_sm = (double)it2 * 8 /*chunks*/ * 8 /*warps*/ * 32 * 32 * 4;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"errors\": \"%s / %s\", "
         "\"int8_mma\": {\"instr\": \"tcgen05.mma.cta_group::1.kind::i8 M128 N128 K32, SS\", \"cycles_per_mma\": %.2f, \"ms\": %.4f, \"dense_TOPs\": %.1f}, "
         "\"tmem_read\": {\"instr\": \"tcgen05.ld.32x32b.x32, 8 warps / SM\", \"bytes_per_clk_per_sm\": %.2f, \"ms\": %.4f, \"TBps_chip\": %.2f}}\n",
         p.name, sms, cudaGetErrorString(err), cudaGetErrorString(err2), cyc1 / mma, ms1, ops1 / (ms1 * 1e-3) / 1e12, bytes_sm / cyc2, ms2,
         bytes_sm * sms / (ms2 * 1e-3) / 1e12);
  return 0;
}
