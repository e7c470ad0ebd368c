// tc_ts_probe.cu -- what bounds tcgen05.mma.kind::i8 when the operands share the shared-memory pipe with other traffic, and
// whether taking A from tensor memory relieves it.  Two parts:
//   (1) correctness of the TMEM-A form: D = A B^T with A written by tcgen05.st.32x32b (lane = row, one 32-bit column = 4
//       consecutive k bytes), K = 32 and K = 64 (second k-step = A columns + 8), checked against the CPU;
//   (2) cycles per MMA for {A from smem, A from TMEM} x N in {64, 128, 256}, alone and with 2 / 4 / 8 warps storing 16 B per
//       lane to shared memory in a loop (what the operand-expanding producers of the match kernel do);
//   (3) the same for N = 128 from shared memory with what else the match kernel does around its MMAs: a fence.proxy.async
//       after every 16 stores (the producers publish a tile row), 8 warps draining the other accumulator with tcgen05.ld.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tc_ts_probe tools/microbench/tc_ts_probe.cu && /tmp/tc_ts_probe
// Prints one JSON object.  Not part of the product library.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W2;\nbra W1;\nW2:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {  // K-major, no swizzle: LBO 128 B (k-chunks), SBO 2048 B (8-row groups)
  return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_i8(uint32_t n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}

__host__ __device__ inline int a_val(int r, int k) { return ((r * 7 + k * 3) % 5) - 2; }
__host__ __device__ inline int b_val(int n, int k) { return ((n * 5 + k) % 7) - 3; }

// ---- (1) correctness ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_check(int ksteps, int32_t* out /* 128 x 64 */) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 256; i += blockDim.x) {
    const int n = i >> 8, k = i & 255;
    smem[(n >> 3) * 2048 + (k >> 4) * 128 + (n & 7) * 16 + (k & 15)] = (uint8_t)(int8_t)(k < 32 * ksteps ? b_val(n, k) : 0);
  }
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
  const int row = threadIdx.x;
  for (int ks = 0; ks < ksteps; ks++) {  // A rows -> TMEM columns 256 + 8 ks .. : lane = row, column c = k bytes 4c .. 4c + 3
    uint32_t w[8];
    for (int c = 0; c < 8; c++) {
      uint32_t v = 0;
      for (int b = 0; b < 4; b++) v |= (uint32_t)(uint8_t)(int8_t)a_val(row, 32 * ks + 4 * c + b) << (8 * b);
      w[c] = v;
    }
    const uint32_t ta = tm + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)(8 * ks);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint64_t db = make_desc(smem_u32(smem));
    for (int ks = 0; ks < ksteps; ks++) mma_ts(tm, tm + 256u + (uint32_t)(8 * ks), db + (uint64_t)(ks * 16), idesc_i8(64), ks > 0 ? 1u : 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    const uint32_t ta = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(ta)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 8; c++) out[row * 64 + c0 + c] = (int32_t)v[c];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  (void)lane;
}

// ---- (2) rate under shared-memory contention ------------------------------------------------------------------------------
// warp 0: TMEM allocation; thread 32: MMA issuer; warps 2 .. 2 + n_store: 16-byte stores to a scratch region of shared memory
template <int TS, int N>
__global__ void __launch_bounds__(320, 1) k_rate(int iters, int n_store, long long* cycles, uint32_t* sink, int fence_every = 0,
                                                 int ld_warps = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];  // A tile 32 KiB | B tile up to 64 KiB | scratch 32 KiB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 1);
  if (threadIdx.x == 0) { stop = 0; mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
    const uint32_t a_t = tm + 256u + (N == 256 ? 0u : 0u);  // A columns (garbage values: only the rate matters)
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const uint32_t d = N == 256 ? tm : tm + (uint32_t)((it & 1) * 128);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (TS) mma_ts(d, (N == 256 ? tm + 384u : a_t) + (uint32_t)(8 * k), db + (uint64_t)(k * 16), idesc_i8(N), k > 0 ? 1u : 0u);
        else mma_ss(d, da + (uint64_t)(k * 16), db + (uint64_t)(k * 16), idesc_i8(N), k > 0 ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x] = clock64() - t0;
    stop = 1;
  } else if (ld_warps > 0 && warp >= 2 && warp < 2 + ld_warps) {
    // drain the accumulator the MMAs are NOT writing?  here simply columns 256..383 (never written): lane quadrant = warp & 3
    const uint32_t t0 = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)(((warp - 2) >> 2) * 64);
    uint32_t acc = 0, n = 0;
    while (!stop) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(t0 + (n & 1u) * 32u)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= v[0] ^ v[31];
      n++;
    }
    if (lane == 0) sink[blockIdx.x * 8 + (warp - 2)] = n + (acc & 0u);
  } else if (ld_warps == 0 && warp >= 2 && warp < 2 + n_store) {
    const uint32_t base = smem_u32(smem) + 32768u + 65536u + (uint32_t)((warp - 2) * 4096) + (uint32_t)lane * 16u;
    uint32_t x = threadIdx.x, n = 0;
    while (!stop) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + (uint32_t)(u * 512)), "r"(x) : "memory");
        x = x * 1664525u + 1013904223u;
      }
      n += 8;
      if (fence_every > 0 && (n % (uint32_t)fence_every) == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (lane == 0) sink[blockIdx.x * 8 + (warp - 2)] = n;  // 16-byte stores per lane issued while the MMAs ran
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// ---- (4) how far the issuing thread may run ahead of the tensor pipe: bursts of 9 MMAs (one accumulator of the match kernel)
// separated by `delay` cycles in which the thread issues nothing.  While delay < (queued work) the pipe never starves and a
// burst costs 9 x 64 cycles; the knee gives the depth of the MMA queue.
__global__ void __launch_bounds__(64, 1) k_queue(int bursts, int delay, long long* cycles) {
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
    for (int it = 0; it < bursts; it++) {
      const uint32_t d = tm + (uint32_t)((it & 3) * 128);
#pragma unroll
      for (int k = 0; k < 9; k++) mma_ss(d, da + (uint64_t)((k & 7) * 16), db + (uint64_t)((k & 7) * 16), idesc_i8(128), k > 0 ? 1u : 0u);
      const long long t1 = clock64();
      while (clock64() - t1 < delay) {}
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int TS, int N>
static void run_rate(int sms, long long* d_cyc, uint32_t* d_sink, bool first) {
  const int iters = 2000;
  const size_t smem = 32768 + 65536 + 32768;
  cudaFuncSetAttribute(k_rate<TS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int ns : {0, 2, 4, 8}) {
    cudaMemset(d_sink, 0, sizeof(uint32_t) * sms * 8);
    k_rate<TS, N><<<sms, 320, smem>>>(iters, ns, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    std::vector<uint32_t> hs(sms * 8);
    cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs.data(), d_sink, sizeof(uint32_t) * sms * 8, cudaMemcpyDeviceToHost);
    double cyc = 0, st = 0;
    for (int i = 0; i < sms; i++) cyc += (double)h[i] / sms;
    for (int i = 0; i < sms * 8; i++) st += (double)hs[i] / sms;
    // store bytes per clk per SM = 16-byte stores per lane x 32 lanes x 16 B / cycles
    printf("%s{\"a_from\": \"%s\", \"N\": %d, \"store_warps\": %d, \"cycles_per_mma\": %.2f, \"ideal\": %d, \"store_bytes_per_clk\": %.1f, \"err\": \"%s\"}",
           (first && ns == 0) ? "" : ", ", TS ? "tmem" : "smem", N, ns, cyc / (iters * 8.0), N / 2, st * 512.0 / cyc, cudaGetErrorString(e));
  }
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  long long* d_cyc;
  uint32_t* d_sink;
  int32_t* d_out;
  cudaMalloc(&d_cyc, sizeof(long long) * sms);
  cudaMalloc(&d_sink, sizeof(uint32_t) * sms * 8);
  cudaMalloc(&d_out, sizeof(int32_t) * 128 * 64);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"check\": [", p.name, sms);
  cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int ks = 1; ks <= 2; ks++) {
    cudaMemset(d_out, 0xff, sizeof(int32_t) * 128 * 64);
    k_check<<<1, 128, 16384>>>(ks, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int32_t> h(128 * 64);
    cudaMemcpy(h.data(), d_out, sizeof(int32_t) * 128 * 64, cudaMemcpyDeviceToHost);
    int bad = 0, first_bad = -1;
    for (int r = 0; r < 128; r++)
      for (int n = 0; n < 64; n++) {
        int ref = 0;
        for (int k = 0; k < 32 * ks; k++) ref += a_val(r, k) * b_val(n, k);
        if (h[r * 64 + n] != ref) {
          if (first_bad < 0) first_bad = r * 64 + n;
          bad++;
        }
      }
    printf("%s{\"k_steps\": %d, \"mismatches\": %d, \"first_bad\": %d, \"got0\": %d, \"err\": \"%s\"}", ks > 1 ? ", " : "", ks, bad, first_bad, h[0],
           cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("]}\n"); return 1; }
  }
  printf("], \"rate\": [");
  run_rate<0, 128>(sms, d_cyc, d_sink, true);
  run_rate<0, 256>(sms, d_cyc, d_sink, false);
  run_rate<0, 64>(sms, d_cyc, d_sink, false);
  run_rate<1, 64>(sms, d_cyc, d_sink, false);
  run_rate<1, 128>(sms, d_cyc, d_sink, false);
  run_rate<1, 256>(sms, d_cyc, d_sink, false);
  printf("], \"around_the_mma\": [");
  {
    const int iters = 2000;
    const size_t smem = 32768 + 65536 + 32768;
    struct Cfg { const char* what; int ns, fence, ld; } cfgs[] = {
        {"8 store warps, fence.proxy.async every 16 stores", 8, 16, 0}, {"8 store warps, fence every 64 stores", 8, 64, 0},
        {"2 store warps, fence every 16 stores", 2, 16, 0},             {"8 warps of tcgen05.ld.x32 + wait", 0, 0, 8},
        {"4 warps of tcgen05.ld.x32 + wait", 0, 0, 4}};
    bool first = true;
    for (const Cfg& c : cfgs) {
      cudaMemset(d_sink, 0, sizeof(uint32_t) * sms * 8);
      k_rate<0, 128><<<sms, 320, smem>>>(iters, c.ns, d_cyc, d_sink, c.fence, c.ld);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<long long> h(sms);
      std::vector<uint32_t> hs(sms * 8);
      cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      cudaMemcpy(hs.data(), d_sink, sizeof(uint32_t) * sms * 8, cudaMemcpyDeviceToHost);
      double cyc = 0, st = 0;
      for (int i = 0; i < sms; i++) cyc += (double)h[i] / sms;
      for (int i = 0; i < sms * 8; i++) st += (double)hs[i] / sms;
      printf("%s{\"what\": \"%s\", \"cycles_per_mma\": %.2f, \"side_ops_per_1000_clk\": %.1f, \"err\": \"%s\"}", first ? "" : ", ", c.what,
             cyc / (iters * 8.0), st * 1000.0 / cyc, cudaGetErrorString(e));
      first = false;
    }
  }
  printf("], \"issue_queue\": [");
  {
    cudaFuncSetAttribute(k_queue, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    bool first = true;
    for (int delay : {0, 64, 128, 192, 256, 320, 384, 448, 512, 576, 640, 768}) {
      const int bursts = 1000;
      k_queue<<<sms, 64, 65536>>>(bursts, delay, d_cyc);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<long long> h(sms);
      cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      double cyc = 0;
      for (int i = 0; i < sms; i++) cyc += (double)h[i] / sms;
      printf("%s{\"idle_cycles_after_burst\": %d, \"cycles_per_9_mma_burst\": %.1f, \"err\": \"%s\"}", first ? "" : ", ", delay, cyc / bursts,
             cudaGetErrorString(e));
      first = false;
    }
  }
  printf("]}\n");
  return 0;
}
