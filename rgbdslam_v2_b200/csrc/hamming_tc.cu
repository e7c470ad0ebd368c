// hamming_tc.cu -- Hamming N x M brute-force match on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// bruteForceSearchORB (features.cpp:168-182) for all query rows of all pairs, exact:
//   256-bit descriptors are expanded once per node to +-1 int8 vectors (bit set -> +1, clear -> -1).
//   For two descriptors a, b:  a . b = 256 - 2 * hamming(a, b)   (exact in int32)
//   => argmin_j hd(q_i, t_j)  ==  argmax_j (Q T^T)_ij, ties -> lowest j  (features.cpp:176 strict <).
// The N x M distance matrix is therefore an int8 GEMM with a row-arg-max epilogue:
//   tcgen05.mma.kind::i8  M=128 (queries -> TMEM lanes)  N=256 (train rows -> TMEM columns)  K=32 x 8 steps,
//   int32 accumulators in TMEM, double buffered (2 x 256 columns) so the epilogue of tile n overlaps the
//   MMAs of tile n+1.
//
// Data movement: the expansion kernel writes each node's int8 matrix already in the UMMA canonical
// K-major no-swizzle shared-memory layout, in 128-row tiles of 32 KiB:
//     tile[row_group 16][k_chunk 16][row_in_group 8][16 B]
// so a whole operand tile is ONE contiguous global block and is staged with a single
// cp.async.bulk (TMA bulk copy, completes on an mbarrier).  No tensor map / swizzle is needed and the
// MMA reads conflict-free 8x16 B core matrices.
//
// Warp roles (192 threads, 1 CTA / SM, persistent over work items = (pair, 128-query tile)):
//   warp 0 lane 0 : bulk-copy producer (A tile per item, B tiles of 256 train rows, 2-stage rings)
//   warp 1 lane 0 : tcgen05.mma issuer
//   warps 2..5    : epilogue -- tcgen05.ld of the accumulator, running arg-max per query row, final store
#include "kernels.h"

namespace rb200 {

// ---------------------------------------------------------------------------------------------
// +-1 int8 expansion into the tiled layout described above.  One thread per 16-byte chunk.
__global__ void __launch_bounds__(256) expand_i8_kernel(const ExpandJob* __restrict__ jobs) {
  const ExpandJob job = jobs[blockIdx.y];
  const int c = blockIdx.x * 256 + threadIdx.x;  // chunk index in output order
  if (c >= job.n_pad * 16) return;
  const int tile = c >> 11;          // 2048 chunks per 128-row tile
  const int rg = (c >> 7) & 15;      // row group
  const int kc = (c >> 3) & 15;      // 16-byte K chunk
  const int rr = c & 7;              // row in group
  const int row = tile * 128 + rg * 8 + rr;
  uint4 out = make_uint4(0, 0, 0, 0);
  if (row < job.n) {
    const unsigned bits = (unsigned)job.desc[(size_t)row * 32 + kc * 2] | ((unsigned)job.desc[(size_t)row * 32 + kc * 2 + 1] << 8);
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      unsigned v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) v |= (((bits >> (i * 4 + b)) & 1u) ? 0x01u : 0xFFu) << (8 * b);
      w[i] = v;
    }
    out = make_uint4(w[0], w[1], w[2], w[3]);
  }
  reinterpret_cast<uint4*>(job.out)[c] = out;
}

// The same expansion for the nodes of one rgbdslam_b200_nodes_create chunk: node f has its descriptors at desc + f * K * 32,
// its feature count in n[f] (device memory: no host round trip) and its tiles at out + f * n_pad * 256.
__global__ void __launch_bounds__(256) expand_i8_strided_kernel(const uint8_t* __restrict__ desc, int8_t* __restrict__ out,
                                                                const int* __restrict__ n, int K, int n_pad) {
  const int f = blockIdx.y;
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n_pad * 16) return;
  const int tile = c >> 11, rg = (c >> 7) & 15, kc = (c >> 3) & 15, rr = c & 7;
  const int row = tile * 128 + rg * 8 + rr;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (row < n[f]) {
    const uint8_t* d = desc + ((size_t)f * K + row) * 32 + kc * 2;
    const unsigned bits = (unsigned)d[0] | ((unsigned)d[1] << 8);
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      unsigned v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) v |= (((bits >> (i * 4 + b)) & 1u) ? 0x01u : 0xFFu) << (8 * b);
      w[i] = v;
    }
    o = make_uint4(w[0], w[1], w[2], w[3]);
  }
  reinterpret_cast<uint4*>(out + (size_t)f * n_pad * 256)[c] = o;
}

cudaError_t launch_expand_i8_strided(const uint8_t* desc, int8_t* out, const int* d_n, int nframes, int K, int n_pad,
                                     cudaStream_t stream) {
  if (nframes <= 0 || n_pad <= 0) return cudaSuccess;
  dim3 grid((n_pad * 16 + 255) / 256, nframes);
  expand_i8_strided_kernel<<<grid, 256, 0, stream>>>(desc, out, d_n, K, n_pad);
  return cudaGetLastError();
}

cudaError_t launch_expand_i8(const ExpandJob* d_jobs, int njobs, int max_n_pad, cudaStream_t stream) {
  if (njobs <= 0 || max_n_pad <= 0) return cudaSuccess;
  dim3 grid((max_n_pad * 16 + 255) / 256, njobs);
  expand_i8_kernel<<<grid, 256, 0, stream>>>(d_jobs);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// Pull a block of global memory into L2 ahead of its bulk copy (no shared memory, no completion tracking).
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
// One lane of a fully converged warp (elect.sync): the way to issue tcgen05 / bulk-copy instructions from warp-uniform code.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"):
//   bits [0,14)  start address >> 4
//   bits [16,30) leading-dimension byte offset >> 4  = distance between the two 8x16 B core matrices of
//                one K=32 B step                      = 128 B  (k_chunk stride of the tile layout)
//   bits [32,46) stride-dimension byte offset >> 4   = distance between consecutive 8-row groups = 2048 B
//   bits [46,48) descriptor version = 1 (Blackwell);  bits [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) |
         (1ull << 46);
}

constexpr uint32_t kTileA = 128 * 256;   // 32 KiB: 128 query rows x 256 int8
constexpr uint32_t kTileB = 256 * 256;   // 64 KiB: 256 train rows x 256 int8
constexpr uint32_t kSmemBars = 2 * kTileA + 2 * kTileB;
constexpr uint32_t kTcSmemBytes = kSmemBars + 128;
constexpr int kTcThreads = 192;
constexpr int kNoBest = (int)0x80000000;
// instruction descriptor: D=S32 (2<<4), A=INT8 (1<<7), B=INT8 (1<<10), both K-major, N=256, M=128
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

// instruction descriptor for the SIFT path: D=F32 (1<<4), A=BF16 (1<<7), B=BF16 (1<<10), K-major, N=256, M=128
constexpr uint32_t kIdescBF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

// MODE 0: Hamming (int8 +-1 operands, int32 accumulators, arg-max of the dot product).
// MODE 1: SIFT L2 (bf16 operands, 128-d rows are also 256 B => identical tile geometry; fp32 accumulators; the
//         epilogue keeps the 4 best candidates per query by score 2 a.b - |b|^2; exact fp32 re-ranking follows).
template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) tc_match_kernel(const HamItem* __restrict__ items, int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + 2 * kTileA;
  const uint32_t bars = sA + kSmemBars;
  // barrier slots (8 B each): full_a[2] 0,1 | empty_a[2] 2,3 | full_b[2] 4,5 | empty_b[2] 6,7 | tmem_full[2] 8,9 |
  // tmem_empty[2] 10,11 ; tmem base pointer at slot 12
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kSmemBars + 96);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 10; i++) mbar_init(bar(i), 1);
    mbar_init(bar(10), 4);
    mbar_init(bar(11), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_ptr_smem))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const HamItem item = items[it];
        mbar_wait(bar(2 + sa), pa ^ 1);
        mbar_expect_tx(bar(0 + sa), kTileA);
        bulk_g2s(sA + sa * kTileA, item.a, kTileA, bar(0 + sa));
        if (++sa == 2) { sa = 0; pa ^= 1; }
        for (int nb = 0; nb < item.n_btiles; nb++) {
          mbar_wait(bar(6 + sb), pb ^ 1);
          mbar_expect_tx(bar(4 + sb), kTileB);
          bulk_g2s(sB + sb * kTileB, item.b + (size_t)nb * kTileB, kTileB, bar(4 + sb));
          if (++sb == 2) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {  // warp-uniform walk, one elected lane issues (see tc_match256_kernel)
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, pacc = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int n_btiles = items[it].n_btiles;
      mbar_wait(bar(0 + sa), pa);
      tc_fence_after();
      for (int nb = 0; nb < n_btiles; nb++) {
        mbar_wait(bar(4 + sb), pb);
        mbar_wait(bar(10 + acc), pacc ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = make_desc(sA + sa * kTileA), db = make_desc(sB + sb * kTileB);
#pragma unroll
          for (int k = 0; k < 8; k++) {
            if (MODE == 0)
              tc_mma_i8(tmem_base + acc * 256, da + (uint64_t)(k * 16), db + (uint64_t)(k * 16), kIdescI8, k > 0 ? 1u : 0u);
            else
              tc_mma_bf16(tmem_base + acc * 256, da + (uint64_t)(k * 16), db + (uint64_t)(k * 16), kIdescBF16, k > 0 ? 1u : 0u);
          }
          tc_commit(bar(6 + sb));    // B stage may be refilled once these MMAs retire
          tc_commit(bar(8 + acc));   // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++sb == 2) { sb = 0; pb ^= 1; }
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      if (elect_one()) tc_commit(bar(2 + sa));  // A stage free
      __syncwarp();
      if (++sa == 2) { sa = 0; pa ^= 1; }
    }
  } else {
    const int wq = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = wq * 32 + lane;
    uint32_t acc = 0, pacc = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const HamItem item = items[it];
      int best = kNoBest;
      float s0 = -3.0e38f, s1 = -3.0e38f, s2 = -3.0e38f, s3 = -3.0e38f;  // MODE 1: 4 best scores, descending
      int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
      for (int nb = 0; nb < item.n_btiles; nb++) {
        mbar_wait(bar(8 + acc), pacc);
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * 256;
#pragma unroll 1
        for (int c = 0; c < 8; c++) {
          uint32_t v[32];
          tc_ld32(t0 + c * 32, v);
          tc_wait_ld();
          const int col0 = nb * 256 + c * 32;
          if (MODE == 0) {
            if (col0 + 32 <= item.nsearch) {
#pragma unroll
              for (int j = 0; j < 32; j++) best = max(best, (int)v[j] * 65536 + (65535 - (col0 + j)));
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (col0 + j < item.nsearch) best = max(best, (int)v[j] * 65536 + (65535 - (col0 + j)));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const int col = col0 + j;
              if (col < item.nsearch) {
                const float sc = fmaf(2.f, __uint_as_float(v[j]), -__ldg(item.bnorm + col));
                if (sc > s3) {  // insert (ties keep the earlier = lower index)
                  if (sc > s2) {
                    s3 = s2; i3 = i2;
                    if (sc > s1) {
                      s2 = s1; i2 = i1;
                      if (sc > s0) { s1 = s0; i1 = i0; s0 = sc; i0 = col; }
                      else { s1 = sc; i1 = col; }
                    } else { s2 = sc; i2 = col; }
                  } else { s3 = sc; i3 = col; }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(10 + acc));
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      if (row < item.nq_valid) {
        if (MODE == 0) {
          int2 o = make_int2(257, -1);  // features.cpp:172-173
          if (best != kNoBest) {
            const int s = best >> 16;  // dot product = 256 - 2*hd
            o.x = (256 - s) >> 1;
            o.y = 65535 - (best & 0xFFFF);
          }
          item.out[row] = o;
        } else {
          reinterpret_cast<int4*>(item.out)[row] = make_int4(i0, i1, i2, i3);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Variant with a 256-query A block per work item (two M=128 accumulator halves) and 128-row B tiles:
//   smem  : A 2 x 64 KiB (double buffered across items) + B 2 x 32 KiB
//   TMEM  : [stage 2][half 2][128 columns]  (512 columns)
//   L2->SM traffic per 256 queries: 64 KiB + nt/128 * 32 KiB  (1.8x less than the 128-query kernel, whose operand
//   fetch rate -- not the tensor pipe -- limits it: ncu shows 42 % tensor-pipe activity at ~29 B/clk/SM)
//   warps : 0 = bulk-copy producer, 1 = MMA issuer, 2..9 = epilogue (warp group g drains half g)
constexpr uint32_t kA256 = 256 * 256;  // 64 KiB
constexpr uint32_t kB128 = 128 * 256;  // 32 KiB
// Ring depths: the query block (A, 64 KiB) is double-buffered across items, the train tiles (B, 32 KiB) stream through
// a 3-deep ring -- one tile feeds ~1.1 k cycles of MMA, an L2 / HBM fetch takes longer than that, so two stages stall
// the tensor pipe on every tile (measured 61 us vs 75 us for the shallower variants).  224 KiB + barriers <= 227 KiB.
#ifndef RB200_TC_L2_PREFETCH
#define RB200_TC_L2_PREFETCH 0  // measured: no gain (bench `value` 1.50-1.52 M with, 1.54-1.59 M pairs/s without)
#endif
#ifndef RB200_TC256_ASTAGES
#define RB200_TC256_ASTAGES 2
#endif
#ifndef RB200_TC256_BSTAGES
#define RB200_TC256_BSTAGES 3
#endif
constexpr int kASt = RB200_TC256_ASTAGES, kBSt = RB200_TC256_BSTAGES;
constexpr uint32_t kSmem256Bars = kASt * kA256 + kBSt * kB128;
constexpr uint32_t kTc256SmemBytes = kSmem256Bars + 256;
static_assert(kTc256SmemBytes <= 232448, "tc_match256: shared memory over the 227 KiB per-CTA limit");
// barrier slots
constexpr int kBarAFull = 0, kBarAEmpty = kASt, kBarBFull = 2 * kASt, kBarBEmpty = 2 * kASt + kBSt,
              kBarAccFull = 2 * kASt + 2 * kBSt, kBarAccEmpty = kBarAccFull + 2, kBarCount = kBarAccEmpty + 2;
static_assert(kBarCount * 8 <= 192, "barrier area");
constexpr int kTc256Threads = 320;
constexpr uint32_t kIdescI8_N128 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescU8_N128 = (2u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);  // a/b format 0 = unsigned 8-bit
constexpr uint32_t kIdescBF16_N128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

#ifdef RB200_PROFILE_TC
// [role: loader, MMA issuer, epilogue warps][cycles waiting on A, on B, on accumulators, total, participants]
__device__ unsigned long long g_tc_prof[3][8];
extern "C" int rb200_debug_tc_profile(unsigned long long* out24, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out24, g_tc_prof, sizeof(unsigned long long) * 24);
  if (e == cudaSuccess && reset) {
    unsigned long long z[24] = {0};
    e = cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
  }
  return (int)e;
}
#endif
template <int MODE>
__global__ void __launch_bounds__(kTc256Threads, 1) tc_match256_kernel(const HamItem* __restrict__ items, int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + kASt * kA256;
  const uint32_t bars = sA + kSmem256Bars;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kSmem256Bars + 192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBarAccEmpty; i++) mbar_init(bar(i), 1);
    mbar_init(bar(kBarAccEmpty), 8);  // tmem_empty: one arrival per epilogue warp
    mbar_init(bar(kBarAccEmpty + 1), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_ptr_smem))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
#ifdef RB200_PROFILE_TC
  long long pf_a = 0, pf_b = 0, pf_acc = 0, pf_t0 = clock64();
#define RB200_TIMED_WAIT(ACC, ...) { const long long c0_ = clock64(); mbar_wait(__VA_ARGS__); ACC += clock64() - c0_; }
#else
#define RB200_TIMED_WAIT(ACC, ...) mbar_wait(__VA_ARGS__);
#endif

  if (warp == 0) {
    if (lane == 0) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const HamItem item = items[it];
#if RB200_TC_L2_PREFETCH
        // Experiment (off by default): the operands of this CTA's NEXT item go to L2 now.  No gain -- a cold launch is only
        // ~2 us slower than a warm one (59 vs 57 us), the 3-deep ring already covers the HBM latency.
        if (it + (int)gridDim.x < n_items) {
          const HamItem nx = items[it + gridDim.x];
          bulk_prefetch_l2(nx.a, kA256);
          for (int nb = 0; nb < nx.n_btiles; nb++) bulk_prefetch_l2(nx.b + (size_t)nb * kB128, kB128);
        }
#endif
        RB200_TIMED_WAIT(pf_a, bar(kBarAEmpty + sa), pa ^ 1)
        mbar_expect_tx(bar(kBarAFull + sa), kA256);
        bulk_g2s(sA + sa * kA256, item.a, kA256, bar(kBarAFull + sa));
        if (++sa == kASt) { sa = 0; pa ^= 1; }
        for (int nb = 0; nb < item.n_btiles; nb++) {
          RB200_TIMED_WAIT(pf_b, bar(kBarBEmpty + sb), pb ^ 1)
          mbar_expect_tx(bar(kBarBFull + sb), kB128);
          bulk_g2s(sB + sb * kB128, item.b + (size_t)nb * kB128, kB128, bar(kBarBFull + sb));
          if (++sb == kBSt) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the pipeline and one elected lane issues the tcgen05 instructions.  Issuing from inside an
    // `if (lane == 0)` region made the compiler wrap every UTCIMMA in an ELECT / BRA.U.ANY retry loop and rebuild both
    // descriptors from the shared-memory address: ~13 dependent instructions = 82 cycles per MMA that needs 68 cycles of
    // tensor time (in-kernel clock64 profile: the issuer waited on barriers only 15 % of the time, i.e. it WAS the bottleneck).
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, pacc = 0;
    constexpr uint32_t kIdesc = MODE == 0 ? kIdescI8_N128 : (MODE == 2 ? kIdescU8_N128 : kIdescBF16_N128);
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int n_btiles = items[it].n_btiles;
      RB200_TIMED_WAIT(pf_a, bar(kBarAFull + sa), pa)
      tc_fence_after();
      for (int nb = 0; nb < n_btiles; nb++) {
        RB200_TIMED_WAIT(pf_b, bar(kBarBFull + sb), pb)
        RB200_TIMED_WAIT(pf_acc, bar(kBarAccEmpty + acc), pacc ^ 1)
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = make_desc(sA + sa * kA256), db = make_desc(sB + sb * kB128);  // start-address field += bytes >> 4
#pragma unroll
          for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int k = 0; k < (MODE == 2 ? 4 : 8); k++) {  // u8 SIFT rows carry 128 B of data: 4 k-steps of 32 B
              const uint32_t d = tmem_base + acc * 256 + h * 128;
              const uint64_t dak = da + (uint64_t)((h * kTileA + k * 256) >> 4), dbk = db + (uint64_t)((k * 256) >> 4);
              if (MODE == 1) tc_mma_bf16(d, dak, dbk, kIdesc, k > 0 ? 1u : 0u);
              else tc_mma_i8(d, dak, dbk, kIdesc, k > 0 ? 1u : 0u);
            }
          }
          tc_commit(bar(kBarBEmpty + sb));
          tc_commit(bar(kBarAccFull + acc));
        }
        __syncwarp();
        if (++sb == kBSt) { sb = 0; pb ^= 1; }
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      if (elect_one()) tc_commit(bar(kBarAEmpty + sa));
      __syncwarp();
      if (++sa == kASt) { sa = 0; pa ^= 1; }
    }
  } else {
    const int h = (warp - 2) >> 2;  // accumulator half drained by this warp group
    const int wq = warp & 3;        // TMEM lane quadrant
    const int row = h * 128 + wq * 32 + lane;
    uint32_t acc = 0, pacc = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const HamItem item = items[it];
      int best = kNoBest;
      float s0 = -3.0e38f, s1 = -3.0e38f, s2 = -3.0e38f, s3 = -3.0e38f;
      int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
      long long u8_best = -1;  // MODE 2: (dot << 17) | (0x1FFFF - tie priority); -1 = no positive dot yet
      int u8_next = 0;
      for (int nb = 0; nb < item.n_btiles; nb++) {
        RB200_TIMED_WAIT(pf_acc, bar(kBarAccFull + acc), pacc)
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * 256 + h * 128;
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
          uint32_t v[32];
          tc_ld32(t0 + c * 32, v);
          tc_wait_ld();
          const int col0 = nb * 128 + c * 32;
          if (MODE == 0) {
            // (a software-pipelined drain with the next tcgen05.ld in flight and 4 independent arg-max accumulators was
            //  measured SLOWER: 63.5 vs 55.7 us -- the epilogue is not the critical path, the shared-memory operand
            //  bandwidth of the M128 x N128 MMAs is; see tc_match_wide_kernel)
            if (col0 + 32 <= item.nsearch) {
#pragma unroll
              for (int j = 0; j < 32; j++) best = max(best, (int)v[j] * 65536 + (65535 - (col0 + j)));
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (col0 + j < item.nsearch) best = max(best, (int)v[j] * 65536 + (65535 - (col0 + j)));
            }
          } else if (MODE == 2) {
            // SiftGPU RowMatch / ColMatch bookkeeping (ProgramCU.cu:1708-1736, 1463-1478, 1771-1777): strict >, only
            // positive dots register, the runner-up VALUE counts duplicates of the maximum.
            // common case first: a dot product below the current best only feeds the runner-up value (one IMNMX); the 64-bit
            // (dot, tie priority) key is built only for candidates that reach the best dot product
            const int nvalid = item.nsearch - col0;  // columns j < nvalid exist
            int best_dot = u8_best >= 0 ? (int)(u8_best >> 17) : 0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const int dv = j < nvalid ? (int)v[j] : 0;  // dot products are >= 0 (unsigned operands); 0 never registers
              if (dv >= best_dot && dv > 0) {
                const int col = col0 + j;
                const int prio = item.pad_ ? col : (((col & 31) << 12) | (col >> 5));
                const long long key = ((long long)dv << 17) | (long long)(0x1FFFF - prio);
                if (key > u8_best) {
                  if (u8_best >= 0) u8_next = max(u8_next, (int)(u8_best >> 17));
                  u8_best = key;
                  best_dot = dv;
                } else {
                  u8_next = max(u8_next, dv);
                }
              } else {
                u8_next = max(u8_next, dv);
              }
            }
          } else {
            // |b|^2 of the chunk's 32 train rows: one coalesced load per lane, handed round by shuffles (a global load per
            // element inside the branchy insertion made this epilogue 20 x slower than the MMAs it drains); columns past the
            // last train row get +inf, i.e. a score of -inf that never enters the list
            const float bn_l = (col0 + lane < item.nsearch) ? __ldg(item.bnorm + col0 + lane) : __int_as_float(0x7f800000);
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const float sc = fmaf(2.f, __uint_as_float(v[j]), -__shfl_sync(0xffffffffu, bn_l, j));
              if (sc > s3) {  // rare after the first tiles: insert (ties keep the earlier = lower index)
                const int col = col0 + j;
                if (sc > s2) {
                  s3 = s2; i3 = i2;
                  if (sc > s1) {
                    s2 = s1; i2 = i1;
                    if (sc > s0) { s1 = s0; i1 = i0; s0 = sc; i0 = col; }
                    else { s1 = sc; i1 = col; }
                  } else { s2 = sc; i2 = col; }
                } else { s3 = sc; i3 = col; }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBarAccEmpty + acc));
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
      if (row < item.nq_valid) {
        if (MODE == 0) {
          int2 o = make_int2(257, -1);
          if (best != kNoBest) {
            const int s = best >> 16;
            o.x = (256 - s) >> 1;
            o.y = 65535 - (best & 0xFFFF);
          }
          item.out[row] = o;
        } else if (MODE == 2) {
          int4 o = make_int4(0, -1, u8_next, 0);
          if (u8_best >= 0) {
            const int prio = 0x1FFFF - (int)(u8_best & 0x1FFFF);
            o.x = (int)(u8_best >> 17);
            o.y = item.pad_ ? prio : (((prio & 0xFFF) << 5) | (prio >> 12));
          }
          reinterpret_cast<int4*>(item.out)[row] = o;
        } else {
          reinterpret_cast<int4*>(item.out)[row] = make_int4(i0, i1, i2, i3);
        }
      }
    }
  }
#ifdef RB200_PROFILE_TC
  if (lane == 0 && MODE == 0) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : 2);
    atomicAdd(&g_tc_prof[role][0], (unsigned long long)pf_a);
    atomicAdd(&g_tc_prof[role][1], (unsigned long long)pf_b);
    atomicAdd(&g_tc_prof[role][2], (unsigned long long)pf_acc);
    atomicAdd(&g_tc_prof[role][3], (unsigned long long)(clock64() - pf_t0));
    atomicAdd(&g_tc_prof[role][4], 1ull);
  }
#endif
#undef RB200_TIMED_WAIT
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

template <int MODE>
static cudaError_t launch_tc256(const HamItem* d_items, int n_items, int sm_count, cudaStream_t stream) {
  if (n_items <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_match256_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc256SmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = n_items < sm_count ? n_items : sm_count;
  tc_match256_kernel<MODE><<<grid, kTc256Threads, kTc256SmemBytes, stream>>>(d_items, n_items);
  return cudaGetLastError();
}
cudaError_t launch_hamming_tc256(const HamItem* d_items, int n_items, int sm_count, cudaStream_t stream) {
  return launch_tc256<0>(d_items, n_items, sm_count, stream);
}
cudaError_t launch_l2_tc256(const HamItem* d_items, int n_items, int sm_count, cudaStream_t stream) {
  return launch_tc256<1>(d_items, n_items, sm_count, stream);
}
cudaError_t launch_siftgpu_tc256(const HamItem* d_items, int n_items, int sm_count, cudaStream_t stream) {
  return launch_tc256<2>(d_items, n_items, sm_count, stream);
}

// ---------------------------------------------------------------------------------------------
// Path 3 (default): 256-query items against 256-row train tiles.
// ncu on tc_match256_kernel: tensor pipe 65 % active, L2 21 %, HBM 31 % -- and each M128 x N128 x K32 MMA reads 8 KiB of
// operands from shared memory for 68 cycles of math (120 B/clk of a 128 B/clk port, plus the bulk-copy writes): the
// shared-memory operand bandwidth is the limiter.  N = 256 halves the A re-reads per MAC (12 KiB per 136 cycles = 88 B/clk).
// TMEM then holds exactly two M128 x N256 accumulators, one per query half; they double-buffer EACH OTHER: while half 1
// accumulates tile b, the epilogue drains half 0 of tile b.  A ring: 3 x 32 KiB (one query half per slot), B ring:
// 2 x 64 KiB.  16 epilogue warps: (half, column half, lane quadrant); the two column halves of a row meet in shared memory.
constexpr int kWideEpiWarps = 16;
constexpr int kWideThreads = 64 + kWideEpiWarps * 32;
constexpr int kWASlots = 3, kWBSlots = 2;
constexpr uint32_t kWideBarsOff = kWASlots * kTileA + kWBSlots * kTileB;  // 224 KiB
constexpr uint32_t kWideSmemBytes = kWideBarsOff + 256 + 1024;
static_assert(kWideSmemBytes <= 232448, "tc_match_wide: shared memory over the 227 KiB per-CTA limit");
constexpr int kWAFull = 0, kWAEmpty = kWASlots, kWBFull = 2 * kWASlots, kWBEmpty = kWBFull + kWBSlots, kWAccFull = kWBEmpty + kWBSlots,
              kWAccEmpty = kWAccFull + 2;
static_assert((kWAccEmpty + 2) * 8 <= 192, "barrier area");

__global__ void __launch_bounds__(kWideThreads, 1) tc_match_wide_kernel(const HamItem* __restrict__ items, int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + kWASlots * kTileA;
  const uint32_t bars = sA + kWideBarsOff;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kWideBarsOff + 192);
  int* s_best = reinterpret_cast<int*>(smem + kWideBarsOff + 256);  // [half][128 rows]: partial arg-max of column half 1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWAccEmpty; i++) mbar_init(bar(i), 1);
    mbar_init(bar(kWAccEmpty), 8);  // one arrival per epilogue warp of the half
    mbar_init(bar(kWAccEmpty + 1), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32((const void*)tmem_ptr_smem))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const HamItem item = items[it];
        const int n_halves = item.nq_valid > 128 ? 2 : 1;
        for (int h = 0; h < n_halves; h++) {
          mbar_wait(bar(kWAEmpty + sa), pa ^ 1);
          mbar_expect_tx(bar(kWAFull + sa), kTileA);
          bulk_g2s(sA + sa * kTileA, item.a + (size_t)h * kTileA, kTileA, bar(kWAFull + sa));
          if (++sa == kWASlots) { sa = 0; pa ^= 1; }
        }
        for (int nb = 0; nb < item.n_btiles; nb++) {
          mbar_wait(bar(kWBEmpty + sb), pb ^ 1);
          mbar_expect_tx(bar(kWBFull + sb), kTileB);
          bulk_g2s(sB + sb * kTileB, item.b + (size_t)nb * kTileB, kTileB, bar(kWBFull + sb));
          if (++sb == kWBSlots) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {  // warp-uniform walk, one elected lane issues (see tc_match256_kernel)
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, pacc0 = 0, pacc1 = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const HamItem item = items[it];
      const int n_halves = item.nq_valid > 128 ? 2 : 1;
      uint32_t slot[2], phase[2];
      for (int h = 0; h < n_halves; h++) {
        slot[h] = sa;
        phase[h] = pa;
        if (++sa == kWASlots) { sa = 0; pa ^= 1; }
      }
      for (int nb = 0; nb < item.n_btiles; nb++) {
        mbar_wait(bar(kWBFull + sb), pb);
        for (int h = 0; h < n_halves; h++) {
          if (nb == 0) mbar_wait(bar(kWAFull + slot[h]), phase[h]);
          mbar_wait(bar(kWAccEmpty + h), (h ? pacc1 : pacc0) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = make_desc(sA + slot[h] * kTileA), db = make_desc(sB + sb * kTileB);
            const uint32_t d = tmem_base + h * 256;
#pragma unroll
            for (int k = 0; k < 8; k++) tc_mma_i8(d, da + (uint64_t)(k * 16), db + (uint64_t)(k * 16), kIdescI8, k > 0 ? 1u : 0u);
            tc_commit(bar(kWAccFull + h));
            if (h == n_halves - 1) tc_commit(bar(kWBEmpty + sb));
          }
          __syncwarp();
          if (h) pacc1 ^= 1; else pacc0 ^= 1;
        }
        if (++sb == kWBSlots) { sb = 0; pb ^= 1; }
      }
      if (elect_one())
        for (int h = 0; h < n_halves; h++) tc_commit(bar(kWAEmpty + slot[h]));
      __syncwarp();
    }
  } else {
    const int e = warp - 2;
    const int h = e >> 3;         // query half whose accumulator this warp drains
    const int c = (e >> 2) & 1;   // column half of the 256-column tile
    const int wq = warp & 3;      // TMEM lane quadrant this warp may access
    const int row = h * 128 + wq * 32 + lane;
    uint32_t pacc = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const HamItem item = items[it];
      if (h == 1 && item.nq_valid <= 128) continue;  // the second half is never issued for short items
      int best = kNoBest;
      for (int nb = 0; nb < item.n_btiles; nb++) {
        mbar_wait(bar(kWAccFull + h), pacc);
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + h * 256 + c * 128;
#pragma unroll 1
        for (int ch = 0; ch < 4; ch++) {
          uint32_t v[32];
          tc_ld32(t0 + ch * 32, v);
          tc_wait_ld();
          const int col0 = nb * 256 + c * 128 + ch * 32;
          if (col0 + 32 <= item.nsearch) {
#pragma unroll
            for (int j = 0; j < 32; j++) best = max(best, (int)v[j] * 65536 + (65535 - (col0 + j)));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (col0 + j < item.nsearch) best = max(best, (int)v[j] * 65536 + (65535 - (col0 + j)));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kWAccEmpty + h));
        pacc ^= 1;
      }
      // the two column halves of a row meet here (8 warps = 256 threads per query half, named barriers 1 and 2)
      if (c == 1) s_best[h * 128 + wq * 32 + lane] = best;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + h) : "memory");
      if (c == 0) {
        best = max(best, s_best[h * 128 + wq * 32 + lane]);
        if (row < item.nq_valid) {
          int2 o = make_int2(257, -1);
          if (best != kNoBest) {
            const int sdot = best >> 16;
            o.x = (256 - sdot) >> 1;
            o.y = 65535 - (best & 0xFFFF);
          }
          item.out[row] = o;
        }
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + h) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

cudaError_t launch_hamming_tc_wide(const HamItem* d_items, int n_items, int sm_count, cudaStream_t stream) {
  if (n_items <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_match_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWideSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = n_items < sm_count ? n_items : sm_count;
  tc_match_wide_kernel<<<grid, kWideThreads, kWideSmemBytes, stream>>>(d_items, n_items);
  return cudaGetLastError();
}

cudaError_t launch_hamming_tc(const HamItem* d_items, int n_items, int sm_count, cudaStream_t stream) {
  if (n_items <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_match_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = n_items < sm_count ? n_items : sm_count;
  tc_match_kernel<0><<<grid, kTcThreads, kTcSmemBytes, stream>>>(d_items, n_items);
  return cudaGetLastError();
}

// SIFT L2: items carry bf16 operand tiles, bnorm (|b|^2 of the bf16-rounded train rows) and an int4 output per query.
cudaError_t launch_l2_tc(const HamItem* d_items, int n_items, int sm_count, cudaStream_t stream) {
  if (n_items <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_match_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = n_items < sm_count ? n_items : sm_count;
  tc_match_kernel<1><<<grid, kTcThreads, kTcSmemBytes, stream>>>(d_items, n_items);
  return cudaGetLastError();
}

}  // namespace rb200
