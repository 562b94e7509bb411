// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05 / TMEM PTX wrappers.
// Everything is inline PTX (no CUTLASS dependency) so the kernels compile in seconds.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace lb {

#define LB_DEVICE __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
LB_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

LB_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// cross-GPU flag waits (NVLink peer memory).  Every wait is BOUNDED: a dead or wedged peer must not hang the GPU
// forever — after `timeout_ns` the waiter prints who/what it was waiting for and traps, which surfaces
// as a CUDA error on the host (the trainer's emergency path / the launcher's restart then take over).
// ---------------------------------------------------------------------------------------------
LB_DEVICE unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
LB_DEVICE uint32_t ld_acquire_sys_u32(const uint32_t* a) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(a) : "memory");
  return v;
}
constexpr unsigned long long kDefaultSpinTimeoutNs = 120ull * 1000ull * 1000ull * 1000ull;  // 2 min
// wait until *addr >= target (wrap-safe signed distance); `what`/`who` only label the diagnostic
static __device__ __noinline__ void spin_timeout_trap(const uint32_t* addr, uint32_t target, int what, int who) {
  printf("[libai_b200] device spin-wait timed out: kind=%d rank/peer=%d flag=%p value=%u target=%u (block %d)\n", what, who,
         (const void*)addr, ld_acquire_sys_u32(addr), target, (int)blockIdx.x);
  __trap();
}
LB_DEVICE void spin_wait_ge_sys(const uint32_t* addr, uint32_t target, unsigned long long timeout_ns, int what, int who) {
  if (static_cast<int32_t>(ld_acquire_sys_u32(addr) - target) >= 0) return;
  const unsigned long long t0 = globaltimer_ns();
  const unsigned long long budget = timeout_ns ? timeout_ns : kDefaultSpinTimeoutNs;
  uint32_t polls = 0;
  while (static_cast<int32_t>(ld_acquire_sys_u32(addr) - target) < 0) {
    if ((++polls & 1023u) == 0 && globaltimer_ns() - t0 > budget) spin_timeout_trap(addr, target, what, who);
  }
}

// ---------------------------------------------------------------------------------------------
// counter-based RNG for dropout: Philox4x32 (7 rounds — passes BigCrush, Salmon et al. 2011) keyed by the PyTorch CUDA
// generator's (seed, offset) and indexed by ELEMENT POSITION, so forward, backward and any re-tiling regenerate the
// same mask.  The (seed, offset) pair comes from `at::CUDAGeneratorImpl::philox_cuda_state`: plain values in eager
// mode, device pointers + an intra-graph offset while a CUDA graph is being captured (PyTorch rewrites the pointed-to
// values before every replay) — so captured transformer blocks draw fresh masks on every replay.
// `salt` distinguishes tensor-parallel ranks in sharded regions (0 in replicated regions: identical masks).
// ---------------------------------------------------------------------------------------------
struct RngArgs {
  unsigned long long seed_val, offset_val;   // eager: the values
  const long long* seed_ptr;                 // capture: where PyTorch keeps them
  const long long* offset_ptr;
  unsigned int offset_intragraph;
  int captured;
  unsigned long long salt;
};
LB_DEVICE void rng_resolve(const RngArgs& a, unsigned long long& seed, unsigned long long& offset) {
  if (a.captured) {
    seed = static_cast<unsigned long long>(*a.seed_ptr);
    offset = static_cast<unsigned long long>(*a.offset_ptr) + a.offset_intragraph;
  } else {
    seed = a.seed_val;
    offset = a.offset_val;
  }
  seed ^= a.salt * 0x9E3779B97F4A7C15ull;
}
LB_DEVICE uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
LB_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
LB_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
LB_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

LB_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
LB_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
LB_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
LB_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) – 2D / 3D tiled loads into shared memory, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
LB_DEVICE void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
LB_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
LB_DEVICE void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
LB_DEVICE void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2,
                           int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// shared -> global tiled store (bulk async group completion)
LB_DEVICE void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
LB_DEVICE void tma_store_4d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1, int32_t c2,
                            int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
LB_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
LB_DEVICE void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
template <int N>
LB_DEVICE void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
template <uint32_t kCols>
LB_DEVICE void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
LB_DEVICE void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}
LB_DEVICE void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
LB_DEVICE void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// ---- CTA pair (cta_group::2): two SMs of one TPC work on one 256-row tile.  Each CTA stages its own 128 rows of A and
// HALF of the B tile in shared memory; the MMA (issued by the leader CTA only) reads both halves, each CTA's TMEM
// receives the accumulator rows of its own 128 rows.  PTX forms as in cute/arch/{mma_sm100_umma,copy_sm100_tma,
// tmem_allocator_sm100}.hpp and cutlass/arch/barrier.h of the vendored CUTLASS headers.
LB_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
LB_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
LB_DEVICE void tmem_alloc_2cta(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
LB_DEVICE void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}
// arrive (count 1) on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster (the form CUTLASS'
// ClusterBarrier::arrive(cta_id) uses; an explicit `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR per arrival,
// which throttled the peer's TMA producer to one k-block per fence: profiles/r2_10_ncu_gemm_2cta_first.txt)
LB_DEVICE void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// 2-D tiled TMA load issued by either CTA of a pair; the transaction bytes are credited to the mbarrier of the EVEN
// (leader) CTA at the same offset (peer bit of the shared::cluster address cleared)
LB_DEVICE void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
LB_DEVICE void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have completed
LB_DEVICE void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16/fp16 inputs, fp32 accumulate; one thread issues.
LB_DEVICE void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 8-bit float operands (E4M3/E5M2 selected by the instruction descriptor), 32 elements of K per instruction
LB_DEVICE void umma_f8_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (used by attention: P stays in tensor memory)
LB_DEVICE void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
LB_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread i <-> TMEM lane base+i)
LB_DEVICE void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
LB_DEVICE void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
LB_DEVICE void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
LB_DEVICE void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
LB_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
LB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp)
// ---------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, 128-byte swizzle, dense tiles whose rows are 128 bytes:
//   start address [0,14) (>>4), LBO [16,30) (>>4), SBO [32,46) (>>4), version=1 [46,48), layout=2 (SW128) [61,64)
LB_DEVICE uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// instruction descriptor for kind::f16 with bf16 inputs and fp32 accumulation
//   c_format=F32 (1) [4,6), a_format=BF16 (1) [7,10), b_format=BF16 (1) [10,13),
//   a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((n >> 3) << 17) | ((m >> 4) << 24);
}

// kind::f8f6f4 with E4M3 x E4M3 inputs (a_format = b_format = 0), fp32 accumulation, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_e4m3(uint32_t m, uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// numeric helpers
// ---------------------------------------------------------------------------------------------
LB_DEVICE uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
LB_DEVICE float2 unpack_bf16(uint32_t u) {
  // a bf16 is the upper half of an fp32: one shift for the low element, one mask for the high one (the intrinsic
  // compiled to PRMT + IMAD for the high half)
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_GELU_TANH = 2, ACT_RELU = 3, ACT_SILU = 4, ACT_QUICK_GELU = 5,
                 ACT_RESADD = 6 /* GEMM epilogue only: out = acc (+ bias) + pre_in, i.e. the residual add */ };

// erfc(|x|/sqrt2) and exp(-x^2/2) from one MUFU.EX2 + one MUFU.RCP (Abramowitz-Stegun 7.1.26, |err(erf)| < 1.5e-7 -
// three orders of magnitude below bf16 resolution).  The libm erff costs ~60 instructions with a divergent branch,
// which made the 4-warp GEMM epilogue (one thread per accumulator row) the bottleneck of the h->4h projection.
LB_DEVICE void gelu_terms(float x, float& erfc_abs, float& gauss) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));  // exp(-x^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  erfc_abs = poly * t * e;
  gauss = e;
}
LB_DEVICE float gelu_fast(float x) {
  float q, e;
  gelu_terms(x, q, e);
  const float h = 0.5f * q;                 // 0.5 * erfc(|x|/sqrt2)
  return x * (x >= 0.f ? 1.0f - h : h);     // no cancellation on the negative side
}
LB_DEVICE float gelu_grad_fast(float x) {
  float q, e;
  gelu_terms(x, q, e);
  const float h = 0.5f * q;
  const float cdf = x >= 0.f ? 1.0f - h : h;
  return fmaf(x * 0.3989422804014327f, e, cdf);
}
// 2^x on the SFU without the denormal-range fix-up branch that exp2f() compiles to (one BSSY/BRA/BSYNC triple per
// element in the softmax loops); ex2.approx(-inf) = +0.
LB_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
LB_DEVICE float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

LB_DEVICE float act_fwd(float x, int act) {
  switch (act) {
    case ACT_GELU:
      return gelu_fast(x);
    case ACT_GELU_TANH: {
      float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.0f + tanh_fast(u));
    }
    case ACT_RELU:
      return x > 0.f ? x : 0.f;
    case ACT_SILU:
      return x / (1.0f + __expf(-x));
    case ACT_QUICK_GELU:
      return x / (1.0f + __expf(-1.702f * x));
    default:
      return x;
  }
}
// d act(x) / dx
LB_DEVICE float act_grad(float x, int act) {
  switch (act) {
    case ACT_GELU:
      return gelu_grad_fast(x);
    case ACT_GELU_TANH: {
      float x2 = x * x;
      float u = 0.7978845608028654f * (x + 0.044715f * x * x2);
      float t = tanh_fast(u);
      float du = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x2);
      return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
    }
    case ACT_RELU:
      return x > 0.f ? 1.f : 0.f;
    case ACT_SILU: {
      float s = 1.0f / (1.0f + __expf(-x));
      return s * (1.0f + x * (1.0f - s));
    }
    case ACT_QUICK_GELU: {
      float s = 1.0f / (1.0f + __expf(-1.702f * x));
      return s * (1.0f + 1.702f * x * (1.0f - s));
    }
    default:
      return 1.f;
  }
}

LB_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
LB_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace lb

// ---------------------------------------------------------------------------------------------
// host: tensor-map creation through the driver entry point (no -lcuda link dependency)
// ---------------------------------------------------------------------------------------------
namespace lb_host {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();
// rank-`rank` bf16 tensor; dims/strides innermost first (strides in bytes, strides[0] implied = 2)
bool make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);
bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, CUtensorMapSwizzle swizzle);
}  // namespace lb_host
