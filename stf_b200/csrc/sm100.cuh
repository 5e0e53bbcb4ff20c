// Thin inline-PTX layer for the Blackwell (sm_100a) features libstf_b200 uses:
// mbarrier, 1-D bulk TMA copies (cp.async.bulk -> UBLKCP), tcgen05.mma kind::tf32 with TMEM
// accumulators (UTC*MMA), tcgen05.ld (LDTM), TMEM allocation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stf {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp (elect.sync): ptxas keeps warp-uniform operands of the guarded single-thread
// instructions (tcgen05.mma / commit / bulk copies) in uniform registers instead of moving them per lane.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// One probe (the hardware suspends the warp for a bounded, implementation-defined time if the phase is pending).
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait for the phase with the given parity.  The inner loop is three instructions (probe, branch, count), so
// waiting warps leave the issue slots to the working ones; every 4096 failed probes the wall clock is checked
// and a protocol bug traps (launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 4096;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "@p bra D_%=;\n\t"
        "sub.u32 n, n, 1;\n\t"
        "setp.ne.u32 p, n, 0;\n\t"
        "@p bra W_%=;\n\t"
        "setp.eq.u32 p, n, 1;\n\t"   // n == 0 here: p = false
        "D_%=:\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 20000000000LL) __trap();
  }
}

// Non-blocking probes (mbarrier.test_wait) in a tight loop, for hand-offs that sit on a short per-step critical path (the
// fused MLP kernel: two hand-offs per 64-column chunk).  try_wait may park the thread for an implementation-defined time when
// the phase is pending; measured on B200 the two variants performed the same in that kernel (the MMA issuer bound it), so
// this is a choice of semantics -- never parked -- not a measured win.
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 65536;\n\t"
        "W_%=:\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "@p bra D_%=;\n\t"
        "sub.u32 n, n, 1;\n\t"
        "setp.ne.u32 p, n, 0;\n\t"
        "@p bra W_%=;\n\t"
        "setp.eq.u32 p, n, 1;\n\t"
        "D_%=:\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 20000000000LL) __trap();
  }
}

// Same, for roles that usually wait long (the epilogue warps wait a whole K loop for their accumulator): back off
// with nanosleep between probes so that the waiting warps do not compete for issue slots with the working ones.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (uint32_t spins = 1;; ++spins) {
    __nanosleep(128);
    if (mbar_try_wait(bar, parity)) return;
    if ((spins & 4095u) == 0 && clock64() - t0 > 20000000000LL) __trap();
  }
}

// Roles whose wait is off the critical path (deep rings ahead of them): sleep `ns` between probes.  ncu showed the
// tight-spin variant executing ~40 % of all warp instructions of the kernel in landing-barrier probes, competing for
// issue slots with the epilogue warps that bound the tile rate.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, uint32_t ns) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (uint32_t spins = 1;; ++spins) {
    __nanosleep(ns);
    if (mbar_try_wait(bar, parity)) return;
    if ((spins & 4095u) == 0 && clock64() - t0 > 20000000000LL) __trap();
  }
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- 1-D bulk TMA copy
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Same copy, delivered to the same shared-memory offset (and signalled on the same mbarrier offset) of every CTA of the
// cluster selected by cta_mask: CTAs that stream the same weight tile fetch it from L2 once.
__device__ __forceinline__ void bulk_copy_g2s_multicast(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                                        uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp as alloc
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, K-major operand, no swizzle ("interleaved" canonical layout):
// 8-row x 16-byte core matrices; LBO = byte distance between the two 16-byte K chunks of one
// MMA, SBO = byte distance between consecutive 8-row groups.  Field layout as in CUTLASS
// cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1
// layout_type[61,64)=0.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major (InstrDescriptor in the
// same header): c_format[4,6)=1 (F32), a_format[7,10)=2 (TF32), b_format[10,13)=2,
// n_dim[17,23)=N>>3, m_dim[24,29)=M>>4.
__host__ __device__ inline uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T, issued by ONE thread for the whole CTA.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all MMAs issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Same arrival on the barrier at this offset in every CTA of cta_mask (a weight stage shared by a CTA pair is free when
// both CTAs' MMAs have read it).
__device__ __forceinline__ void umma_commit_multicast(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ float to_tf32(float x) {  // round-to-nearest (ties away), low 13 bits zero
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace sm100
}  // namespace stf
