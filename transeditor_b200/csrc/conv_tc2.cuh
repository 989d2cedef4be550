// Two-CTA (cta_group::2) variant of the tcgen05 implicit-GEMM convolution.
//
// Why: a single-CTA UMMA at M128 x N128 reads 4 KB of A + 4 KB of B from shared memory per K16 step (64 tensor
// cycles) = the whole 128 B/cycle shared-memory port, and the TMA writes of the next stage need the same port
// again, so the 1-CTA kernel tops out near 50 % tensor utilisation at N=128 and ~67 % at N=256 (measured:
// 1.01 and 1.5-1.6 PFLOP/s).  A CTA PAIR executes one M256 x N UMMA: each SM supplies its own 128 A rows and
// only HALF of the B tile (the halves are exchanged inside the pair), so per-SM shared-memory traffic per
// FLOP drops by 25 % (N128) / 33 % (N256).
//
// Protocol (rank 0 of the 2-CTA cluster is the leader and issues every MMA):
//   full[s]        lives in the LEADER.  Its producer arms it with the bytes of BOTH CTAs; both CTAs' TMA loads
//                  (cp.async.bulk.tensor...cta_group::2) complete_tx on it (peer bit of the address cleared).
//   empty[s]       one per CTA; tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11) releases the slot
//                  in both CTAs once the MMAs that read it are done.
//   tmem_full[b]   one per CTA (multicast commit); each CTA's epilogue warps drain their own 128 TMEM lanes.
//   tmem_empty[b]  lives in the LEADER, 8 arrivals (4 epilogue warps x 2 CTAs, remote mbarrier.arrive).
// Tile mapping: pair q handles M tiles 2q (rank 0) and 2q+1 (rank 1) of one n block; an odd tail tile is
// a dummy (batch coordinate out of range -> TMA zero fill, epilogue writes nothing).
#pragma once
#include "tc_common.cuh"

namespace te {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the even (leader) CTA of a pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask),
      "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0,
                                             int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask),
      "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0,
                                             int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask),
      "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive on the barrier at the same offset in BOTH CTAs once all prior MMAs of this thread are done
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
// arrive on the LEADER's copy of a barrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

}  // namespace te
