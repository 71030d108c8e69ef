// Thin inline-PTX wrappers for the Blackwell (sm_100a) async machinery used by conv_tc.cu:
// mbarrier, TMA (cp.async.bulk.tensor, tiled + im2col), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace ds {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// ---- TMA ---------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int32_t c0, int32_t c1, int32_t c2,
                                            int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// im2col mode over (C, W, H, N): base pixel (w, h, n) = top-left of the filter window, (off_w, off_h) = tap
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int32_t c,
                                                   int32_t w, int32_t h, int32_t n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

// TMA stores (smem tile -> global through a tensor map; out-of-bounds rows / columns are clipped), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// global[tile] += smem[tile] (fp32 add performed in L2)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// 4-D variants: the box (c, w, h, n) is clipped against the tensor in every dimension - the halo-tile convolution stores a tile of
// (rows x padded width) pixels and lets the TMA unit drop the padding columns and the rows past the image
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, uint32_t src, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_shared_v4_u32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// ---- tcgen05 -----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs (kind::f16), fp32 accumulate.  One thread issues.
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the SMs of one TPC execute one M=256 MMA ----------------------
// Each CTA stages its own 128 rows of A and HALF of the B tile; the leader (cluster rank 0) issues the MMAs, which read
// both CTAs' shared memory and write each CTA's own TMEM.  All loads signal the LEADER's full barrier; tcgen05.commit
// multicasts the "slot free" / "accumulator ready" arrivals to both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t bar_cluster, uint32_t dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(const CUtensorMap* m, uint32_t bar_cluster, uint32_t dst, int32_t c, int32_t w,
                                                        int32_t h, int32_t n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte-swizzled smem operand descriptor (rows of 128 B, 8-row atoms 1024 B apart).
// Field layout: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// ---- warp-uniform issue of the split-bf16 product -------------------------------------------------------------------------
// The MMA warp runs its loops with all 32 lanes converged and hands the elected lane's predicate to these blocks: an
// `if (lane == 0) { tcgen05.mma ... }` region is divergent code, in which the compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY
// waterfall loop and rebuilds each descriptor with a dozen uniform-datapath instructions - measured ~120 clk per MMA on the issuing
// thread (profiles/r02_ncu_halo_issue_bound.txt), i.e. narrow tiles (N <= 128: <= 64 clk of tensor work per MMA) were bound by
// instruction issue, not by operands.  Descriptors travel as 32-bit words: only the start-address field (low word) changes.
constexpr uint32_t UMMA_DESC_HI_SW128 = 0x40004040u;      // SBO 1024 B | version 1 | SWIZZLE_128B  (bits 32..63 of umma_desc_k_sw128)
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | 0x10000u; }

// acc (+)= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi for one 16-channel step; `elected` != 0 on exactly one lane of the converged warp
template <bool PAIR>
__device__ __forceinline__ void mma3_f16(uint32_t acc, uint32_t ah_lo, uint32_t al_lo, uint32_t bh_lo, uint32_t bl_lo, uint32_t idesc,
                                         uint32_t accumulate, uint32_t elected) {
  if (PAIR) {
    asm volatile(
        "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 dah, dal, dbh, dbl;\n\t"
        "setp.ne.b32 pe, %7, 0;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %7, %7;\n\t"
        "mov.b64 dah, {%1, %8};\n\tmov.b64 dal, {%2, %8};\n\tmov.b64 dbh, {%3, %8};\n\tmov.b64 dbl, {%4, %8};\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], dal, dbh, %5, pa;\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], dah, dbl, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], dah, dbh, %5, pt;\n\t}" ::"r"(acc),
        "r"(ah_lo), "r"(al_lo), "r"(bh_lo), "r"(bl_lo), "r"(idesc), "r"(accumulate), "r"(elected), "r"(UMMA_DESC_HI_SW128)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 dah, dal, dbh, dbl;\n\t"
        "setp.ne.b32 pe, %7, 0;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, %7, %7;\n\t"
        "mov.b64 dah, {%1, %8};\n\tmov.b64 dal, {%2, %8};\n\tmov.b64 dbh, {%3, %8};\n\tmov.b64 dbl, {%4, %8};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %5, pa;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %5, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %5, pt;\n\t}" ::"r"(acc),
        "r"(ah_lo), "r"(al_lo), "r"(bh_lo), "r"(bl_lo), "r"(idesc), "r"(accumulate), "r"(elected), "r"(UMMA_DESC_HI_SW128)
        : "memory");
  }
}
// tcgen05.commit by the elected lane of a converged warp (the lane that issued the MMAs)
template <bool PAIR>
__device__ __forceinline__ void mma_commit_elected(uint32_t bar, uint32_t elected) {
  if (PAIR) {
    const uint16_t mask = 3;
    asm volatile(
        "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %1, 0;\n\t"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t}" ::"r"(bar),
        "r"(elected), "h"(mask)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pe;\n\tsetp.ne.b32 pe, %1, 0;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
        "r"(elected)
        : "memory");
  }
}

// instruction descriptor: D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2, both K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t m, uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// kind::f16 with bf16 operands: D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1, both K-major
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

}  // namespace ptx
}  // namespace ds
