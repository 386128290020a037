// Inline-PTX wrappers for the sm_100a features the convolution kernels use:
// mbarrier, TMA (tiled + im2col loads, tiled stores), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp (the canonical way to issue warp-level TMA / tcgen05 instructions: ptxas knows
// that exactly one thread is active behind this predicate).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef DRS_MBAR_WAIT_OP
#define DRS_MBAR_WAIT_OP "mbarrier.try_wait.parity.shared::cta.b64"
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t" DRS_MBAR_WAIT_OP
      " p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a pipeline bug must not hang the GPU box.  After DRS_MBAR_SPIN_LIMIT failed try_waits (each suspends
// for a hardware-defined interval: seconds in total) the waiting thread records where it was stuck in the
// (host-mapped) diagnostic words and traps.
#ifndef DRS_MBAR_SPIN_LIMIT
#define DRS_MBAR_SPIN_LIMIT (1u << 22)   // ~4 us per failed try_wait: traps after ~15 s
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* diag, uint32_t tag) {
  if (mbar_try_wait(bar, parity)) return;
  // The bound is a spin count, not a clock: ptxas if-converts a conditional %globaltimer read into an unconditional
  // one, and that read (hundreds of ns) then sits in every iteration of the wait loop of the producer / MMA threads.
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins == DRS_MBAR_SPIN_LIMIT) {
      if (diag) {
        diag[0] = 0xDEAD0000u | tag;
        diag[1] = blockIdx.x;
        diag[2] = threadIdx.x;
        diag[3] = parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// address-based variants (shared-memory addresses kept as running 32-bit values in the issue loops)
__device__ __forceinline__ uint32_t mbar_try_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t" DRS_MBAR_WAIT_OP
      " p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity, volatile uint32_t* diag, uint32_t tag) {
  if (mbar_try_wait_addr(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait_addr(bar, parity)) {
    if (++spins == DRS_MBAR_SPIN_LIMIT) {
      if (diag) {
        diag[0] = 0xDEAD0000u | tag;
        diag[1] = blockIdx.x;
        diag[2] = bar;
        diag[3] = parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_addr(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// im2col-mode load of `pixelsPerColumn` consecutive output pixels starting at base pixel (w,h,n),
// channels [c, c+channelsPerPixel), displaced by the filter-tap offset (off_w, off_h).
__device__ __forceinline__ void tma_load_im2col_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c, int32_t w,
                                                   int32_t h, int32_t n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_addr(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_addr(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c, int32_t w,
                                                        int32_t h, int32_t n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// One K block = NMMA consecutive K=16 MMAs (descriptor address field += 2 per step: 32 bytes inside the swizzle span).
// Descriptors are passed as 32-bit halves (a_lo / b_lo carry the address, `hi` is common to both operands); `acc` = 0
// makes the first MMA overwrite the accumulator.
template <int NMMA>
__device__ __forceinline__ void umma_f16_kblock(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc);
template <>
__device__ __forceinline__ void umma_f16_kblock<4>(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t.reg .b64 da, db;\n\t.reg .b32 x, y;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 t, %5, %5;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 x, %1, 2;\n\tadd.u32 y, %2, 2;\n\tmov.b64 da, {x, %3};\n\tmov.b64 db, {y, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, t;\n\t"
      "add.u32 x, %1, 4;\n\tadd.u32 y, %2, 4;\n\tmov.b64 da, {x, %3};\n\tmov.b64 db, {y, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, t;\n\t"
      "add.u32 x, %1, 6;\n\tadd.u32 y, %2, 6;\n\tmov.b64 da, {x, %3};\n\tmov.b64 db, {y, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, t;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_f16_kblock<2>(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t.reg .b64 da, db;\n\t.reg .b32 x, y;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 t, %5, %5;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 x, %1, 2;\n\tadd.u32 y, %2, 2;\n\tmov.b64 da, {x, %3};\n\tmov.b64 db, {y, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, t;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}
// 32 lanes x 32 columns of fp32: thread `lane` receives row (taddr.lane + lane), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 format, version 1).  layout_type: 2 = SWIZZLE_128B,
// 4 = SWIZZLE_64B, 6 = SWIZZLE_32B, 0 = none.  Offsets are in bytes (encoded >> 4).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type & 7) << 61;
  return d;
}

}  // namespace ptx

// Instruction descriptor for tcgen05.mma kind::f16 (host + device).
//   fmt: 0 = F16, 1 = BF16 ; major: 0 = K-major, 1 = MN-major ; accumulator fp32.
static inline __host__ __device__ uint32_t make_idesc_f16(int M, int N, int a_fmt, int b_fmt, int a_major, int b_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= static_cast<uint32_t>(a_fmt & 7) << 7;
  d |= static_cast<uint32_t>(b_fmt & 7) << 10;
  d |= static_cast<uint32_t>(a_major & 1) << 15;
  d |= static_cast<uint32_t>(b_major & 1) << 16;
  d |= static_cast<uint32_t>((N >> 3) & 0x3F) << 17;
  d |= static_cast<uint32_t>((M >> 4) & 0x1F) << 24;
  return d;
}
