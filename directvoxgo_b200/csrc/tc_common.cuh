// tc_common.cuh -- sm_100a tensor-core building blocks for the rgbnet kernels: tcgen05.mma
// (kind::f16 with FP16 operands, fp32 accumulate in TMEM), shared-memory matrix descriptors, TMEM
// allocation and loads, mbarrier completion tracking.  Inline PTX only; sm_100a only.
//
// Why FP16 operands and not TF32: both carry a 10-bit mantissa (same rounding error), but for 32-bit
// operands the tensor core accepts MN-major ("transposed") tiles only in the SWIZZLE_128B_BASE32B
// layout (measured: a no-swizzle MN-major tf32 descriptor yields zeros; CUTLASS: "for mn-major tf32
// operands, SW128_32B is the only available smem layout"), which cannot alias a K-major tile.  The
// backward GEMMs need every activation / weight tile in BOTH orientations, so tf32 would double the
// shared-memory footprint past 227 KB.  With 16-bit operands ONE physical layout serves both.
//
// Operand tiles in shared memory: the no-swizzle "interleaved" canonical layout of 8x16-byte core
// matrices (8 rows x 8 halves):
//
//     addr(r, c) = base + (r/8)*group_stride + (c/8)*128 + (r%8)*16 + (c%8)*2        (fp16, T = 8)
//
// r is the index that is NOT contiguous in memory (sample index for activations, out-feature for
// weights [out][in]); c is the contiguous one.  group_stride = (cols/8)*128 bytes.
// The same bytes can be handed to the tensor core either way round:
//   * K-major operand   (r = M/N index, c = K index):  LBO = 128 (next 16-byte K chunk),
//                        SBO = group_stride (next 8 rows);  one K=16 MMA step advances 256 B
//   * MN-major operand  (r = K index,  c = M/N index):  SBO = 128 (next 16-byte M/N chunk),
//                        LBO = group_stride (next 8 K rows); one K=16 MMA step advances 2*group_stride
// which is what lets the backward GEMMs (dW = dZ^T * H, reduction over samples) read the very
// buffers the forward GEMMs wrote, with no transposes.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dvgo {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Byte offset of element (r, c) in a canonical tile with `cols` fp16 columns (cols % 8 == 0).
__host__ __device__ __forceinline__ uint32_t tile_off(int r, int c, int cols) {
  return static_cast<uint32_t>((r >> 3) * (cols >> 3) * 128 + (c >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2);
}
__host__ __device__ constexpr uint32_t tile_bytes(int rows, int cols) {
  return static_cast<uint32_t>(((rows + 7) / 8) * (cols / 8) * 128);
}
__host__ __device__ constexpr uint32_t group_stride(int cols) { return static_cast<uint32_t>((cols / 8) * 128); }
constexpr int kMmaK = 16;  // K elements consumed by one kind::f16 tcgen05.mma

// fp32 pair -> packed fp16 pair, round-to-nearest, SATURATING to +-65504 (one F2FP.SATFINITE instruction): a
// feature / activation / scaled gradient beyond FP16's range clamps instead of becoming inf and poisoning the
// training state through Adam (the reference's fp32 nn.Linear has no such range limit).  NaN stays NaN and is
// reported through the status word (counters[1] bit 1).
__device__ __forceinline__ __half2 pack2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ uint2 pack4(const float4& v) {
  const __half2 a = pack2_sat(v.x, v.y), b = pack2_sat(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// Survivor tiles in global memory ("X~ tiles" / "dZ3 tiles"): tile t holds survivors [128 t, 128 t + 128) as one
// [128 rows][cols] fp16 block in the canonical operand layout above, so a CTA pulls a whole tile with ONE bulk copy
// and hands it to the tensor core untouched.  Producers cover the rows up to the next multiple of 256 with zeros (the
// backward kernel works on tile pairs) and the buffers hold one tile more than ceil(capacity / 128).
__host__ __device__ constexpr int64_t tiles_for(int64_t surv_cap) { return (surv_cap + 127) / 128 + 1; }

// One elected lane of a fully converged warp (elect.sync).  MMA issue MUST sit under this predicate inside a
// warp-uniform branch: under a plain `if (threadIdx.x == 0)` nvcc cannot prove a single active thread and wraps EVERY
// tcgen05.mma in an ELECT / BRA.U.ANY serialisation loop (measured round 1 as "no MMA issues faster than 49-72 cycles");
// under elect.sync the UTCHMMAs are emitted back to back.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// 64-bit shared-memory matrix descriptor (sm_100 "version 1", no swizzle, base offset 0).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version for Blackwell
  return d;         // layout_type (bits 61-63) = 0: SWIZZLE_NONE
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, int cols) {
  return make_desc(saddr, 128u, group_stride(cols));
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, int cols) {
  return make_desc(saddr, group_stride(cols), 128u);
}

// Instruction descriptor: D = fp32 (c_format 1), A = B = fp16 (format 0), dense, no negate.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows = TMEM lanes, K contiguous along the columns, two fp16
// per 32-bit column, so one K = 16 step spans 8 columns) is read from tensor memory instead of shared memory.
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// registers -> TMEM: 32 lanes (this warp's quarter) x 8 consecutive 32-bit columns (= 16 packed fp16 per lane).
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* u) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
               :
               : "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};\n"
      :
      : "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]),
        "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Make all previously issued MMAs arrive on an mbarrier when they complete.
__device__ __forceinline__ void mma_commit(uint32_t bar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :
               : "r"(bar_saddr)
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar_saddr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar_saddr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_saddr) : "memory");
}
// Wait for the phase with the given parity; bounded spin so a protocol bug traps instead of hanging
// the GPU (a hung box is a strike on the shared pool).
__device__ __forceinline__ void mbar_wait(uint32_t bar_saddr, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar_saddr), "r"(parity)
        : "memory");
    if (spin > (1u << 22)) __trap();
  }
}

// ---- bulk asynchronous copy global -> shared (the TMA engine's 1-D form; SASS UBLKCP) ----
// One thread arms the mbarrier with the byte count, then issues the copy; the barrier's phase completes when the
// bytes have landed (complete_tx).  dst / src 16-byte aligned, bytes a multiple of 16.  No registers, no scoreboard:
// the survivor tiles of the rgbnet kernels arrive this way (their producers write them in the operand layout).
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_saddr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t bar_saddr) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(dst_saddr), "l"(src), "r"(bytes), "r"(bar_saddr)
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// Non-blocking test of an mbarrier phase (the polling MMA issuer serves whichever tile context is ready first).
__device__ __forceinline__ bool mbar_test(uint32_t bar_saddr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar_saddr), "r"(parity)
      : "memory");
  return done != 0;
}

// TMEM allocation (one full warp executes these).
__device__ __forceinline__ void tmem_alloc(uint32_t dst_saddr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_saddr), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// TMEM -> registers: 32 lanes (this warp's quarter) x 16 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15}, [%16];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc
}  // namespace dvgo
