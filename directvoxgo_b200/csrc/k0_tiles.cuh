// k0_tiles.cuh -- one survivor's row of the rgbnet's X~ tile: trilinear k0 features (lib/dvgo.py:509) + the ray's share
// of the row (view embedding, the constant 1, padding), as saturated fp16 in the tensor core's operand layout
// (tc_common.cuh).  Shared by k0_gather_tiles_kernel (rows to global memory, fused_march.cu) and by the producer warps
// of mlp_fwd_gather_kernel (rows straight into the shared-memory tile the layer-1 MMA reads, fused_mlp.cu).
//
// In the operand layout a row is K1/8 CHUNKS of 16 bytes, one per 128-byte core matrix, and the same chunk of 8
// consecutive rows is one contiguous 128-byte line: every store below is a 16-byte chunk and the G = C/4 threads of
// consecutive survivors write the same chunk index in the same instruction, so global stores leave the SM as full
// 32-byte sectors (a first version with 8-byte stores per thread cost +47 us per step in partial-sector writes).
//   chunk k < ceil(G/2)  : feature units 2k, 2k+1 -> thread 2k of the survivor (unit 2k+1 arrives by one shuffle)
//   other chunks         : embedding / padding columns, dealt round-robin to the G threads
//   chunks that lie entirely past C + pe_stride are never written: the buffers are zero-initialised
#pragma once
#include "fused_scene.cuh"
#include "tc_common.cuh"

namespace dvgo {

template <bool kSmem>
__device__ __forceinline__ void tile_store16(uint8_t* p, const uint4& v) {
  if (kSmem) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tc::smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  } else {
    *reinterpret_cast<uint4*>(p) = v;
  }
}

// Threads per survivor and survivors per warp for a channel count.
template <int C>
struct K0TileShape {
  static constexpr int G = (C % 4 == 0) ? C / 4 : 1;
  static constexpr int SPW = 32 / G;
};

// Executed by ALL 32 lanes of a warp (shuffles inside).  Lane = (sub = lane / G, q = lane % G); `active`: the lane has a
// row to write; `live`: that row is a survivor (else it is written as zeros).  `cn`, `r` (ray index): the survivor's
// corners and ray (ignored unless live; cn.valid must be 0 otherwise).  trow = address of (row, column 0) of the tile.
template <int C, bool kSmem>
__device__ __forceinline__ void k0_tile_row(const float* __restrict__ k0, const Corner8& cn, int r, bool active,
                                            bool live, const uint8_t* __restrict__ pe16, int K1, int used_chunks,
                                            uint8_t* __restrict__ trow, int q) {
  constexpr int G = K0TileShape<C>::G;
  constexpr int FC = (G + 1) / 2;               // chunks that hold feature columns (C % 4 == 0 path)
  // the ray's share of the row, already fp16 (view_embedding's rows16): chunk k at byte 16 k
  const uint8_t* __restrict__ e = pe16 + static_cast<int64_t>(r) * (K1 * 2);
  if (C % 4 == 0) {
    // Every lane runs the same instruction stream (predicated): embedding chunk loads, the eight corner loads, then
    // the stores.  (Dealing the chunks out under divergent branches serialised one load round trip per branch.)
    constexpr int NJ = (8 - FC + G - 1) / G;          // embedding / padding chunks per thread, at most
    const int k_first = FC + (q + G - 1) % G;         // this thread's chunks: k_first + j G
    uint4 pv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = k_first + j * G;
      pv[j] = make_uint4(0u, 0u, 0u, 0u);
      if (live && k < used_chunks) pv[j] = __ldg(reinterpret_cast<const uint4*>(e + 16 * k));
    }
    const bool odd_tail = (q & 1) == 0 && q + 1 >= G;  // last feature chunk of an odd G: its upper unit is embedding
    uint2 ph = make_uint2(0u, 0u);
    if (live && odd_tail) ph = __ldg(reinterpret_cast<const uint2*>(e + 8 * (q + 1)));
    // All eight corner loads are issued unconditionally (a corner outside the grid reads voxel 0 and is skipped by
    // the predicated multiply-add below): with predicated loads ptxas gave successive loads the same destination
    // registers and waited for each before issuing the next -- 7 dependent memory round trips per thread, 65 % of
    // the kernel's stall samples (ncu, round 2).
    float4 f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      f[k] = __ldg(reinterpret_cast<const float4*>(k0 + static_cast<int64_t>(cn.ok(k) ? cn.off(k) : 0) * C + q * 4));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!cn.ok(k)) continue;
      const float wk = cn.w(k);
      acc[0] = fma_(f[k].x, wk, acc[0]); acc[1] = fma_(f[k].y, wk, acc[1]);
      acc[2] = fma_(f[k].z, wk, acc[2]); acc[3] = fma_(f[k].w, wk, acc[3]);
    }
    const uint2 mine = tc::pack4(make_float4(acc[0], acc[1], acc[2], acc[3]));
    uint2 nb;                         // the feature unit of the next thread of the same survivor
    nb.x = __shfl_down_sync(0xffffffffu, mine.x, 1);
    nb.y = __shfl_down_sync(0xffffffffu, mine.y, 1);
    if (active) {
      if ((q & 1) == 0) {             // feature chunk q/2: units q (mine) and q + 1 (neighbour, or embedding)
        const uint2 hi = odd_tail ? ph : nb;
        tile_store16<kSmem>(trow + (q >> 1) * 128, make_uint4(mine.x, mine.y, hi.x, hi.y));
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = k_first + j * G;
        if (k < used_chunks) tile_store16<kSmem>(trow + k * 128, pv[j]);
      }
    }
  } else if (active) {   // one thread per survivor, channel counts that are no multiple of 4 (3, 6, 9)
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!cn.ok(k)) continue;
      const float* __restrict__ v = k0 + static_cast<int64_t>(cn.off(k)) * C;
      const float wk = cn.w(k);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fma_(__ldg(v + c), wk, acc[c]);
    }
    constexpr int CK = (C + 7) / 8;      // chunks that hold at least one feature column
#pragma unroll
    for (int k = 0; k < CK; ++k) {       // feature halves, then whatever the ray's row holds in the rest of the chunk
      uint4 rowc = make_uint4(0u, 0u, 0u, 0u);
      if (live) rowc = __ldg(reinterpret_cast<const uint4*>(e + 16 * k));
      __half h[8];
      *reinterpret_cast<uint4*>(h) = rowc;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (8 * k + j < C) h[j] = __low2half(tc::pack2_sat(acc[8 * k + j], 0.f));
      tile_store16<kSmem>(trow + k * 128, *reinterpret_cast<const uint4*>(h));
    }
    for (int k = CK; k < used_chunks; ++k) {
      uint4 rowc = make_uint4(0u, 0u, 0u, 0u);
      if (live) rowc = __ldg(reinterpret_cast<const uint4*>(e + 16 * k));
      tile_store16<kSmem>(trow + k * 128, rowc);
    }
  }
}

}  // namespace dvgo
