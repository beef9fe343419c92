// Tile geometry and device/host pieces shared by the per-layer kernel (tdnn_gemm.cu) and the whole-stack kernel
// (tdnn_stack.cu): one output tile = 256 frames x 256 channels per CTA pair, 128-byte K chunks, 8 epilogue warps.
#pragma once
#include <stdlib.h>
#include <string.h>

#include "ptx.cuh"
#include "xvec_internal.h"

namespace xvec {

constexpr int BM_CTA = 128;            // frame rows per CTA
constexpr int BM = 2 * BM_CTA;         // frame rows per tile (CTA pair)
constexpr int BN = XVEC_TILE_N;        // 256 output channels per tile
constexpr int BN_CTA = BN / 2;         // weight rows staged by each CTA
constexpr int BK_BYTES = 128;          // one 128B-swizzle atom row: 64 bf16 or 32 tf32 channels
constexpr int A_BYTES = BM_CTA * BK_BYTES;  // 16 KiB
constexpr int B_BYTES = BN_CTA * BK_BYTES;  // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;             // two per TMEM lane quarter, each owning half (128) of the tile's columns
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int OUT_BUF_BYTES = 32 * 128;  // one TMA-store box: 32 rows x 128 bytes
constexpr int OUT_BUFS = 2;              // staging boxes per epilogue warp (TMA stores in flight)

// {lo, hi} floats -> packed bf16x2 (lo in the low half), round-to-nearest; the _relu form clamps negatives to 0.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Statistics-pooling epilogue of one accumulator tile for one epilogue warp (replaces the reads of torch.mean / torch.std
// in stat_pool, main.py:59-63): 32 rows x 32 columns at a time, transposed through smem so that lane == column; column sums
// of r = relu(acc + bias) and r^2 per utterance present in the 32-row block, packed f32x2 arithmetic, fixed order.
struct PoolArgs {
  int rows, n;
  const float* bias;
  const int* row_utt;
  const int* blk_slot_base;
  float* part;
};
// tbase: TMEM address of this warp's lane quarter in the accumulator buffer; row0: first frame row of the warp's 32 rows;
// [cbeg, cend): the warp's columns of the tile; tr: the warp's dense 32x32 fp32 transpose tile (XOR-swizzled by 4-row groups);
// release(): called once, right after the warp's last tcgen05.ld of the tile has landed (it is NOT called when the warp
// reads nothing — no pooled rows or no valid columns — the caller then releases the buffer itself).
template <class Release>
__device__ __forceinline__ void pool_epilogue_tile(const PoolArgs& p, uint32_t tbase, int row0, int n0, int cbeg, int cend, float* tr,
                                                   int lane, Release&& release) {
  const int row = row0 + lane;
  const int c_last = min(cend, p.n - n0) - 1;  // last valid column of this warp's range (may be < cbeg)
  const int my_u = (row < p.rows) ? __ldg(p.row_utt + row) : -1;
  const int slot0 = __ldg(p.blk_slot_base + (row0 >> 5));
  const unsigned valid = __ballot_sync(0xffffffffu, my_u >= 0);
  for (int c = cbeg; c < cend && n0 + c < p.n; c += 32) {
    if (valid == 0u) break;  // block has no pooled rows (warp-uniform)
    uint32_t v[32];
    tmem_ld_32x32(tbase + c, v);
    tmem_ld_wait();
    if (c + 32 > c_last) release();
#pragma unroll
    for (int j = 0; j < 32; ++j)  // tr[column j][row lane], row group (lane/4) stored at slot (lane/4 ^ j%8)
      tr[j * 32 + ((((lane >> 2) ^ (j & 7)) << 2) | (lane & 3))] = __uint_as_float(v[j]);
    __syncwarp();
    const int col = n0 + c + lane;  // this lane now owns one column
    const float b = (col < p.n && p.bias) ? __ldg(p.bias + col) : 0.f;
    float4 z[8];                    // the column's 32 rows
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = *reinterpret_cast<const float4*>(tr + lane * 32 + ((i ^ (lane & 7)) << 2));
    __syncwarp();
    const float2 b2 = make_float2(b, b);
    unsigned remaining = valid;
    int seg = 0;
    while (remaining) {  // one pass per utterance present in this 32-row block (warp-uniform)
      const int lo = __ffs(remaining) - 1;
      const int u = __shfl_sync(0xffffffffu, my_u, lo);
      const unsigned m = __ballot_sync(0xffffffffu, my_u == u);
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
      if (m == 0xffffffffu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float2 a = __fadd2_rn(make_float2(z[i].x, z[i].y), b2);
          float2 d = __fadd2_rn(make_float2(z[i].z, z[i].w), b2);
          a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f);
          d.x = fmaxf(d.x, 0.f); d.y = fmaxf(d.y, 0.f);
          s2 = __fadd2_rn(s2, a);
          q2 = __ffma2_rn(a, a, q2);
          s2 = __fadd2_rn(s2, d);
          q2 = __ffma2_rn(d, d, q2);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float2 a = __fadd2_rn(make_float2(z[i].x, z[i].y), b2);
          float2 d = __fadd2_rn(make_float2(z[i].z, z[i].w), b2);
          a.x = ((m >> (4 * i + 0)) & 1u) ? fmaxf(a.x, 0.f) : 0.f;
          a.y = ((m >> (4 * i + 1)) & 1u) ? fmaxf(a.y, 0.f) : 0.f;
          d.x = ((m >> (4 * i + 2)) & 1u) ? fmaxf(d.x, 0.f) : 0.f;
          d.y = ((m >> (4 * i + 3)) & 1u) ? fmaxf(d.y, 0.f) : 0.f;
          s2 = __fadd2_rn(s2, a);
          q2 = __ffma2_rn(a, a, q2);
          s2 = __fadd2_rn(s2, d);
          q2 = __ffma2_rn(d, d, q2);
        }
      }
      if (col < p.n) {
        float* dst = p.part + static_cast<size_t>(slot0 + seg) * 2 * p.n + col;
        dst[0] = s2.x + s2.y;
        dst[p.n] = q2.x + q2.y;
      }
      remaining &= ~m;
      ++seg;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
// 2-D row-major tensor map with 128-byte swizzle (the inner box is one 128-byte chunk).
inline int make_tmap_2d(CUtensorMap* map, const void* ptr, int dtype, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                        uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return set_error(XVEC_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const uint64_t es = dtype == XVEC_BF16 ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) return set_error(XVEC_E_ARG, "matrix base pointer must be 16-byte aligned");
  if ((ld_elems * es) % 16 != 0) return set_error(XVEC_E_ARG, "row stride must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * es};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == XVEC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(XVEC_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
  return XVEC_OK;
}

// L2 eviction hints of the TMA traffic: activations in, weights, activations out.  XVEC_L2HINT = three digits
// (0 normal, 1 evict-first, 2 evict-last) overrides the default.
inline void l2_policies(unsigned long long* pol_a, unsigned long long* pol_b, unsigned long long* pol_y) {
  static const unsigned long long pol_tab[3] = {L2_EVICT_NORMAL, L2_EVICT_FIRST, L2_EVICT_LAST};
  static int hint[3] = {-1, 0, 0};
  if (hint[0] < 0) {
    const char* h = getenv("XVEC_L2HINT");
    const char* def = "122";
    if (!h || strlen(h) != 3) h = def;
    int v[3];
    for (int i = 0; i < 3; ++i) v[i] = (h[i] >= '0' && h[i] <= '2') ? h[i] - '0' : 0;
    hint[2] = v[2];
    hint[1] = v[1];
    hint[0] = v[0];
  }
  *pol_a = pol_tab[hint[0]];
  *pol_b = pol_tab[hint[1]];
  *pol_y = pol_tab[hint[2]];
}

}  // namespace xvec
