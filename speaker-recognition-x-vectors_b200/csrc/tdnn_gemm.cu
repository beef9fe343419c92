// TDNN layer as a tcgen05 GEMM over the flat frame matrix — sm_100a only.
//
//   y[r, n] = epi( sum_{tap j} sum_{ch} x[r + off_j, ch] * W[n, j*Cin + ch] )
//
// replaces tdnn_layer.py:26-41 (get_time_context + torch.cat + nn.Linear + ReLU + eval BatchNorm1d) without ever
// materialising the unfolded (rows x taps*Cin) tensor: the K loop runs over (tap, 128-byte channel chunk) and the
// TMA producer simply shifts the row coordinate of the A tile by the tap's frame offset.  Rows past the end of the
// matrix and channels past Cin are zero-filled by TMA, the packed weights carry matching zero padding.
//
// Persistent, warp-specialised CTA (192 threads, 1 CTA/SM):
//   warp 0      TMA producer        (A tile 128 rows x 128 B, B tile 256 rows x 128 B, 4-stage mbarrier ring)
//   warp 1      TMEM owner + tcgen05.mma issuer (UMMA 128x256x{16 bf16 | 8 tf32}, fp32 accumulators in TMEM,
//               two 256-column accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2-5   epilogue: tcgen05.ld -> bias/ReLU/BatchNorm affine -> global store          (EPI_STORE_*)
//                         or -> per-utterance column sums of r and r^2 (statistics pooling)  (EPI_POOL)
#include "ptx.cuh"
#include "xvec_internal.h"
#include <cuda_bf16.h>

namespace xvec {

constexpr int BM = 128;
constexpr int BN = XVEC_TILE_N;  // 256
constexpr int BK_BYTES = 128;    // one 128B-swizzle atom row: 64 bf16 or 32 tf32 channels
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK_BYTES;  // 16 KiB
constexpr int B_BYTES = BN * BK_BYTES;  // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int MAX_NPAD = 1536;          // column parameters staged in shared memory by the store epilogue
constexpr int TR_LD = 33;               // padded row stride of the per-warp 32x32 transpose tile

enum { EPI_STORE_F32 = 0, EPI_STORE_BF16 = 1, EPI_POOL = 2 };

struct GemmParams {
  int rows;     // output rows
  int n;        // valid output columns
  int m_tiles, n_tiles;
  int taps, cpt;  // K loop = taps * cpt chunks of 128 bytes
  int tap_off[XVEC_MAX_TAPS];
  const float* bias;
  const float* scale;
  const float* shift;
  int relu;
  void* out;
  long long ldo;
  int vec_ok;  // output rows are 16-byte aligned -> vector stores
  const int* row_utt;
  const int* blk_slot_base;
  float* part;
};

template <int kEpi>
constexpr int epi_smem_bytes() {
  return kEpi == EPI_POOL ? 4 * 32 * TR_LD * 4 : 3 * MAX_NPAD * 4;
}
template <int kEpi>
constexpr int gemm_smem_bytes() {
  return 1024 + STAGES * STAGE_BYTES + epi_smem_bytes<kEpi>();
}

template <bool kTf32, int kEpi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tdnn_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int tap_off_s[XVEC_MAX_TAPS];

  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_smem = base + STAGES * STAGE_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int BKE = kTf32 ? 32 : 64;  // elements per 128-byte chunk
  const int kblocks = p.taps * p.cpt;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x < XVEC_MAX_TAPS) tap_off_s[threadIdx.x] = p.tap_off[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
  if constexpr (kEpi != EPI_POOL) {
    float* cb = reinterpret_cast<float*>(epi_smem);
    const int npad = p.n_tiles * BN;
    for (int i = threadIdx.x; i < npad; i += GEMM_THREADS) {
      const bool ok = i < p.n;
      cb[i] = (ok && p.bias) ? p.bias[i] : 0.f;
      cb[MAX_NPAD + i] = (ok && p.scale) ? p.scale[i] : 1.f;
      cb[2 * MAX_NPAD + i] = (ok && p.shift) ? p.shift[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t / p.n_tiles) * BM;
        const int n0 = (t % p.n_tiles) * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          const int tap = kb / p.cpt;
          const int ch = kb - tap * p.cpt;
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = base + stage * STAGE_BYTES;
          tma_load_2d(sa, &tmA, &full_bar[stage], ch * BKE, m0 + tap_off_s[tap]);
          tma_load_2d(sa + A_BYTES, &tmB, &full_bar[stage], kb * BKE, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kTf32 ? 2u : 1u, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use = (it >> 1) & 1;
        mbar_wait(&tempty_bar[buf], use ^ 1u, 2);  // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(base + stage * STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(a_addr);
          const uint64_t db = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x 32 bytes of K inside the swizzle atom
            umma_ss<kTf32>(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[buf]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int m0 = (t / p.n_tiles) * BM;
      const int n0 = (t % p.n_tiles) * BN;
      const int buf = it & 1;
      const uint32_t use = (it >> 1) & 1;
      mbar_wait(&tfull_bar[buf], use, 4);
      tc_fence_after();
      const uint32_t tbase = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);
      const int row = m0 + q * 32 + lane;

      if constexpr (kEpi == EPI_POOL) {
        float* tr = reinterpret_cast<float*>(epi_smem) + (warp - 2) * (32 * TR_LD);
        const int my_u = (row < p.rows) ? __ldg(p.row_utt + row) : -1;
        const int slot0 = __ldg(p.blk_slot_base + (m0 >> 5) + q);
        const unsigned valid = __ballot_sync(0xffffffffu, my_u >= 0);
        for (int c = 0; c < BN && n0 + c < p.n; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tbase + c, v);
          tmem_ld_wait();
          if (valid == 0u) continue;  // block has no pooled rows (uniform)
#pragma unroll
          for (int j = 0; j < 32; ++j) tr[lane * TR_LD + j] = __uint_as_float(v[j]);
          __syncwarp();
          const int col = n0 + c + lane;  // this lane now owns one column
          const float b = (col < p.n && p.bias) ? __ldg(p.bias + col) : 0.f;
          unsigned remaining = valid;
          int seg = 0;
          while (remaining) {  // one pass per utterance present in this 32-row block (warp-uniform)
            const int lo = __ffs(remaining) - 1;
            const int u = __shfl_sync(0xffffffffu, my_u, lo);
            const unsigned m = __ballot_sync(0xffffffffu, my_u == u);
            const int hi = 32 - __clz(m);
            float s = 0.f, ss = 0.f;
            if (lo == 0 && hi == 32) {
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const float z = fmaxf(tr[r * TR_LD + lane] + b, 0.f);
                s += z;
                ss = fmaf(z, z, ss);
              }
            } else {
              for (int r = lo; r < hi; ++r) {
                const float z = fmaxf(tr[r * TR_LD + lane] + b, 0.f);
                s += z;
                ss = fmaf(z, z, ss);
              }
            }
            if (col < p.n) {
              float* dst = p.part + static_cast<size_t>(slot0 + seg) * 2 * p.n + col;
              dst[0] = s;
              dst[p.n] = ss;
            }
            remaining &= ~m;
            ++seg;
          }
          __syncwarp();
        }
      } else {
        const float* cb = reinterpret_cast<const float*>(epi_smem);
        for (int c = 0; c < BN && n0 + c < p.n; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tbase + c, v);
          tmem_ld_wait();
          const int col0 = n0 + c;
          float o[32];
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(cb + col0 + j4);
            const float4 sc = *reinterpret_cast<const float4*>(cb + MAX_NPAD + col0 + j4);
            const float4 sh = *reinterpret_cast<const float4*>(cb + 2 * MAX_NPAD + col0 + j4);
            float a0 = __uint_as_float(v[j4 + 0]) + bb.x, a1 = __uint_as_float(v[j4 + 1]) + bb.y;
            float a2 = __uint_as_float(v[j4 + 2]) + bb.z, a3 = __uint_as_float(v[j4 + 3]) + bb.w;
            if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
            o[j4 + 0] = fmaf(a0, sc.x, sh.x);
            o[j4 + 1] = fmaf(a1, sc.y, sh.y);
            o[j4 + 2] = fmaf(a2, sc.z, sh.z);
            o[j4 + 3] = fmaf(a3, sc.w, sh.w);
          }
          if (row < p.rows) {
            const bool full = (col0 + 32 <= p.n) && p.vec_ok;
            if constexpr (kEpi == EPI_STORE_BF16) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 w;
                  __nv_bfloat162 h0 = __floats2bfloat162_rn(o[j + 0], o[j + 1]);
                  __nv_bfloat162 h1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]);
                  __nv_bfloat162 h3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
                  w.x = *reinterpret_cast<uint32_t*>(&h0);
                  w.y = *reinterpret_cast<uint32_t*>(&h1);
                  w.z = *reinterpret_cast<uint32_t*>(&h2);
                  w.w = *reinterpret_cast<uint32_t*>(&h3);
                  *reinterpret_cast<uint4*>(dst + j) = w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.n) dst[j] = __float2bfloat16_rn(o[j]);
              }
            } else {
              float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.n) dst[j] = o[j];
              }
            }
          }
        }
      }
      // all of this warp's TMEM reads of the buffer are complete -> hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int make_tmap_2d(CUtensorMap* map, const void* ptr, int dtype, uint64_t inner, uint64_t outer, uint64_t ld_elems,
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

template <bool kTf32, int kEpi>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t st) {
  static bool configured[64] = {};  // per instantiation and device
  constexpr int smem = gemm_smem_bytes<kEpi>();
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tdnn_gemm_kernel<kTf32, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(XVEC_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  tdnn_gemm_kernel<kTf32, kEpi><<<grid, GEMM_THREADS, smem, st>>>(ta, tb, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(XVEC_E_CUDA, "tdnn_gemm_kernel launch: %s", cudaGetErrorString(e));
  return XVEC_OK;
}

int gemm_dispatch(const void* x, int x_dtype, int64_t x_rows, int cin, int64_t x_ld, const void* w_packed, int n,
                  const int32_t* tap_offsets, int taps, const float* bias, const float* scale, const float* shift, int relu,
                  void* y, int y_dtype, int64_t y_ld, const int32_t* row_utt, const int32_t* blk_slot_base, float* part,
                  int64_t rows, bool pool, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  if (!x || !w_packed || (!pool && !y) || (pool && (!row_utt || !blk_slot_base || !part)))
    return set_error(XVEC_E_ARG, "null pointer argument");
  if (x_dtype != XVEC_F32 && x_dtype != XVEC_BF16) return set_error(XVEC_E_ARG, "bad x_dtype %d", x_dtype);
  if (!pool && y_dtype != XVEC_F32 && y_dtype != XVEC_BF16) return set_error(XVEC_E_ARG, "bad y_dtype %d", y_dtype);
  if (taps < 1 || taps > XVEC_MAX_TAPS) return set_error(XVEC_E_ARG, "taps must be in [1,%d]", XVEC_MAX_TAPS);
  if (rows <= 0 || x_rows <= 0 || cin <= 0 || n <= 0) return set_error(XVEC_E_ARG, "non-positive size");
  if (rows > 0x7fffff00LL || x_rows > 0x7fffff00LL) return set_error(XVEC_E_ARG, "too many rows");
  if ((scale == nullptr) != (shift == nullptr)) return set_error(XVEC_E_ARG, "bn_scale and bn_shift must be given together");
  const bool tf32 = x_dtype == XVEC_F32;
  const int bke = tf32 ? 32 : 64;
  GemmParams p{};
  p.rows = static_cast<int>(rows);
  p.n = n;
  p.m_tiles = static_cast<int>((rows + BM - 1) / BM);
  p.n_tiles = (n + BN - 1) / BN;
  p.taps = taps;
  p.cpt = (cin + bke - 1) / bke;
  for (int j = 0; j < taps; ++j) {
    if (tap_offsets[j] < 0) return set_error(XVEC_E_ARG, "tap offsets must be non-negative");
    p.tap_off[j] = tap_offsets[j];
  }
  if (!pool && p.n_tiles * BN > MAX_NPAD) return set_error(XVEC_E_ARG, "n > %d is not supported by the store epilogue", MAX_NPAD);
  p.bias = bias;
  p.scale = scale;
  p.shift = shift;
  p.relu = relu;
  p.out = y;
  p.ldo = y_ld;
  const int64_t yes = y_dtype == XVEC_BF16 ? 2 : 4;
  p.vec_ok = (!pool && (reinterpret_cast<uintptr_t>(y) & 15u) == 0 && (y_ld * yes) % 16 == 0) ? 1 : 0;
  if (!pool && y_ld < n) return set_error(XVEC_E_ARG, "y_ld < n");
  p.row_utt = row_utt;
  p.blk_slot_base = blk_slot_base;
  p.part = part;

  CUtensorMap ta, tb;
  rc = make_tmap_2d(&ta, x, x_dtype, static_cast<uint64_t>(cin), static_cast<uint64_t>(x_rows), static_cast<uint64_t>(x_ld), bke, BM);
  if (rc) return rc;
  const uint64_t kpad = static_cast<uint64_t>(taps) * p.cpt * bke;
  rc = make_tmap_2d(&tb, w_packed, x_dtype, kpad, static_cast<uint64_t>(p.n_tiles) * BN, kpad, bke, BN);
  if (rc) return rc;

  const int64_t tiles = static_cast<int64_t>(p.m_tiles) * p.n_tiles;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pool) return tf32 ? launch<true, EPI_POOL>(ta, tb, p, grid, st) : launch<false, EPI_POOL>(ta, tb, p, grid, st);
  if (y_dtype == XVEC_BF16)
    return tf32 ? launch<true, EPI_STORE_BF16>(ta, tb, p, grid, st) : launch<false, EPI_STORE_BF16>(ta, tb, p, grid, st);
  return tf32 ? launch<true, EPI_STORE_F32>(ta, tb, p, grid, st) : launch<false, EPI_STORE_F32>(ta, tb, p, grid, st);
}

int read_watchdog() {
  unsigned int v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_watchdog_code, sizeof(v));
  if (e != cudaSuccess) return set_error(XVEC_E_CUDA, "cudaMemcpyFromSymbol: %s", cudaGetErrorString(e));
  return static_cast<int>(v);
}

}  // namespace xvec
