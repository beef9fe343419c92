// Statistics-pooling finalize FUSED with the first segment layer — one launch for the tail of a batch (sm_100a build).
//
//   out[u, :] = act( [mean(u) || std(u)] . W' + bias ),   mean / std of utterance u from the pooling partials of the TDNN5 epilogue
//   replaces: torch.mean / torch.std / torch.cat in stat_pool (main.py:59-63) + segment_layer6 (main.py:45, 87-90)
//
// MEASURED RESULT (B200, 256 utterances, p = 1500, n = 512, warm L2, ncu --cache-control none): 37.9 us against 7.9 + 23.3 us for
// pool_finalize_kernel + fc_small_kernel in bf16 (54.8 against 7.9 + 41.8 in TF32).  It saves a launch and never writes the
// pooled matrix, but a 16-utterance CTA re-reads its K slice of W (48 MB of the 89 MB of L2 traffic; the pair moves 55 MB) and the
// n-halves each finalize the same statistics.  xvec_extract_forward therefore keeps the two launches; this stays an alternative
// entry point (xvec_pool_fc_fused) with its own test.
//
// The unfused tail is pool_finalize_kernel + fc_small_kernel (94 dependent K steps per CTA, bound by L2
// latency).  Here the K dimension (2p = 3000 statistics) is split over
// the grid: CTA (slice, n-half, 16 utterances) finalizes ITS 2 x cs statistics of its 16 utterances straight into shared memory
// (fixed-order float64 reduction over the partial slots, exactly pool_finalize_kernel's arithmetic; the (n_utts, 2p) pooled matrix
// never exists), multiplies them with its K slice of W (24 K steps, weight fragments read from L2 straight into the mma.sync
// register layout, three steps in flight) and writes a float32 partial tile.  The last CTA of a tile to arrive (one atomic per
// CTA) adds the slices IN SLICE ORDER, applies bias / ReLU and stores the result — deterministic, no second launch; it leaves the
// arrival counter at zero for the next launch.
//
// Like fc_small.cu this is the legacy mma.sync tensor path on purpose: 128 threads, < 14 KiB of shared memory, no TMEM, so a CTA
// fits NEXT TO a resident CTA of the next batch's persistent tdnn_stack_kernel (tdnn_stack.cu: STACK_SMEM_LEFT_FOR_TAIL).
// Operands are addressed in BYTES: a K step is 32 bytes of a row (16 bf16 / 8 float32), m16n8k16.bf16 and m16n8k8.tf32 have the
// same fragment byte layout (thread (g, t): words at byte 4t and 4t + 16 of row g).
#include <cuda_bf16.h>
#include <stdint.h>

#include "xvec_internal.h"

namespace xvec {

constexpr int PF_M = 16;              // utterances per CTA
constexpr int PF_N = 256;             // output columns per CTA: 4 warps x 64
constexpr int PF_THREADS = 128;
constexpr int PF_SLICE_BYTES = 768;   // K bytes per slice and half ([means | stds] -> 2 x 384): cs = 192 bf16 / 96 float32 columns
constexpr int PF_HALF_BYTES = PF_SLICE_BYTES / 2;
constexpr int PF_STEPS = PF_SLICE_BYTES / 32;
constexpr int PF_PITCH = PF_SLICE_BYTES + 16;  // shared-memory row pitch: ldmatrix rows land in distinct 16-byte bank groups
constexpr int PF_DEPTH = 3;           // K steps of weight fragments in flight per thread
static_assert(PF_M * PF_PITCH + 256 <= 20 * 1024, "must fit next to a resident stack CTA");

struct PoolFcParams {
  const float* part;
  const int* slot_start;
  const int* n_rows;
  int n_utts, p;
  const float* scale;
  const float* shift;
  const uint8_t* w;       // (n, 2p) row-major in the operand dtype
  long long ldw_bytes;
  const float* bias;
  int n, relu;
  void* out;
  long long ldo;          // elements
  float* ws;              // [ksplit][m_tiles * PF_M][n_pad] float32 partial tiles
  unsigned* counters;     // [m_tiles * n_halves], zero on entry, zero on exit
  int ksplit, cs, es;     // slices, statistics columns per slice, bytes per operand element
  int n_pad, rows_pad;
};

__device__ __forceinline__ void pf_ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
template <bool kTf32>
__device__ __forceinline__ void pf_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (kTf32)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool kTf32, bool kOutBf16>
__global__ void __launch_bounds__(PF_THREADS, 4)  // <= 128 registers: 16 K per CTA, what a resident stack CTA (352 x 136) leaves free
pool_fc_kernel(const PoolFcParams p) {
  __shared__ __align__(16) uint8_t xs[PF_M * PF_PITCH];
  __shared__ double inv_n[PF_M][2];
  __shared__ int s_nrows[PF_M], s_slot0[PF_M], s_slot1[PF_M];
  __shared__ unsigned s_last;
  const int slice = blockIdx.x, nh = blockIdx.y, mt = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u0 = mt * PF_M;
  const int c0 = slice * p.cs;  // first statistics column of this slice

  // ---------------------------------------------------------------- phase 1: finalize 16 utterances x cs columns into shared memory
  if (tid < PF_M) {
    const int u = u0 + tid;
    const int n0 = u < p.n_utts ? p.n_rows[u] : 0;
    s_nrows[tid] = n0;
    s_slot0[tid] = u < p.n_utts ? p.slot_start[u] : 0;
    s_slot1[tid] = u < p.n_utts ? p.slot_start[u + 1] : 0;
    inv_n[tid][0] = n0 > 0 ? 1.0 / n0 : 0.0;
    inv_n[tid][1] = n0 > 1 ? 1.0 / (n0 - 1) : 0.0;
  }
  __syncthreads();
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  // a work item = 4 adjacent statistics columns of one utterance: float4 loads, the sums of up to four slots (8 loads) in flight
  // per thread — the phase is bound by the latency of these loads, so width matters more than arithmetic
  const int c4 = p.cs >> 2;
  for (int idx = tid; idx < PF_M * c4; idx += PF_THREADS) {
    const int ul = idx / c4, cl = (idx - ul * c4) << 2;
    const int col = c0 + cl;
    float m[4] = {0.f, 0.f, 0.f, 0.f}, sd[4] = {0.f, 0.f, 0.f, 0.f};
    if (u0 + ul < p.n_utts && col < p.p) {
      // same arithmetic as pool_finalize_kernel (pool.cu): slots in order, float64, unbiased variance, NaN for a single frame
      double S[4] = {0.0, 0.0, 0.0, 0.0}, Q[4] = {0.0, 0.0, 0.0, 0.0};
      const int sl1 = s_slot1[ul];
      int sl = s_slot0[ul];
      const bool vec = col + 4 <= p.p;  // whole float4 inside the row (p is a multiple of 4 on this path, so always)
      for (; sl < sl1; sl += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* src = p.part + static_cast<size_t>(min(sl + i, sl1 - 1)) * 2 * p.p + col;
          if (vec) {
            a[i] = __ldcg(reinterpret_cast<const float4*>(src));
            b[i] = __ldcg(reinterpret_cast<const float4*>(src + p.p));
          } else {
            a[i] = make_float4(src[0], col + 1 < p.p ? src[1] : 0.f, col + 2 < p.p ? src[2] : 0.f, 0.f);
            b[i] = make_float4(src[p.p], col + 1 < p.p ? src[p.p + 1] : 0.f, col + 2 < p.p ? src[p.p + 2] : 0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (sl + i >= sl1) break;  // the clamped loads past the last slot are not added
          S[0] += static_cast<double>(a[i].x); S[1] += static_cast<double>(a[i].y); S[2] += static_cast<double>(a[i].z); S[3] += static_cast<double>(a[i].w);
          Q[0] += static_cast<double>(b[i].x); Q[1] += static_cast<double>(b[i].y); Q[2] += static_cast<double>(b[i].z); Q[3] += static_cast<double>(b[i].w);
        }
      }
      const int n = s_nrows[ul];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (col + e >= p.p) break;
        const double mean = n > 0 ? S[e] * inv_n[ul][0] : nan;
        double var = n > 1 ? (Q[e] - S[e] * S[e] * inv_n[ul][0]) * inv_n[ul][1] : nan;
        if (var < 0.0) var = 0.0;
        const float sc = p.scale ? p.scale[col + e] : 1.f;
        const float sh = p.shift ? p.shift[col + e] : 0.f;
        m[e] = static_cast<float>(mean * sc + sh);
        sd[e] = fabsf(sc) * sqrtf(static_cast<float>(var));
      }
    }
    uint8_t* row = xs + ul * PF_PITCH;
    if constexpr (kTf32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(row) + cl) = make_float4(m[0], m[1], m[2], m[3]);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(row + PF_HALF_BYTES) + cl) = make_float4(sd[0], sd[1], sd[2], sd[3]);
    } else {
      __nv_bfloat162* dm = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(row) + cl);
      __nv_bfloat162* ds = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(row + PF_HALF_BYTES) + cl);
      dm[0] = __floats2bfloat162_rn(m[0], m[1]);
      dm[1] = __floats2bfloat162_rn(m[2], m[3]);
      ds[0] = __floats2bfloat162_rn(sd[0], sd[1]);
      ds[1] = __floats2bfloat162_rn(sd[2], sd[3]);
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2: (16 x 768 bytes of K) x (64 columns per warp)
  const int g = lane >> 2, t = lane & 3;
  const int ncol0 = nh * PF_N + warp * 64;  // this warp's first output column
  // weight fragments straight from global memory (L2): thread (g, t) of n8-tile j needs the 32-bit words at byte 4t and 4t + 16
  // of the 32-byte K block of row ncol0 + 8j + g.  Words whose statistics column is >= p (padding of the last slice) read as 0.
  const uint8_t* wrow[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int nr = ncol0 + 8 * j + g;
    wrow[j] = nr < p.n ? p.w + static_cast<long long>(nr) * p.ldw_bytes : nullptr;
  }
  auto fetch = [&](int s, uint32_t (&b)[8][2]) {
    // step s covers bytes [32 s, 32 s + 32) of the slice: first half = means (W columns c0 ..), second half = stds (W columns p + c0 ..)
    const bool second = s >= PF_STEPS / 2;
    const int off = (second ? s - PF_STEPS / 2 : s) * 32;                       // byte offset inside the half
    const long long base = (static_cast<long long>(second ? p.p : 0) + c0) * p.es + off;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int byte = 4 * t + 16 * h;
      const int colw = c0 + (off + byte) / p.es;  // first statistics column of this word (p is even: a bf16 pair is in or out as one)
      const bool ok = s < PF_STEPS && colw < p.p;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        b[j][h] = (ok && wrow[j]) ? __ldg(reinterpret_cast<const uint32_t*>(wrow[j] + base + byte)) : 0u;
    }
  };
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  uint32_t bq[PF_DEPTH][8][2];
#pragma unroll
  for (int d = 0; d < PF_DEPTH; ++d) fetch(d, bq[d]);
  for (int s0 = 0; s0 < PF_STEPS; s0 += PF_DEPTH) {
#pragma unroll
    for (int d = 0; d < PF_DEPTH; ++d) {
      const int s = s0 + d;
      if (s >= PF_STEPS) break;
      uint32_t a[4];
      pf_ldmatrix_x4(a, &xs[((lane & 7) + 8 * ((lane >> 3) & 1)) * PF_PITCH + 32 * s + 16 * (lane >> 4)]);
#pragma unroll
      for (int j = 0; j < 8; ++j) pf_mma<kTf32>(acc[j], a, bq[d][j][0], bq[d][j][1]);
      fetch(s + PF_DEPTH, bq[d]);  // past the end: zeros, no access
    }
  }

  // ---------------------------------------------------------------- phase 3: partial tile out; the last slice to arrive reduces
  float* tile = p.ws + (static_cast<size_t>(slice) * p.rows_pad + u0) * p.n_pad + ncol0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = 8 * j + 2 * t;
    *reinterpret_cast<float2*>(tile + static_cast<size_t>(g) * p.n_pad + col) = make_float2(acc[j][0], acc[j][1]);
    *reinterpret_cast<float2*>(tile + static_cast<size_t>(g + 8) * p.n_pad + col) = make_float2(acc[j][2], acc[j][3]);
  }
  __threadfence();
  __syncthreads();
  unsigned* counter = p.counters + mt * gridDim.y + nh;
  if (tid == 0) s_last = atomicAdd(counter, 1u) == static_cast<unsigned>(p.ksplit) - 1u ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();  // the other slices' tiles are visible (each fenced before its arrival)
  // slice order (the result does not depend on which CTA arrived last); all 16 positions of a thread per slice in one batch of
  // independent loads — a loop over the slices per position would serialise 128 L2 round trips
  float2 v[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j][0] = v[j][1] = make_float2(0.f, 0.f);
  const float* wbase = p.ws + static_cast<size_t>(u0 + g) * p.n_pad + ncol0 + 2 * t;
  for (int s = 0; s < p.ksplit; ++s) {
    const float* sp = wbase + static_cast<size_t>(s) * p.rows_pad * p.n_pad;
    float2 q[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      q[j][0] = __ldcg(reinterpret_cast<const float2*>(sp + 8 * j));
      q[j][1] = __ldcg(reinterpret_cast<const float2*>(sp + static_cast<size_t>(8) * p.n_pad + 8 * j));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        v[j][h].x += q[j][h].x;
        v[j][h].y += q[j][h].y;
      }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = ncol0 + 8 * j + 2 * t;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = u0 + g + 8 * h;
      if (row >= p.n_utts || col >= p.n) continue;
      float2 r = v[j][h];
      if (p.bias) {
        r.x += p.bias[col];
        if (col + 1 < p.n) r.y += p.bias[col + 1];
      }
      if (p.relu) {  // torch.relu keeps NaN (a single-frame utterance's std): (v < 0 ? 0 : v), not fmaxf
        r.x = r.x < 0.f ? 0.f : r.x;
        r.y = r.y < 0.f ? 0.f : r.y;
      }
      if constexpr (kOutBf16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ldo + col;
        o[0] = __float2bfloat16_rn(r.x);
        if (col + 1 < p.n) o[1] = __float2bfloat16_rn(r.y);
      } else {
        float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col;
        o[0] = r.x;
        if (col + 1 < p.n) o[1] = r.y;
      }
    }
  }
  if (tid == 0) *counter = 0u;  // every slice has arrived: ready for the next launch
}

static int pf_cs(int dtype) { return PF_HALF_BYTES / (dtype == XVEC_BF16 ? 2 : 4); }

bool pool_fc_supported(int n_utts, int p, int n, int dtype, int64_t w_ld, const void* w) {
  const int es = dtype == XVEC_BF16 ? 2 : 4;
  return n_utts > 0 && p > 0 && n > 0 && n % 2 == 0 && p % 4 == 0 && w_ld >= 2LL * p && (w_ld * es) % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(w) & 3u) == 0 && (dtype == XVEC_BF16 || dtype == XVEC_F32) && n_utts <= 0x7fffff00 / PF_M;
}

int64_t pool_fc_workspace_bytes(int n_utts, int p, int n, int dtype) {
  if (n_utts <= 0 || p <= 0 || n <= 0) return 0;
  const int cs = pf_cs(dtype);
  const int64_t ksplit = (p + cs - 1) / cs, m_tiles = (n_utts + PF_M - 1) / PF_M, n_halves = (n + PF_N - 1) / PF_N;
  const int64_t counters = (m_tiles * n_halves * 4 + 255) / 256 * 256;
  return counters + ksplit * m_tiles * PF_M * n_halves * PF_N * 4;
}

int pool_fc_dispatch(const float* part, const int32_t* slot_start, const int32_t* n_rows, int n_utts, int p, const float* scale,
                     const float* shift, const void* w, int dtype, int64_t w_ld, const float* bias, int n, int relu, void* out,
                     int out_dtype, int64_t out_ld, void* ws, int64_t ws_bytes, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  if (!part || !slot_start || !n_rows || !w || !out || !ws) return set_error(XVEC_E_ARG, "null pointer argument");
  if (!pool_fc_supported(n_utts, p, n, dtype, w_ld, w)) return set_error(XVEC_E_ARG, "shape / alignment not supported by the fused pooling + segment kernel");
  if ((scale == nullptr) != (shift == nullptr)) return set_error(XVEC_E_ARG, "bn_scale and bn_shift must be given together");
  if (out_dtype != XVEC_F32 && out_dtype != XVEC_BF16) return set_error(XVEC_E_ARG, "bad out_dtype %d", out_dtype);
  if (out_ld < n) return set_error(XVEC_E_ARG, "out_ld < n");
  if (ws_bytes < pool_fc_workspace_bytes(n_utts, p, n, dtype) || (reinterpret_cast<uintptr_t>(ws) & 15u))
    return set_error(XVEC_E_ARG, "workspace must be 16-byte aligned and hold xvec_pool_fc_workspace_bytes() bytes");
  PoolFcParams q{};
  q.part = part;
  q.slot_start = slot_start;
  q.n_rows = n_rows;
  q.n_utts = n_utts;
  q.p = p;
  q.scale = scale;
  q.shift = shift;
  q.w = static_cast<const uint8_t*>(w);
  q.es = dtype == XVEC_BF16 ? 2 : 4;
  q.ldw_bytes = w_ld * q.es;
  q.bias = bias;
  q.n = n;
  q.relu = relu;
  q.out = out;
  q.ldo = out_ld;
  q.cs = pf_cs(dtype);
  q.ksplit = (p + q.cs - 1) / q.cs;
  const int m_tiles = (n_utts + PF_M - 1) / PF_M, n_halves = (n + PF_N - 1) / PF_N;
  q.n_pad = n_halves * PF_N;
  q.rows_pad = m_tiles * PF_M;
  q.counters = static_cast<unsigned*>(ws);
  q.ws = reinterpret_cast<float*>(static_cast<char*>(ws) + (static_cast<int64_t>(m_tiles) * n_halves * 4 + 255) / 256 * 256);
  const dim3 grid(q.ksplit, n_halves, m_tiles);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == XVEC_BF16) {
    if (out_dtype == XVEC_BF16) pool_fc_kernel<false, true><<<grid, PF_THREADS, 0, st>>>(q);
    else pool_fc_kernel<false, false><<<grid, PF_THREADS, 0, st>>>(q);
  } else {
    if (out_dtype == XVEC_BF16) pool_fc_kernel<true, true><<<grid, PF_THREADS, 0, st>>>(q);
    else pool_fc_kernel<true, false><<<grid, PF_THREADS, 0, st>>>(q);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(XVEC_E_CUDA, "pool_fc_kernel launch: %s", cudaGetErrorString(e));
  return XVEC_OK;
}

}  // namespace xvec
