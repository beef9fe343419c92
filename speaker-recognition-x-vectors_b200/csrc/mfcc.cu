// MFCC front end on the GPU: waveform -> the (frames x 24) float32 matrix the TDNN path reads.
//
// replaces: python_speech_features.mfcc(sample, 16000, numcep=24, nfilt=26, nfft=512) at dataset.py:130 and the min-max
// normalisation of the signal before it (dataset.py:216-219).  Parameters are the reference's (fixed):
//   preemphasis 0.97, 400-sample frames every 160 samples (zero padded at the end, rectangular window), |rfft_512|^2/512,
//   26 triangular mel filters on integer bins (0..8 kHz), log, orthonormal DCT-II (first 24), lifter L=22,
//   coefficient 0 = log(frame energy).
// One warp per frame: the 512-point real FFT is a 256-point complex Stockham radix-4 FFT (4 passes through a 2 KiB shared
// ping-pong buffer) plus the real-input split; mel / log / DCT are tiny per-lane loops.  Latency/shared-memory bound, not a
// tensor-core problem: 306 k frames per 1024-utterance set are ~15 GFLOP.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "xvec_internal.h"

namespace xvec {

constexpr int MF_FRAME_LEN = 400;
constexpr int MF_FRAME_STEP = 160;
constexpr int MF_NFFT = 512;
constexpr int MF_NH = 256;     // complex FFT size
constexpr int MF_NBINS = 257;
constexpr int MF_NFILT = 26;
constexpr int MF_NUMCEP = 24;
constexpr int MF_WARPS = 4;
constexpr int MF_FRAMES_PER_CTA = 16;
constexpr float MF_EPS = 2.220446049250313e-16f;  // numpy float64 eps, what the reference's package substitutes for zeros

__constant__ int c_mel_bins[MF_NFILT + 2];
__constant__ float c_dct_lift[MF_NUMCEP * MF_NFILT];  // orthonormal DCT-II rows with the lifter folded in

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

template <bool kInt16>
__global__ void __launch_bounds__(MF_WARPS * 32)
mfcc_kernel(const void* __restrict__ wav, const long long* __restrict__ wav_start, const int* __restrict__ wav_len,
            const long long* __restrict__ row_start, const int* __restrict__ n_frames, const float* __restrict__ norm_offset,
            const float* __restrict__ norm_scale, float* __restrict__ out, long long ld) {
  __shared__ float2 tw256[MF_NH];       // exp(-2 pi i t / 256)
  __shared__ float2 tw512[MF_NBINS];    // exp(-2 pi i k / 512)
  __shared__ float2 bufs[MF_WARPS][2][MF_NH];
  __shared__ float pspec[MF_WARPS][MF_NBINS + 7];
  __shared__ float logfb[MF_WARPS][32];

  const int u = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < MF_NH; t += blockDim.x) {
    float s, c;
    sincospif(-2.0f * t / MF_NH, &s, &c);
    tw256[t] = make_float2(c, s);
  }
  for (int t = threadIdx.x; t < MF_NBINS; t += blockDim.x) {
    float s, c;
    sincospif(-2.0f * t / MF_NFFT, &s, &c);
    tw512[t] = make_float2(c, s);
  }
  __syncthreads();

  const int nf = n_frames[u];
  const int len = wav_len[u];
  const long long w0 = wav_start[u];
  const float off = norm_offset ? norm_offset[u] : 0.f;
  const float scl = norm_scale ? norm_scale[u] : 1.f;
  auto sample = [&](int idx) -> float {  // normalised signal value, 0 outside [0, len)
    if (idx < 0 || idx >= len) return 0.f;
    float v;
    if constexpr (kInt16) v = static_cast<float>(reinterpret_cast<const short*>(wav)[w0 + idx]);
    else v = reinterpret_cast<const float*>(wav)[w0 + idx];
    return (v + off) * scl;
  };
  auto emph = [&](int idx, int frame_pos) -> float {  // pre-emphasised, zero beyond the signal and beyond the 400-sample frame
    if (frame_pos >= MF_FRAME_LEN || idx >= len) return 0.f;
    return sample(idx) - 0.97f * sample(idx - 1);
  };

  for (int fi = warp; fi < MF_FRAMES_PER_CTA; fi += MF_WARPS) {
    const int f = blockIdx.x * MF_FRAMES_PER_CTA + fi;
    if (f >= nf) break;  // warp-uniform
    float2* a = bufs[warp][0];
    float2* b = bufs[warp][1];
    // z[n] = y[2n] + i y[2n+1]
    for (int n = lane; n < MF_NH; n += 32) {
      const int p0 = 2 * n;
      a[n] = make_float2(emph(f * MF_FRAME_STEP + p0, p0), emph(f * MF_FRAME_STEP + p0 + 1, p0 + 1));
    }
    __syncwarp();
    // Stockham radix-4, natural-order output
#pragma unroll
    for (int ns = 1; ns < MF_NH; ns *= 4) {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj;        // butterfly 0..63
        const int k = j & (ns - 1);
        const int tstep = (MF_NH / 4 / ns) * k;  // twiddle index of exp(-2 pi i k / (4 ns))
        float2 v0 = a[j];
        float2 v1 = cmul(a[j + 64], tw256[tstep]);
        float2 v2 = cmul(a[j + 128], tw256[2 * tstep]);
        float2 v3 = cmul(a[j + 192], tw256[3 * tstep]);
        const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
        const float2 md = make_float2(d13.y, -d13.x);  // -i * (v1 - v3)
        const int j0 = ((j - k) << 2) + k;             // (j / ns) * 4 ns + k
        b[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
        b[j0 + ns] = make_float2(d02.x + md.x, d02.y + md.y);
        b[j0 + 2 * ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
        b[j0 + 3 * ns] = make_float2(d02.x - md.x, d02.y - md.y);
      }
      __syncwarp();
      float2* t = a; a = b; b = t;
    }
    // real-input split: X[k] = E + W512^k O,  E = (Z[k] + conj Z[N-k]) / 2,  O = -i (Z[k] - conj Z[N-k]) / 2
    float esum = 0.f;
    for (int k = lane; k < MF_NBINS; k += 32) {
      const float2 zk = a[k & (MF_NH - 1)];
      const float2 zc = a[(MF_NH - k) & (MF_NH - 1)];
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y));
      const float2 o = make_float2(d.y, -d.x);
      const float2 wo = cmul(tw512[k], o);
      const float xr = e.x + wo.x, xi = e.y + wo.y;
      const float pw = (xr * xr + xi * xi) * (1.0f / MF_NFFT);
      pspec[warp][k] = pw;
      esum += pw;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
    __syncwarp();
    if (lane < MF_NFILT) {
      const int b0 = c_mel_bins[lane], b1 = c_mel_bins[lane + 1], b2 = c_mel_bins[lane + 2];
      float acc = 0.f;
      for (int i = b0; i < b1; ++i) acc += pspec[warp][i] * (static_cast<float>(i - b0) / static_cast<float>(b1 - b0));
      for (int i = b1; i < b2; ++i) acc += pspec[warp][i] * (static_cast<float>(b2 - i) / static_cast<float>(b2 - b1));
      logfb[warp][lane] = logf(acc == 0.f ? MF_EPS : acc);
    }
    __syncwarp();
    if (lane < MF_NUMCEP) {
      float c = 0.f;
#pragma unroll
      for (int j = 0; j < MF_NFILT; ++j) c = fmaf(c_dct_lift[lane * MF_NFILT + j], logfb[warp][j], c);
      if (lane == 0) c = logf(esum == 0.f ? MF_EPS : esum);  // appendEnergy
      out[(row_start[u] + f) * ld + lane] = c;
    }
    __syncwarp();
  }
}

// per-utterance min / max -> (offset, scale) of the reference's normalisation: x -= min; x /= max(x)  (dataset.py:217-218)
template <bool kInt16>
__global__ void __launch_bounds__(256)
wav_minmax_kernel(const void* __restrict__ wav, const long long* __restrict__ wav_start, const int* __restrict__ wav_len,
                  float* __restrict__ norm_offset, float* __restrict__ norm_scale) {
  const int u = blockIdx.x;
  const int len = wav_len[u];
  const long long w0 = wav_start[u];
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    float v;
    if constexpr (kInt16) v = static_cast<float>(reinterpret_cast<const short*>(wav)[w0 + i]);
    else v = reinterpret_cast<const float*>(wav)[w0 + i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  __shared__ float smn[8], smx[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    norm_offset[u] = -mn;
    norm_scale[u] = 1.0f / (mx - mn);  // a constant signal gives inf/nan exactly like the reference's division by zero
  }
}

static int upload_tables() {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && done[dev]) return XVEC_OK;
  int bins[MF_NFILT + 2];
  const double sr = 16000.0;
  const double hi_mel = 2595.0 * log10(1.0 + (sr / 2.0) / 700.0);
  for (int i = 0; i < MF_NFILT + 2; ++i) {
    const double mel = hi_mel * i / (MF_NFILT + 1);
    const double hz = 700.0 * (pow(10.0, mel / 2595.0) - 1.0);
    bins[i] = static_cast<int>(floor((MF_NFFT + 1) * hz / sr));
  }
  float dct[MF_NUMCEP * MF_NFILT];
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < MF_NUMCEP; ++k) {
    const double lift = 1.0 + 11.0 * sin(pi * k / 22.0);
    const double norm = k == 0 ? sqrt(1.0 / (4.0 * MF_NFILT)) : sqrt(1.0 / (2.0 * MF_NFILT));
    for (int j = 0; j < MF_NFILT; ++j) dct[k * MF_NFILT + j] = static_cast<float>(2.0 * cos(pi * k * (2 * j + 1) / (2.0 * MF_NFILT)) * norm * lift);
  }
  cudaError_t e = cudaMemcpyToSymbol(c_mel_bins, bins, sizeof(bins));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_dct_lift, dct, sizeof(dct));
  if (e != cudaSuccess) return set_error(XVEC_E_CUDA, "cudaMemcpyToSymbol: %s", cudaGetErrorString(e));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return XVEC_OK;
}

}  // namespace xvec

using namespace xvec;

extern "C" {

int64_t xvec_mfcc_num_frames(int64_t n_samples) {
  if (n_samples <= MF_FRAME_LEN) return 1;
  return 1 + (n_samples - MF_FRAME_LEN + MF_FRAME_STEP - 1) / MF_FRAME_STEP;
}

int xvec_wav_minmax(const void* wav_dev, int wav_is_int16, const int64_t* wav_start_dev, const int32_t* wav_len_dev, int n_utts,
                    float* norm_offset_dev, float* norm_scale_dev, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  if (!wav_dev || !wav_start_dev || !wav_len_dev || !norm_offset_dev || !norm_scale_dev) return set_error(XVEC_E_ARG, "null pointer argument");
  if (n_utts <= 0) return set_error(XVEC_E_ARG, "bad n_utts");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (wav_is_int16)
    wav_minmax_kernel<true><<<n_utts, 256, 0, st>>>(wav_dev, reinterpret_cast<const long long*>(wav_start_dev), wav_len_dev, norm_offset_dev, norm_scale_dev);
  else
    wav_minmax_kernel<false><<<n_utts, 256, 0, st>>>(wav_dev, reinterpret_cast<const long long*>(wav_start_dev), wav_len_dev, norm_offset_dev, norm_scale_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(XVEC_E_CUDA, "wav_minmax_kernel launch: %s", cudaGetErrorString(e));
  return XVEC_OK;
}

int xvec_mfcc(const void* wav_dev, int wav_is_int16, const int64_t* wav_start_dev, const int32_t* wav_len_dev,
              const int64_t* row_start_dev, const int32_t* n_frames_dev, int n_utts, int max_frames, const float* norm_offset_dev,
              const float* norm_scale_dev, float* out_dev, int64_t out_ld, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  if (!wav_dev || !wav_start_dev || !wav_len_dev || !row_start_dev || !n_frames_dev || !out_dev) return set_error(XVEC_E_ARG, "null pointer argument");
  if (n_utts <= 0 || n_utts > 65535 || max_frames <= 0 || out_ld < MF_NUMCEP) return set_error(XVEC_E_ARG, "bad n_utts (1..65535) / max_frames / out_ld");
  if ((norm_offset_dev == nullptr) != (norm_scale_dev == nullptr)) return set_error(XVEC_E_ARG, "norm_offset and norm_scale must be given together");
  rc = upload_tables();
  if (rc) return rc;
  dim3 grid((max_frames + MF_FRAMES_PER_CTA - 1) / MF_FRAMES_PER_CTA, n_utts);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (wav_is_int16)
    mfcc_kernel<true><<<grid, MF_WARPS * 32, 0, st>>>(wav_dev, reinterpret_cast<const long long*>(wav_start_dev), wav_len_dev,
                                                      reinterpret_cast<const long long*>(row_start_dev), n_frames_dev, norm_offset_dev,
                                                      norm_scale_dev, out_dev, out_ld);
  else
    mfcc_kernel<false><<<grid, MF_WARPS * 32, 0, st>>>(wav_dev, reinterpret_cast<const long long*>(wav_start_dev), wav_len_dev,
                                                       reinterpret_cast<const long long*>(row_start_dev), n_frames_dev, norm_offset_dev,
                                                       norm_scale_dev, out_dev, out_ld);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(XVEC_E_CUDA, "mfcc_kernel launch: %s", cudaGetErrorString(e));
  return XVEC_OK;
}

}  // extern "C"
