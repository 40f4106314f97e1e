// Kaldi-compatible log-mel filter bank features on the GPU (the step immediately before the hot path).
//
// Reference: AudioFeaturizer.featurize / _compute_fbank (ppasr/data_utils/featurizer/audio_featurizer.py:37-69,120-138):
// -20 dB RMS normalisation (data_utils/audio.py:287-304), conversion to int16 scale (audio.py:549-574), then
// paddleaudio.compliance.kaldi.fbank(waveform, n_mels, frame_length=25, frame_shift=10, dither=0, sr) -- a third-party port of
// Kaldi's compute-fbank-feats with the defaults snip_edges, remove_dc_offset, preemphasis 0.97, povey window, round to power
// of two (512), power spectrum, low_freq 20, high_freq = Nyquist, log(max(e, FLT_EPSILON)); restated in oracle/fbank_oracle.py
// and checked there against torchaudio.compliance.kaldi.fbank.
//
// One CTA (128 threads) per frame: 400 samples -> DC removal -> pre-emphasis -> povey window -> 512-point radix-2 FFT in
// shared memory (fp32) -> power spectrum -> sparse triangular mel filters -> log. ~25 kFLOP and 1.6 KB read per frame:
// HBM / latency bound, fp32 on CUDA cores on purpose (bf16 tensor-core DFTs cannot hold a speech frame's dynamic range).
#include <cmath>
#include <mutex>
#include <vector>

#include "common.h"
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace ppasr {

void count_launch();

constexpr int FB_WIN = 400, FB_SHIFT = 160, FB_NFFT = 512, FB_THREADS = 128, FB_MAXW = 2048;

struct FbankTables {
  float* window = nullptr;   // [400]
  float2* twiddle = nullptr; // [256] exp(-2 pi i k / 512)
  int* mel_start = nullptr;  // [n_mels] first fft bin with a non-zero weight
  int* mel_ofs = nullptr;    // [n_mels + 1] offsets into mel_w
  float* mel_w = nullptr;    // packed non-zero weights
  int n_mels = 0, sample_rate = 0, nw = 0;
};

// per-utterance gain of AudioSegment.normalize(target_db): 10^((target_db - rms_db) / 20), rms_db = 10 log10(mean(x^2))
__global__ void fbank_gain_kernel(const float* __restrict__ audio, long long stride, const int* __restrict__ nsamp, int N,
                                  float target_db, float* __restrict__ gain) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  const int n = nsamp ? nsamp[b] : N;
  const float* x = audio + (size_t)b * stride;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i] * (double)x[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    const double ms = n > 0 ? t / n : 0.0;
    const double rms_db = 10.0 * log10(ms == 0.0 ? 1.0 : ms);  // audio.py:526-529: an all-zero segment counts as 0 dB
    gain[b] = (float)pow(10.0, ((double)target_db - rms_db) / 20.0);
  }
}

__global__ void __launch_bounds__(FB_THREADS)
fbank_kernel(const float* __restrict__ audio, long long stride, const int* __restrict__ nsamp, int N, const float* __restrict__ gain,
             const FbankTables tb, float* __restrict__ out, int Tmax) {
  __shared__ float2 buf[FB_NFFT];
  __shared__ float xs[FB_WIN];
  __shared__ float2 tw[FB_NFFT / 2];
  __shared__ float pw[FB_NFFT / 2 + 1];
  __shared__ float red[FB_THREADS / 32];
  const int b = blockIdx.y, m = blockIdx.x, tid = threadIdx.x;
  const int n = nsamp ? nsamp[b] : N;
  const int frames = n >= FB_WIN ? 1 + (n - FB_WIN) / FB_SHIFT : 0;
  float* orow = out + ((size_t)b * Tmax + m) * tb.n_mels;
  if (m >= frames) {  // padding frame of a shorter utterance
    for (int i = tid; i < tb.n_mels; i += FB_THREADS) orow[i] = 0.f;
    return;
  }
  const float g = gain ? gain[b] * 32768.0f : 32768.0f;
  const float* x = audio + (size_t)b * stride + (size_t)m * FB_SHIFT;
  // int16 conversion of audio.py:549-574: scale, clip, truncate toward zero
  float lsum = 0.f;
  for (int i = tid; i < FB_WIN; i += FB_THREADS) {
    const float v = truncf(fminf(fmaxf(x[i] * g, -32768.f), 32767.f));
    xs[i] = v;
    lsum += v;
  }
  for (int i = tid; i < FB_NFFT / 2; i += FB_THREADS) tw[i] = tb.twiddle[i];
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0) red[tid >> 5] = lsum;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int w = 0; w < FB_THREADS / 32; ++w) mean += red[w];
  mean *= (1.0f / FB_WIN);
  // DC removal, pre-emphasis (x[i] - 0.97 x[max(i-1,0)]), povey window, bit-reversed load into the FFT buffer
  for (int i = tid; i < FB_NFFT; i += FB_THREADS) {
    float v = 0.f;
    if (i < FB_WIN) {
      const float c = xs[i] - mean;
      const float p = xs[i > 0 ? i - 1 : 0] - mean;
      v = (c - 0.97f * p) * tb.window[i];
    }
    buf[__brev((unsigned)i) >> 23] = make_float2(v, 0.f);  // 9-bit reversal
  }
  __syncthreads();
#pragma unroll 1
  for (int s = 0; s < 9; ++s) {
    const int half = 1 << s;
    for (int idx = tid; idx < FB_NFFT / 2; idx += FB_THREADS) {
      const int pos = idx & (half - 1);
      const int i = ((idx >> s) << (s + 1)) + pos;
      const int j = i + half;
      const float2 w = tw[pos << (8 - s)];
      const float2 a = buf[i], c = buf[j];
      const float2 t = make_float2(w.x * c.x - w.y * c.y, w.x * c.y + w.y * c.x);
      buf[i] = make_float2(a.x + t.x, a.y + t.y);
      buf[j] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
  for (int k = tid; k <= FB_NFFT / 2; k += FB_THREADS) pw[k] = buf[k].x * buf[k].x + buf[k].y * buf[k].y;
  __syncthreads();
  for (int bin = tid; bin < tb.n_mels; bin += FB_THREADS) {
    const int k0 = tb.mel_start[bin], o0 = tb.mel_ofs[bin], cnt = tb.mel_ofs[bin + 1] - o0;
    float e = 0.f;
    for (int q = 0; q < cnt; ++q) e = fmaf(tb.mel_w[o0 + q], pw[k0 + q], e);
    orow[bin] = logf(fmaxf(e, 1.1920928955078125e-07f));
  }
}

static FbankTables g_tables;
static std::mutex g_tables_mu;

static int build_tables(int n_mels, int sample_rate) {
  std::lock_guard<std::mutex> lk(g_tables_mu);
  if (g_tables.window && g_tables.n_mels == n_mels && g_tables.sample_rate == sample_rate) return PPASR_OK;
  std::vector<float> win(FB_WIN);
  for (int i = 0; i < FB_WIN; ++i)  // povey: hann(periodic = false) ^ 0.85
    win[i] = std::pow(0.5f - 0.5f * std::cos(2.0 * M_PI * i / (FB_WIN - 1)), 0.85);
  std::vector<float2> twd(FB_NFFT / 2);
  for (int k = 0; k < FB_NFFT / 2; ++k) twd[k] = make_float2((float)std::cos(-2.0 * M_PI * k / FB_NFFT), (float)std::sin(-2.0 * M_PI * k / FB_NFFT));
  // get_mel_banks: triangular filters on the mel scale 1127 ln(1 + f / 700), low 20 Hz, high = Nyquist, fp32 like the reference
  auto mel = [](float f) { return 1127.0f * std::log(1.0f + f / 700.0f); };
  const float nyq = 0.5f * sample_rate;
  const float bw = (float)sample_rate / FB_NFFT;
  const float mlo = mel(20.0f), mhi = mel(nyq);
  const float delta = (mhi - mlo) / (n_mels + 1);
  std::vector<int> start(n_mels), ofs(n_mels + 1);
  std::vector<float> w;
  for (int b = 0; b < n_mels; ++b) {
    const float left = mlo + b * delta, center = mlo + (b + 1.0f) * delta, right = mlo + (b + 2.0f) * delta;
    int first = -1, last = -1;
    std::vector<float> row(FB_NFFT / 2);
    for (int k = 0; k < FB_NFFT / 2; ++k) {
      const float mk = mel(bw * k);
      const float up = (mk - left) / (center - left), down = (right - mk) / (right - center);
      const float v = std::max(0.0f, std::min(up, down));
      row[k] = v;
      if (v > 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    if (first < 0) first = last = 0;
    start[b] = first;
    ofs[b] = (int)w.size();
    for (int k = first; k <= last; ++k) w.push_back(row[k]);
  }
  ofs[n_mels] = (int)w.size();
  auto up = [](const void* src, size_t bytes, void** dst) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  FbankTables t;
  PPASR_CUDA_CHECK(up(win.data(), win.size() * 4, (void**)&t.window));
  PPASR_CUDA_CHECK(up(twd.data(), twd.size() * 8, (void**)&t.twiddle));
  PPASR_CUDA_CHECK(up(start.data(), start.size() * 4, (void**)&t.mel_start));
  PPASR_CUDA_CHECK(up(ofs.data(), ofs.size() * 4, (void**)&t.mel_ofs));
  PPASR_CUDA_CHECK(up(w.data(), w.size() * 4, (void**)&t.mel_w));
  t.n_mels = n_mels, t.sample_rate = sample_rate, t.nw = (int)w.size();
  g_tables = t;  // (tables of a previous configuration are leaked on purpose: a few KB, other streams may still read them)
  return PPASR_OK;
}

}  // namespace ppasr

using namespace ppasr;

extern "C" {

int ppasr_b200_fbank_frames(int32_t n_samples) { return n_samples >= FB_WIN ? 1 + (n_samples - FB_WIN) / FB_SHIFT : 0; }

int ppasr_b200_fbank(const float* audio, int32_t B, int64_t stride, int32_t N, const int32_t* n_samples, int32_t n_mels,
                     int32_t sample_rate, int32_t db_normalize, float target_db, float* gain_ws, float* out, int32_t Tmax,
                     void* stream) {
  PPASR_REQUIRE(audio && out && B > 0 && N > 0 && stride >= N, "bad arguments");
  PPASR_REQUIRE(n_mels > 3 && n_mels <= 256 && sample_rate == 16000, "fbank: n_mels in (3, 256], 16 kHz audio (25 ms / 10 ms frames)");
  PPASR_REQUIRE(!db_normalize || gain_ws, "db_normalize needs a [B] float workspace");
  PPASR_REQUIRE(Tmax >= ppasr_b200_fbank_frames(N), "Tmax too small");
  if (Tmax == 0) return PPASR_OK;
  int rc = build_tables(n_mels, sample_rate);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (db_normalize) {
    fbank_gain_kernel<<<B, 1024, 0, st>>>(audio, (long long)stride, n_samples, N, target_db, gain_ws);
    count_launch();
  }
  fbank_kernel<<<dim3(Tmax, B), FB_THREADS, 0, st>>>(audio, (long long)stride, n_samples, N, db_normalize ? gain_ws : nullptr,
                                                    g_tables, out, Tmax);
  count_launch();
  PPASR_CUDA_CHECK(cudaGetLastError());
  return PPASR_OK;
}

}  // extern "C"
