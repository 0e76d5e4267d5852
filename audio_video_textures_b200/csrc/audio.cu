// (f3) Audio front end of the contrastive path: waveform -> log-mel spectrogram -> VGGish examples.
//   replaces: cvt/utils/mel_features.py (frame :22-46, periodic_hann :49-69, stft_magnitude :72-93,
//             log_mel_spectrogram :188-223) and the example framing of cvt/utils/vggish_utils.py:57-69.
// The reference computes everything in float64 numpy; the kernel does the same in fp64 (the work is tiny: 205k
// FMAs per 10 ms frame), so the fp32 outputs agree with the reference to rounding.  One CTA per STFT frame:
// window the samples into shared memory (zero padded to the FFT length), evaluate the fft_len/2+1 DFT bins
// directly from a shared twiddle table (index (k*n) mod fft_len), magnitudes, the mel matrix product and the log.
#include "common.cuh"

namespace {

constexpr int LM_THREADS = 256;

__global__ void __launch_bounds__(LM_THREADS)
logmel_kernel(const double *__restrict__ wave, int64_t n_samples, int channels, int win_len, int hop, int fft_len,
              const double *__restrict__ window, const double *__restrict__ mel, int n_mel, double log_offset,
              float *__restrict__ out) {
    extern __shared__ double lm_smem[];
    double *xw = lm_smem;                       // [fft_len]
    double *tc = xw + fft_len;                  // cos(2 pi i / fft_len)
    double *ts = tc + fft_len;                  // sin
    double *mag = ts + fft_len;                 // [fft_len / 2 + 1]
    const int64_t f = blockIdx.x;
    const int n_bins = fft_len / 2 + 1;
    for (int i = threadIdx.x; i < fft_len; i += LM_THREADS) {
        double v = 0.0;
        if (i < win_len) {
            const int64_t s = f * hop + i;
            if (channels == 1) v = wave[s];
            else {                              // mono = mean over channels (vggish_utils.py:42-43), data [n_samples, channels]
                double a = 0.0;
                for (int c = 0; c < channels; ++c) a += wave[s * channels + c];
                v = a / channels;
            }
            v *= window[i];
        }
        xw[i] = v;
        double sn, cs;
        sincospi(2.0 * i / fft_len, &sn, &cs);
        tc[i] = cs;
        ts[i] = sn;
    }
    __syncthreads();
    const int mask = fft_len - 1;               // fft_len is a power of two (mel_features.py:207)
    for (int k = threadIdx.x; k < n_bins; k += LM_THREADS) {
        double re = 0.0, im = 0.0;
        int idx = 0;
        for (int n = 0; n < win_len; ++n) {
            re = fma(xw[n], tc[idx], re);
            im = fma(-xw[n], ts[idx], im);
            idx = (idx + k) & mask;
        }
        mag[k] = sqrt(re * re + im * im);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n_mel; j += LM_THREADS) {
        double acc = 0.0;
        for (int k = 0; k < n_bins; ++k) acc = fma(mag[k], mel[(int64_t)k * n_mel + j], acc);
        out[f * n_mel + j] = (float)log(acc + log_offset);
    }
}

// examples[e, t, :] = logmel[e * hop + t, :]   (mel_features.frame applied to the feature rows)
__global__ void __launch_bounds__(LM_THREADS)
frame_examples_kernel(const float *__restrict__ logmel, int n_mel, int win, int hop, float *__restrict__ out) {
    const int64_t e = blockIdx.x;
    const float *src = logmel + e * hop * n_mel;
    float *dst = out + e * (int64_t)win * n_mel;
    for (int i = threadIdx.x; i < win * n_mel; i += LM_THREADS) dst[i] = src[i];
}

}  // namespace

extern "C" int avtex_logmel(const double *wave, int64_t n_samples, int channels, int win_len, int hop, int fft_len,
                            const double *window, const double *mel, int n_mel, double log_offset, float *out,
                            int64_t n_frames, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(n_samples >= win_len && win_len >= 1 && hop >= 1 && channels >= 1 && n_mel >= 1,
                  "logmel: bad arguments n_samples=%lld win=%d hop=%d", (long long)n_samples, win_len, hop);
    AVTEX_REQUIRE(fft_len >= win_len && (fft_len & (fft_len - 1)) == 0 && fft_len <= 4096,
                  "logmel: fft_len=%d must be a power of two >= the window and <= 4096", fft_len);
    AVTEX_REQUIRE(n_frames >= 1 && (n_frames - 1) * hop + win_len <= n_samples && n_frames < (int64_t(1) << 31),
                  "logmel: %lld frames do not fit in %lld samples", (long long)n_frames, (long long)n_samples);
    const size_t smem = (size_t)(3 * fft_len + fft_len / 2 + 1) * sizeof(double);
    AVTEX_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    logmel_kernel<<<(unsigned)n_frames, LM_THREADS, smem, as_stream(stream)>>>(wave, n_samples, channels, win_len, hop,
                                                                              fft_len, window, mel, n_mel, log_offset, out);
    AVTEX_LAUNCH_CHECK();
    return 0;
}

extern "C" int avtex_frame_examples(const float *logmel, int64_t n_frames, int n_mel, int win, int hop, float *out,
                                    int64_t n_examples, int device, void *stream) {
    AVTEX_ENTER(device);
    AVTEX_REQUIRE(win >= 1 && hop >= 1 && n_mel >= 1 && n_examples >= 1 && (n_examples - 1) * hop + win <= n_frames &&
                      n_examples < (int64_t(1) << 31),
                  "frame_examples: %lld examples of %d frames do not fit in %lld rows", (long long)n_examples, win,
                  (long long)n_frames);
    frame_examples_kernel<<<(unsigned)n_examples, LM_THREADS, 0, as_stream(stream)>>>(logmel, n_mel, win, hop, out);
    AVTEX_LAUNCH_CHECK();
    return 0;
}
