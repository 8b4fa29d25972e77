// mico_b200 -- K11: Kaldi-compatible log-mel filterbank front-end (model/audioprocessor.py:38-46 ->
// torchaudio.compliance.kaldi.fbank with the defaults the reference uses: 25 ms / 10 ms frames at 16 kHz, snip_edges,
// remove_dc_offset, pre-emphasis 0.97, povey window, zero-pad to 512, power spectrum, mel banks 20 Hz .. Nyquist,
// log(max(x, eps))) fused with the reference's normalisation (x - 15.41663) / (2 * 6.55582).
//
// One 128-thread block per frame: the 400-sample frame is read with coalesced loads straight from the waveform
// (frames overlap by 60 %, so the re-reads hit L2), DC removal + pre-emphasis + window in registers/shared memory,
// a 512-point radix-2 FFT in shared memory (9 stages, 256 butterflies each = 2 per thread), |X|^2 for the 257 bins,
// then each thread produces mel bins t, t+128, ... as a 257-term dot product and writes log-mel rows coalesced.
// HBM-bound by design: 640 B read and 4*num_mel B written per frame.
#include "common.cuh"
#include "host_utils.h"

namespace mico {
namespace {

constexpr int kFftN = 512, kFftLog = 9, kFbThreads = 128;

__global__ void __launch_bounds__(kFbThreads)
fbank_kernel(const float* __restrict__ wave, int64_t clip_stride, int n_frames, int frame_len, int frame_shift,
             const float* __restrict__ window, const float* __restrict__ mel, int num_mel, float in_scale, float preemph,
             float log_floor, float norm_sub, float norm_mul, float* __restrict__ out, int64_t out_clip_stride) {
    __shared__ float re[kFftN], im[kFftN], red[kFbThreads / 32];
    const int clip = blockIdx.y, f = blockIdx.x, t = threadIdx.x;
    const float* src = wave + (int64_t)clip * clip_stride + (int64_t)f * frame_shift;
    // ---- load (x 2^15, audioprocessor.py:39) and DC offset
    float s = 0.f;
    for (int i = t; i < kFftN; i += kFbThreads) {
        const float v = i < frame_len ? src[i] * in_scale : 0.f;
        re[i] = v;
        s += v;
    }
    s = warp_sum(s);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    const float mean = (red[0] + red[1] + red[2] + red[3]) / (float)frame_len;
    // ---- pre-emphasis (x[i] - c * x[i-1], x[-1] := x[0]) on the DC-removed signal, window, bit-reversed scatter
    float w[kFftN / kFbThreads];
#pragma unroll
    for (int j = 0; j < kFftN / kFbThreads; ++j) {
        const int i = t + j * kFbThreads;
        float v = 0.f;
        if (i < frame_len) {
            const float cur = re[i] - mean, prev = re[i > 0 ? i - 1 : 0] - mean;
            v = (cur - preemph * prev) * window[i];
        }
        w[j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kFftN / kFbThreads; ++j) {
        const int i = t + j * kFbThreads;
        const int r = (int)(__brev((unsigned)i) >> (32 - kFftLog));
        re[r] = w[j];
        im[r] = 0.f;
    }
    __syncthreads();
    // ---- radix-2 decimation-in-time FFT
    for (int st = 1; st <= kFftLog; ++st) {
        const int half = 1 << (st - 1);
#pragma unroll
        for (int j = 0; j < (kFftN / 2) / kFbThreads; ++j) {
            const int b = t + j * kFbThreads;              // butterfly index 0..255
            const int pos = b & (half - 1);
            const int i0 = ((b >> (st - 1)) << st) + pos, i1 = i0 + half;
            float sn, cs;
            sincospif(-(float)pos / (float)half, &sn, &cs);
            const float xr = re[i1] * cs - im[i1] * sn, xi = re[i1] * sn + im[i1] * cs;
            const float ar = re[i0], ai = im[i0];
            re[i0] = ar + xr; im[i0] = ai + xi;
            re[i1] = ar - xr; im[i1] = ai - xi;
        }
        __syncthreads();
    }
    // ---- power spectrum into re[0..256]
    float pw[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int i = t + j * kFbThreads;
        pw[j] = i <= kFftN / 2 ? re[i] * re[i] + im[i] * im[i] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int i = t + j * kFbThreads;
        if (i <= kFftN / 2) re[i] = pw[j];
    }
    __syncthreads();
    // ---- mel energies, log, normalise
    float* orow = out + (int64_t)clip * out_clip_stride + (int64_t)f * num_mel;
    for (int m = t; m < num_mel; m += kFbThreads) {
        const float* mw = mel + (int64_t)m * (kFftN / 2 + 1);
        float acc = 0.f;
        for (int k = 0; k <= kFftN / 2; ++k) acc += re[k] * __ldg(mw + k);
        orow[m] = (__logf(fmaxf(acc, log_floor)) - norm_sub) * norm_mul;
    }
}

}  // namespace
}  // namespace mico

extern "C" int mico_fbank(const float* wave, int64_t clip_stride, int n_clips, int n_samples, int frame_len, int frame_shift,
                          const float* window, const float* mel, int num_mel, float in_scale, float preemph, float log_floor,
                          float norm_sub, float norm_mul, float* out, int64_t out_clip_stride, void* stream_) {
    using namespace mico;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MICO_CHECK_ARG(wave && window && mel && out && n_clips > 0 && num_mel > 0);
    MICO_CHECK_ARG(frame_len > 0 && frame_len <= kFftN && frame_shift > 0 && n_samples >= frame_len);
    const int n_frames = 1 + (n_samples - frame_len) / frame_shift;      // snip_edges = True
    MICO_CHECK_ARG(out_clip_stride >= (int64_t)n_frames * num_mel);
    ProfScope prof(kProfOther, (double)n_clips * ((double)n_samples * 4 + (double)n_frames * num_mel * 4), stream);
    dim3 grid(n_frames, n_clips);
    fbank_kernel<<<grid, kFbThreads, 0, stream>>>(wave, clip_stride, n_frames, frame_len, frame_shift, window, mel, num_mel,
                                                  in_scale, preemph, log_floor, norm_sub, norm_mul, out, out_clip_stride);
    MICO_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return MICO_OK;
}
