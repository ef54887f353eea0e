// Row N2: gather + taper + overlap-add of the selected units' MagPhase frames.
//
// Replaces the array part of Synthesiser.concatenateMagPhaseEpoch_sep_files and
// retrieve_magphase_frag (reference script/synth_simple.py:677-747, 538-652) with
// zero_pad_matrix / taper_matrix (script/matrix_operations.py:1-30): for every selected unit the
// fragment of multiepoch + overlap frames around it (zero padded at sentence edges) is weighted by
// a Hann cross-fade and added into the output at hop multiepoch; the leading and trailing
// overlap/2 frames are trimmed; f0 is zeroed where the cross-faded voicing flag is below 0.5.
// The waveform synthesis that follows (magphase.synthesis_from_lossless, external) stays out of
// scope.  Arithmetic follows the reference: float32 frames times float64 taper, float64 sums in
// path order.
#include "common.cuh"
#include <vector>

struct snk_frames {
    int device = 0;
    int64_t nframes = 0, nunits = 0;
    int width = 0;
    float *mag = nullptr, *real = nullptr, *imag = nullptr;   // [nframes, width]
    double *f0 = nullptr, *vuv = nullptr;                     // [nframes] (lin_interp_f0 returns float64)
    int64_t *unit_frame = nullptr, *sent_lo = nullptr, *sent_hi = nullptr;   // [nunits]
    cudaStream_t stream = nullptr;
    snk_buf w_path, w_taper, w_out, w_small;
};

namespace {

// one CTA per untrimmed output frame r; threads stride the spectrum bins
__global__ void concat_kernel(const float *__restrict__ mag, const float *__restrict__ real, const float *__restrict__ imag,
                              const double *__restrict__ f0, const double *__restrict__ vuv, int width,
                              const int64_t *__restrict__ unit_frame, const int64_t *__restrict__ sent_lo,
                              const int64_t *__restrict__ sent_hi, const int64_t *__restrict__ path, int64_t P, int m,
                              int overlap, const double *__restrict__ taper_in /*[overlap]*/, int has_fzero,
                              double *__restrict__ omag, double *__restrict__ oreal, double *__restrict__ oimag,
                              double *__restrict__ ofz, double *__restrict__ ovuv) {
    const int extra = overlap / 2;
    const int64_t r = (int64_t)blockIdx.x + extra;          // untrimmed row; rows < extra and the last extra rows are trimmed
    const int flen = m + overlap;
    // path positions whose fragment covers row r: p*m <= r < p*m + flen
    int64_t p_lo = (r - flen + 1 + m - 1) / m;
    if (r - flen + 1 < 0) p_lo = 0;
    int64_t p_hi = r / m;
    if (p_hi > P - 1) p_hi = P - 1;
    const int64_t orow = r - extra;
    double fz = 0.0, vv = 0.0;
    for (int c = threadIdx.x; c < width; c += blockDim.x) {
        double am = 0.0, ar = 0.0, ai = 0.0;
        for (int64_t p = p_lo; p <= p_hi; ++p) {             // ascending: the reference's += order
            const int l = (int)(r - p * m);
            const int64_t u = path[p];
            const int64_t f_beg = unit_frame[u] - extra;
            const int64_t src = f_beg + l;
            if (src < sent_lo[u] || src >= sent_hi[u]) continue;          // zero padding (zero_pad_matrix)
            double w = 1.0;
            bool tapered = false;
            if (overlap > 0) {
                if (l < overlap) { w = taper_in[l]; tapered = true; }
                else if (l >= flen - overlap) { w = taper_in[flen - 1 - l]; tapered = true; }  // out taper = flipped in taper
            }
            double vm = (double)mag[src * width + c], vr = (double)real[src * width + c], vi = (double)imag[src * width + c];
            if (tapered) {
                // taper_matrix multiplies IN PLACE: an unpadded fragment is still the float32 slice of the file, so
                // the product is rounded back to float32; a zero-padded fragment was promoted to float64 by vstack
                const bool padded = f_beg < sent_lo[u] || f_beg + flen > sent_hi[u];
                vm = __dmul_rn(vm, w); vr = __dmul_rn(vr, w); vi = __dmul_rn(vi, w);   // no FMA contraction: numpy rounds the product
                if (!padded) { vm = (double)(float)vm; vr = (double)(float)vr; vi = (double)(float)vi; }
            }
            am = __dadd_rn(am, vm); ar = __dadd_rn(ar, vr); ai = __dadd_rn(ai, vi);
        }
        omag[orow * width + c] = am;
        oreal[orow * width + c] = ar;
        oimag[orow * width + c] = ai;
    }
    if (threadIdx.x == 0) {
        for (int64_t p = p_lo; p <= p_hi; ++p) {
            const int l = (int)(r - p * m);
            const int64_t u = path[p];
            const int64_t src = unit_frame[u] - extra + l;
            if (src < sent_lo[u] || src >= sent_hi[u]) continue;
            double w = 1.0;
            if (overlap > 0) {
                if (l < overlap) w = taper_in[l];
                else if (l >= flen - overlap) w = taper_in[flen - 1 - l];
            }
            fz = __dadd_rn(fz, __dmul_rn(f0[src], w));
            vv = __dadd_rn(vv, __dmul_rn(vuv[src], w));
        }
        if (!has_fzero && vv < 0.5) fz = 0.0;                 // synth_simple.py:727-729
        ofz[orow] = fz;
        ovuv[orow] = vv;
    }
}

}  // namespace

extern "C" {

int snk_frames_create(snk_frames **out, int device_id, int64_t nframes, int width, const float *mag, const float *real,
                      const float *imag, const double *f0_interp, const double *vuv, int64_t nunits,
                      const int64_t *unit_frame, const int64_t *sent_lo, const int64_t *sent_hi) {
    SNK_CHECK(out, "out is NULL");
    *out = nullptr;
    SNK_CHECK(mag && real && imag && f0_interp && vuv && unit_frame && sent_lo && sent_hi, "NULL argument");
    SNK_CHECK(nframes >= 1 && width >= 1 && nunits >= 1, "bad frame store shape");
    SNK_CHECK(snk_device_count() > 0, "no CUDA device visible: this engine has no CPU fallback");
    SNK_CUDA(cudaSetDevice(device_id));
    snk_frames *fr = new snk_frames();
    fr->device = device_id; fr->nframes = nframes; fr->nunits = nunits; fr->width = width;
    const size_t fb = (size_t)nframes * width * 4;
    bool ok = cudaMalloc((void **)&fr->mag, fb) == cudaSuccess && cudaMalloc((void **)&fr->real, fb) == cudaSuccess &&
              cudaMalloc((void **)&fr->imag, fb) == cudaSuccess && cudaMalloc((void **)&fr->f0, (size_t)nframes * 8) == cudaSuccess &&
              cudaMalloc((void **)&fr->vuv, (size_t)nframes * 8) == cudaSuccess &&
              cudaMalloc((void **)&fr->unit_frame, (size_t)nunits * 8) == cudaSuccess &&
              cudaMalloc((void **)&fr->sent_lo, (size_t)nunits * 8) == cudaSuccess &&
              cudaMalloc((void **)&fr->sent_hi, (size_t)nunits * 8) == cudaSuccess &&
              cudaStreamCreateWithFlags(&fr->stream, cudaStreamNonBlocking) == cudaSuccess;
    if (ok)
        ok = cudaMemcpy(fr->mag, mag, fb, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(fr->real, real, fb, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(fr->imag, imag, fb, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(fr->f0, f0_interp, (size_t)nframes * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(fr->vuv, vuv, (size_t)nframes * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(fr->unit_frame, unit_frame, (size_t)nunits * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(fr->sent_lo, sent_lo, (size_t)nunits * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(fr->sent_hi, sent_hi, (size_t)nunits * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        snk_set_error("frame store allocation / upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        snk_frames_destroy(fr);
        return 1;
    }
    *out = fr;
    return 0;
}

int snk_frames_destroy(snk_frames *fr) {
    if (!fr) return 0;
    cudaSetDevice(fr->device);
    if (fr->stream) cudaStreamSynchronize(fr->stream);
    cudaFree(fr->mag); cudaFree(fr->real); cudaFree(fr->imag); cudaFree(fr->f0); cudaFree(fr->vuv);
    cudaFree(fr->unit_frame); cudaFree(fr->sent_lo); cudaFree(fr->sent_hi);
    snk_buf_free(&fr->w_path); snk_buf_free(&fr->w_taper); snk_buf_free(&fr->w_out); snk_buf_free(&fr->w_small);
    if (fr->stream) cudaStreamDestroy(fr->stream);
    cudaGetLastError();
    delete fr;
    return 0;
}

int snk_concat_magphase_epoch(snk_frames *fr, const int64_t *path, int64_t P, int multiepoch, int overlap,
                              const double *taper_in, int has_fzero, double *mag, double *real, double *imag, double *fz,
                              double *vuv, double *kernel_ms) {
    SNK_CHECK(fr && path && mag && real && imag && fz && vuv, "NULL argument");
    SNK_CHECK(P >= 1 && multiepoch >= 1, "bad path length / multiepoch");
    SNK_CHECK(overlap >= 0 && overlap % 2 == 0, "frame overlap should be even number");        // synth_simple.py:678
    SNK_CHECK(overlap <= multiepoch, "taper_length (%d) too long for (padded) unit length (%d)", overlap,
              multiepoch + overlap);                                                           // matrix_operations.py:18
    SNK_CHECK(overlap == 0 || taper_in, "taper is NULL");
    for (int64_t p = 0; p < P; ++p) SNK_CHECK(path[p] >= 0 && path[p] < fr->nunits, "unit id %lld out of range", (long long)path[p]);
    SNK_CUDA(cudaSetDevice(fr->device));
    const int64_t rows = P * multiepoch;      // after trimming overlap/2 frames at both ends
    const int W = fr->width;
    SNK_TRY(snk_buf_reserve(&fr->w_path, (size_t)P * 8));
    SNK_TRY(snk_buf_reserve(&fr->w_taper, (size_t)std::max(overlap, 1) * 8));
    SNK_TRY(snk_buf_reserve(&fr->w_out, (size_t)rows * W * 8 * 3));
    SNK_TRY(snk_buf_reserve(&fr->w_small, (size_t)rows * 8 * 2));
    double *o_mag = (double *)fr->w_out.p, *o_real = o_mag + (size_t)rows * W, *o_imag = o_real + (size_t)rows * W;
    double *o_fz = (double *)fr->w_small.p, *o_vuv = o_fz + rows;
    SNK_CUDA(cudaMemcpyAsync(fr->w_path.p, path, (size_t)P * 8, cudaMemcpyHostToDevice, fr->stream));
    if (overlap) SNK_CUDA(cudaMemcpyAsync(fr->w_taper.p, taper_in, (size_t)overlap * 8, cudaMemcpyHostToDevice, fr->stream));
    cudaEvent_t e0, e1;
    SNK_CUDA(cudaEventCreate(&e0));
    SNK_CUDA(cudaEventCreate(&e1));
    SNK_CUDA(cudaEventRecord(e0, fr->stream));
    concat_kernel<<<(unsigned)rows, 256, 0, fr->stream>>>(fr->mag, fr->real, fr->imag, fr->f0, fr->vuv, W, fr->unit_frame,
                                                          fr->sent_lo, fr->sent_hi, (const int64_t *)fr->w_path.p, P,
                                                          multiepoch, overlap, (const double *)fr->w_taper.p, has_fzero,
                                                          o_mag, o_real, o_imag, o_fz, o_vuv);
    SNK_CUDA(cudaGetLastError());
    SNK_CUDA(cudaEventRecord(e1, fr->stream));
    SNK_CUDA(cudaMemcpyAsync(mag, o_mag, (size_t)rows * W * 8, cudaMemcpyDeviceToHost, fr->stream));
    SNK_CUDA(cudaMemcpyAsync(real, o_real, (size_t)rows * W * 8, cudaMemcpyDeviceToHost, fr->stream));
    SNK_CUDA(cudaMemcpyAsync(imag, o_imag, (size_t)rows * W * 8, cudaMemcpyDeviceToHost, fr->stream));
    SNK_CUDA(cudaMemcpyAsync(fz, o_fz, (size_t)rows * 8, cudaMemcpyDeviceToHost, fr->stream));
    SNK_CUDA(cudaMemcpyAsync(vuv, o_vuv, (size_t)rows * 8, cudaMemcpyDeviceToHost, fr->stream));
    SNK_CUDA(cudaStreamSynchronize(fr->stream));
    if (kernel_ms) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *kernel_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

}  // extern "C"
