/* gr4b200.h -- C ABI of the B200-native streaming-DSP hot path for GNU Radio 4 blocks.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. Each entry point names the
 * reference interface (fair-acc/gnuradio4 @ 7a8e2e5, paths relative to /root/reference) whose work it takes over.
 * The reference's device seam is core/include/gnuradio-4.0/Block.hpp:1855-1862 (DeviceEligible && compute_domain is a
 * device => today: warn once, run on CPU); a block's `processBulk_cuda` body calls exactly one function below.
 *
 * Conventions
 *  - Return value: 0 (= gr::work::Status::OK) or a negative gr::work::Status-compatible code
 *    (core/include/gnuradio-4.0/WorkStatus.hpp:12-18). Nothing throws across this boundary; gr4b200_last_error()
 *    returns a thread-local description of the last failure.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Every compute call is asynchronous on
 *    that stream; the caller owns all buffers and keeps them alive until the stream has passed the call.
 *  - Sample pointers are DEVICE pointers unless a name ends in `_host`. complex<float> is interleaved {re, im}
 *    (std::complex<float> layout), 8-byte aligned; 16-byte alignment enables the widest access path.
 *  - One host thread drives one device at a time (gr4b200_init(device) binds the calling thread).
 *  - Plans are opaque handles holding device-resident constants (taps, twiddles, window) and per-stream carry-over
 *    state (FIR history, mixer phase). A plan is used by one stream at a time.
 */
#ifndef GR4B200_H
#define GR4B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GR4B200_ABI_VERSION 1

/* gr::work::Status values (WorkStatus.hpp:12-18) */
#define GR4B200_OK 0
#define GR4B200_DONE (-1)
#define GR4B200_INSUFFICIENT_INPUT_ITEMS (-2)
#define GR4B200_INSUFFICIENT_OUTPUT_ITEMS (-3)
#define GR4B200_ERROR (-100)

/* MathOpImpl / MathOpMultiPortImpl operator selector (blocks/math/include/gnuradio-4.0/math/Math.hpp:25-28,68-71) */
#define GR4B200_OP_ADD 0
#define GR4B200_OP_SUBTRACT 1
#define GR4B200_OP_MULTIPLY 2
#define GR4B200_OP_DIVIDE 3

/* FIR arithmetic mode */
#define GR4B200_FIR_EXACT 1 /* reference summation order, separately rounded mul/add: bit-identical to the CPU path */
#define GR4B200_FIR_FAST 0  /* same order of taps, fused multiply-add: |err| <= gamma_ntaps * sum|b_k x_{n-k}| */
#define GR4B200_FIR_OVERLAP_SAVE 2 /* y = IFFT(FFT(x) . FFT(b)) on blocks of 4096 (complex<float>, full rate, nTaps <= 2049): bound by
                                    * HBM instead of the fp32 pipe; float-transform accuracy, a few 1e-7 of sum|b| * max|x|, NOT the
                                    * reference's rounding -- an opt-in tolerance mode next to the two direct forms */

/* FFT block output flags (blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:109-111) */
#define GR4B200_FFT_OUTPUT_IN_DB 1u
#define GR4B200_FFT_OUTPUT_IN_DEG 2u
#define GR4B200_FFT_UNWRAP_PHASE 4u

typedef struct gr4b200_fir_plan     gr4b200_fir_plan;
typedef struct gr4b200_fft_plan     gr4b200_fft_plan;
typedef struct gr4b200_rotator_plan gr4b200_rotator_plan;
typedef struct gr4b200_pfb_plan     gr4b200_pfb_plan;
typedef struct gr4b200_ring         gr4b200_ring;

/* ---- runtime ---------------------------------------------------------------------------------------------------- */
int         gr4b200_abi_version(void);
const char* gr4b200_last_error(void);
int         gr4b200_device_count(void);
/* binds the calling thread to `device`; replaces nothing in the reference (there is no device runtime), it is what the
 * "cuda" provider registered with ComputeRegistry (core/include/gnuradio-4.0/ComputeDomain.hpp:123-173) calls first */
int gr4b200_init(int device);
int gr4b200_device_sm_count(int device);

/* device / pinned-host memory: what the "cuda" ProviderFn memory_resource (ComputeDomain.hpp:105) allocates from */
void* gr4b200_malloc(size_t bytes);
int   gr4b200_free(void* devicePtr);
void* gr4b200_malloc_host(size_t bytes); /* pinned */
int   gr4b200_free_host(void* hostPtr);
int   gr4b200_memset(void* devicePtr, int value, size_t bytes, void* stream);
int   gr4b200_copy_h2d(void* devicePtr, const void* hostPtr, size_t bytes, void* stream);
int   gr4b200_copy_d2h(void* hostPtr, const void* devicePtr, size_t bytes, void* stream);
int   gr4b200_copy_d2d(void* dst, const void* src, size_t bytes, void* stream);
/* `height` rows of `widthBytes`, row r from src + r * srcPitch to dst + r * dstPitch: one plane out of every FFT frame
 * (the DataSet is built lazily on the host: only the signals a consumer asks for cross the link) */
int   gr4b200_copy_d2h_2d(void* hostPtr, size_t dstPitch, const void* devicePtr, size_t srcPitch, size_t widthBytes, size_t height, void* stream);
/* kernels launched through this library by the calling process so far (every launch is counted where its error is
 * checked): what bench.py reports as gpu_launches */
unsigned long long gr4b200_launch_count(void);

void* gr4b200_stream_create(void);
int   gr4b200_stream_destroy(void* stream);
int   gr4b200_stream_synchronize(void* stream);
void* gr4b200_event_create(void);
int   gr4b200_event_destroy(void* event);
int   gr4b200_event_record(void* event, void* stream);
int   gr4b200_stream_wait_event(void* stream, void* event);
int   gr4b200_event_synchronize(void* event);
int   gr4b200_event_query(void* event); /* 1: everything recorded before it has finished, 0: not yet, < 0: error */
int   gr4b200_event_elapsed_ms(void* start, void* stop, float* ms);

/* ---- HBM edge ring (replaces CircularBuffer<T> for device edges: core/include/gnuradio-4.0/CircularBuffer.hpp:531-577
 * reserve/publish, :839-865 get, :730-759 consume). Single writer, single reader, cursors on the host, storage in HBM.
 * Spans never wrap: reserve/get hand out contiguous ranges only (gr4b200_ring_writable / _available report the
 * contiguous amount), so no mirror half (CircularBuffer.hpp:382-409) and no extra HBM traffic is needed; `history` bytes
 * in front of every read span stay valid (the FIR reads its past samples there instead of carrying a state buffer) --
 * the reader copies the ring tail in front of the base when it leaves the end of the ring; such a ring has one reader.
 * Stream order replaces host-thread order: publish/consume record events, get/reserve make the other stream wait. ---- */
gr4b200_ring* gr4b200_ring_create(int device, size_t capacityBytes, size_t historyBytes);
int           gr4b200_ring_destroy(gr4b200_ring* ring);
size_t        gr4b200_ring_capacity(const gr4b200_ring* ring);
size_t        gr4b200_ring_available(const gr4b200_ring* ring); /* bytes published and not yet consumed */
size_t        gr4b200_ring_writable(const gr4b200_ring* ring);
void*         gr4b200_ring_reserve(gr4b200_ring* ring, size_t bytes, void* stream); /* NULL if not enough contiguous free space; producer stream waits for the last consume */
int           gr4b200_ring_publish(gr4b200_ring* ring, size_t bytes, void* stream); /* records the producer event */
const void*   gr4b200_ring_get(gr4b200_ring* ring, size_t bytes, void* stream);     /* consumer stream waits on it  */
int           gr4b200_ring_consume(gr4b200_ring* ring, size_t bytes, void* stream);
/* more than one consumer on the same edge (CircularBuffer is SPMC: one Writer, N Readers, CircularBuffer.hpp:476-477,
 * 839-865): every reader has its own cursor; space is free once the slowest reader has consumed it. Reader 0 exists from
 * creation (the functions above address it); further readers join before the first publish, as the reference wires all
 * readers at connect time. At most 8 readers per edge. */
int           gr4b200_ring_add_reader(gr4b200_ring* ring); /* index of the new reader (>= 1) or a negative status */
size_t        gr4b200_ring_available_for(const gr4b200_ring* ring, int reader);
const void*   gr4b200_ring_get_for(gr4b200_ring* ring, int reader, size_t bytes, void* stream);
int           gr4b200_ring_consume_for(gr4b200_ring* ring, int reader, size_t bytes, void* stream);
/* a span that runs across the end of the ring, for readers that do not consume whole chunks (Stride<> with overlap,
 * annotated.hpp:121-162): copies `bytes` published bytes starting at the reader's cursor into `dst` (device memory) on
 * `stream`, in stream order behind the publish that covers them. gr4b200_ring_pending_for counts them without the
 * "contiguous" limit of gr4b200_ring_available_for. */
size_t        gr4b200_ring_pending_for(const gr4b200_ring* ring, int reader);
int           gr4b200_ring_read_for(gr4b200_ring* ring, int reader, size_t bytes, void* dst, void* stream);

/* ---- elementwise math ------------------------------------------------------------------------------------------- */
/* MathOpImpl<std::complex<float>, op>::processOne (Math.hpp:38-56): out[i] = in[i] op value. Bit-identical to the
 * reference's std::complex<float> operators (libgcc __mulsc3 / __divsc3 semantics incl. inf/nan recovery). */
int gr4b200_mathop_const_cf32(void* stream, int op, const float* in, float* out, size_t n, float valueRe, float valueIm);
/* MathOpMultiPortImpl::processBulk (Math.hpp:100-107): out = ins[0] op ins[1] op ... (left fold), nInputs in [1,32];
 * `ins_host` is a HOST array of nInputs DEVICE pointers */
int gr4b200_mathop_multi_cf32(void* stream, int op, const float* const* ins_host, size_t nInputs, float* out, size_t n);
/* Decimator<std::complex<float>>::processBulk (blocks/filter/.../time_domain_filter.hpp:234-244): out[j] = in[j*decim] */
int gr4b200_decimate_cf32(void* stream, const float* in, float* out, size_t nIn, size_t decim);

/* ---- sample-format converters -------------------------------------------------------------------------------- */
/* gr::blocks::type::converter::InterleavedToComplex<R, std::complex<float>>::processBulk and
 * ComplexToInterleaved<std::complex<float>, R>::processBulk (blocks/basic/include/gnuradio-4.0/basic/ConverterBlocks.hpp:
 * 233-277) for R = float, int16_t, int8_t: out[i] = {R -> float of in[2i], in[2i+1]} and the reverse with static_cast
 * (truncation toward zero; out-of-range values as the x86-64 build of the reference produces them). `n_complex` counts
 * complex samples; the interleaved buffer holds 2 * n_complex items of the given type. */
#define GR4B200_ITEM_F32 0
#define GR4B200_ITEM_I16 1
#define GR4B200_ITEM_I8 2
int gr4b200_interleaved_to_complex_cf32(void* stream, int item_type, const void* interleaved, float* out, size_t n_complex);
int gr4b200_complex_to_interleaved_cf32(void* stream, int item_type, const float* in, void* interleaved, size_t n_complex);

/* ---- complex mixer ---------------------------------------------------------------------------------------------- */
/* Rotator<std::complex<float>> (blocks/math/include/gnuradio-4.0/math/Rotator.hpp:40-61). The plan carries
 * phase_increment and the accumulated float phase; the device kernel reproduces the reference's sequential float phase
 * recurrence bit-exactly (see DESIGN.md "mixer") and multiplies with the std::complex product rounding. */
gr4b200_rotator_plan* gr4b200_rotator_plan_create(float phaseIncrement, float initialPhase);
int                   gr4b200_rotator_plan_destroy(gr4b200_rotator_plan* plan);
int                   gr4b200_rotator_set_phase(gr4b200_rotator_plan* plan, float accumulatedPhase);
float                 gr4b200_rotator_get_phase(const gr4b200_rotator_plan* plan);
/* Rotator.hpp:41-42: phase_increment = 2 * (pi_f * frequency_shift / sample_rate) in float */
float gr4b200_rotator_phase_increment(float frequencyShift, float sampleRate);
int   gr4b200_rotator_cf32(gr4b200_rotator_plan* plan, void* stream, const float* in, float* out, size_t n);

/* ---- FIR -------------------------------------------------------------------------------------------------------- */
/* fir_filter<T>::processOne (time_domain_filter.hpp:44-47) applied to re and im, and
 * BasicFilterProto<T, Resampling<1,1,false>>::processBulk (time_domain_filter.hpp:190-204) when decimate > 1:
 *   y[n] = sum_k b[k] x[n-k], keep n % decimate == 0 (n counted from the start of each call; nIn % decimate == 0).
 * The plan owns the taps (device) and the (nTaps-1)-sample history carried between calls (zero at creation/reset). */
gr4b200_fir_plan* gr4b200_fir_plan_create(const float* taps_host, size_t nTaps, size_t decimate, int mode);
int               gr4b200_fir_plan_destroy(gr4b200_fir_plan* plan);
int               gr4b200_fir_plan_reset(gr4b200_fir_plan* plan, void* stream);
/* New coefficients for a running filter, with the reference's treatment of the past samples
 * (fir_filter::settingsChanged, time_domain_filter.hpp:39-43): the history buffer holds 32 samples, or bit_ceil(b.size())
 * once a longer filter has been set; a new `b` that still fits KEEPS the past samples (the next nTaps-1 outputs are formed
 * from the old stream), one that does not fit replaces the buffer (zeros). Exact / fast modes; the overlap-save mode returns
 * GR4B200_ERROR (create a new plan). After a fused DDC call only the current filter's nTaps-1 past samples are kept. */
int               gr4b200_fir_plan_set_taps(gr4b200_fir_plan* plan, void* stream, const float* taps_host, size_t nTaps);
int               gr4b200_fir_cf32(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn);
/* real-valued stream, as the reference registers it (time_domain_filter.hpp:20: float) */
int gr4b200_fir_f32(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn);
/* The same filters reading their history from the stream itself: in[-h .. -1], h = gr4b200_fir_plan_history_items(plan),
 * must hold the samples that preceded in[0] (zeros before the stream started) -- true for every span of an HBM ring
 * created with historyBytes >= h items. The plan's carried state is neither read nor updated and no state kernel is
 * launched: consecutive work chunks are independent launches (the reference keeps the past samples in the block's
 * HistoryBuffer, time_domain_filter.hpp:36; here the edge buffer already holds them). */
size_t gr4b200_fir_plan_history_items(const gr4b200_fir_plan* plan);
int    gr4b200_fir_cf32_contiguous(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn);
int    gr4b200_fir_f32_contiguous(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn);

/* FIR design stays on the host (FilterTool.hpp:964-976 generateCoefficients + :415-423 DC normalisation);
 * window type numbering = gr::algorithm::window::Type (fourier/window.hpp:35) */
int  gr4b200_window_f32_host(int windowType, size_t n, float beta, float* out_host);
int  gr4b200_fir_generate_f32_host(size_t nTaps, int windowType, float fc, float beta, int normaliseDc, float* out_host);
long gr4b200_fir_design_f32_host(int filterType, size_t order, double fLow, double fHigh, double fs, double gain, double attenuationDb, double beta, int windowType, float* out_host, size_t capacity);

/* ---- FFT -------------------------------------------------------------------------------------------------------- */
/* gr::algorithm::FFT<std::complex<float>>::compute (algorithm/include/gnuradio-4.0/algorithm/fourier/fft.hpp:113-153):
 * unnormalised forward DFT, natural order; `batch` back-to-back transforms of nfft samples. nfft: power of two in
 * [16, 262144] (one kernel up to 8192 points, two passes of column transforms through a plan-owned scratch buffer
 * above). `window_host` (nullable) = nfft floats multiplied onto re and im before the transform. */
gr4b200_fft_plan* gr4b200_fft_plan_create(size_t nfft, const float* window_host);
int               gr4b200_fft_plan_destroy(gr4b200_fft_plan* plan);
size_t            gr4b200_fft_plan_size(const gr4b200_fft_plan* plan);
int               gr4b200_fft_c2c_cf32(gr4b200_fft_plan* plan, void* stream, const float* in, float* out, size_t batch);
/* gr::algorithm::FFT<float>::compute on real input (fft.hpp:214-258 trySimdFFT_R2C): `batch` transforms of nfft REAL
 * samples each; writes the FULL spectrum, nfft complex bins per transform (bins above nfft/2 are the conjugate mirror,
 * the DC and Nyquist bins are real), the layout the reference unpacks its packed result into. Window as in the plan. */
int               gr4b200_fft_r2c_f32(gr4b200_fft_plan* plan, void* stream, const float* in, float* out, size_t batch);
/* FFT block processBulk + createDataset numerics (blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:147-250): per
 * chunk c of nfft input samples writes signals[c][4][nfft] = {magnitude*2/N fft-shifted, phase fft-shifted, Re, Im}
 * and (if ranges != NULL) ranges[c][4][2] = {min, max} of each signal. flags: GR4B200_FFT_*. */
int gr4b200_fft_block_cf32(gr4b200_fft_plan* plan, void* stream, const float* in, size_t batch, unsigned flags, float* signals, float* ranges);

/* The same block on a REAL stream, FFT<float> (fft.hpp:147-250 with computeFullSpectrum == false): per chunk c of nfft real
 * samples writes signals[c][4][nfft/2] = {magnitude*2/N of bins [0, N/2) (no fft-shift for a half spectrum), phase of the
 * same bins, Re and Im of bins [N/2, N) -- createDataset copies the LAST N/2 bins of the spectrum (fft.hpp:212-217)} and
 * ranges[c][4][2] if not NULL. */
int gr4b200_fft_block_f32(gr4b200_fft_plan* plan, void* stream, const float* in, size_t batch, unsigned flags, float* signals, float* ranges);

/* ---- fused DDC: Rotator -> decimating FIR (compile-time Merge, BlockMerging.hpp:125-138, as device fusion)
 * out[j] = FIR_decim(rotator(in))[j]; same numerics as the two calls back to back. ------------------------------------ */
int gr4b200_ddc_cf32(gr4b200_rotator_plan* mixer, gr4b200_fir_plan* fir, void* stream, const float* in, float* out, size_t nIn);

/* ---- fused FIR -> FFT block: Merge<fir_filter, FFT> (BlockMerging.hpp:125-138) as one kernel. signals[c][4][4096] exactly
 * as gr4b200_fir_cf32 followed by gr4b200_fft_block_cf32 (ranges = NULL) would produce them, bit for bit; the filtered
 * stream is not materialised. Needs a full-rate complex FIR plan, an FFT plan of size 4096, nIn % 4096 == 0 and no
 * phase unwrapping (gr4b200_fir_fft_fused_supported); the FIR history carries over between calls as usual. ------------ */
int gr4b200_fir_fft_fused_supported(const gr4b200_fir_plan* fir, const gr4b200_fft_plan* fft, unsigned flags);
int gr4b200_fir_fft_block_cf32(gr4b200_fir_plan* fir, gr4b200_fft_plan* fft, void* stream, const float* in, size_t nIn, unsigned flags, float* signals);

/* ---- polyphase channelizer (no reference implementation exists: own definition) ------------------- */
gr4b200_pfb_plan* gr4b200_pfb_plan_create(const float* proto_host, size_t nChannels, size_t tapsPerBranch);
int               gr4b200_pfb_plan_destroy(gr4b200_pfb_plan* plan);
int               gr4b200_pfb_plan_reset(gr4b200_pfb_plan* plan, void* stream);
/* stage 1: polyphase FIR bank; in: nFrames*nChannels samples; out: u[frame][branch] */
int gr4b200_pfb_filter_cf32(gr4b200_pfb_plan* plan, void* stream, const float* in, float* out, size_t nFrames);
/* stage 2: nChannels-point FFT over each frame = gr4b200_fft_c2c_cf32 with a plan of size nChannels */
/* both stages in one kernel (the filter-bank outputs never reach HBM): y[frame][channel]; available for 256 channels
 * with 4, 8 or 12 taps per branch (gr4b200_pfb_fused_supported), same history carry-over as the filter stage */
int gr4b200_pfb_fused_supported(const gr4b200_pfb_plan* plan);
int gr4b200_pfb_channelizer_cf32(gr4b200_pfb_plan* plan, void* stream, const float* in, float* out, size_t nFrames);

/* ---- polyphase rational resampler (no reference implementation either: own definition):
 *   y[m] = sum_{k<P} h[(m M) mod L + k L] * x[floor(m M / L) - k],  P = ceil(nTaps / L),  acc = fma(h, x, acc), k ascending.
 * nIn % decimation == 0; writes nIn / decimation * interpolation samples; the (P-1)-sample history carries over. --------- */
typedef struct gr4b200_resampler_plan gr4b200_resampler_plan;
gr4b200_resampler_plan* gr4b200_resampler_plan_create(const float* taps_host, size_t nTaps, size_t interpolation, size_t decimation);
int                     gr4b200_resampler_plan_destroy(gr4b200_resampler_plan* plan);
int                     gr4b200_resampler_plan_reset(gr4b200_resampler_plan* plan, void* stream);
int gr4b200_resampler_cf32(gr4b200_resampler_plan* plan, void* stream, const float* in, float* out, size_t nIn);

/* ---- inter-GPU edges (pipelined mode; the reference's analogue is the multiThreaded job list hand-off through a
 * CircularBuffer, core/include/gnuradio-4.0/Scheduler.hpp:1944-1951) ------------------------------------------------- */
int gr4b200_peer_enable(int device, int peerDevice);
int gr4b200_peer_copy(void* dst, int dstDevice, const void* src, int srcDevice, size_t bytes, void* stream);
/* An edge between two PROCESSES (one per GPU): the consumer exports its edge buffer, the producer maps it and its last
 * kernel stores straight into it over NVLink -- no staging copy, no collective. `handle` is 64 bytes (cudaIpcMemHandle_t),
 * to be carried to the other process by whatever channel the host application has (bench.py: torch.distributed).
 * The reference's analogue: the CircularBuffer between two job lists of the multi-threaded scheduler (Scheduler.hpp:1944-1951). */
int   gr4b200_ipc_export(void* devicePtr, void* handle64);
void* gr4b200_ipc_open(const void* handle64);   /* device pointer valid in the calling process, NULL on failure */
int   gr4b200_ipc_close(void* mappedPtr);
/* Cursors of such an edge live in device memory as 32-bit counters and are moved / awaited by the STREAMS, not by the
 * hosts: `write` stores `value` behind everything enqueued on `stream` so far (a one-thread kernel: plain store + system
 * fence, so the target may be a peer's memory); `wait` holds `stream` until *devicePtr >= value (cuStreamWaitValue32 on the
 * local counter). Neither blocks the calling thread. */
int gr4b200_stream_write_value32(void* stream, void* devicePtr, unsigned value);
int gr4b200_stream_wait_value32(void* stream, void* devicePtr, unsigned value);

#ifdef __cplusplus
}
#endif
#endif /* GR4B200_H */
