/* include/silero_b200.h -- C ABI of the B200-native Silero VAD v3.1 (16 kHz) engine.
 *
 * Drop-in boundary for the C backend of IntendedConsequence/vadc (citations relative to the
 * reference tree). Plain pointers and sizes only; implemented by libsilero_b200.so
 * (vadc_b200/csrc/, hand-written sm_100a CUDA + C host code). There is no CPU fallback: every
 * compute entry point returns SILERO_B200_ERR_CUDA when no usable device exists.
 *
 *   reference interface                                     replaced by
 *   ------------------------------------------------------  -----------------------------------
 *   backend_init            silero.h:48-51 (silero_init :21) silero_b200_create[_from_file]
 *   backend_run             silero.h:53-74                   silero_b200_run_chunks
 *   silero_run_one_batch_with_context   silero_v3.c:72-215   silero_b200_run_chunks
 *   Silero_Context.state_lstm_h/c       tensor.h:89-95       per-stream device state + get/set/reset
 *   run_inference s16->f32 + process_chunks vadc.c:56-103,873-909   silero_b200_run_streams
 *   feed_probability / combine / emit   vadc.c:165-299,1005-1027    vadc_segments_* (vadc_segmenter.h, host) and
 *                                                            silero_b200_run_streams_segments* (same state machine on the device)
 *   load_testtensor(_from_bytes)        tensor.h:201-325     the .testtensor blob passed to create
 *   my_stft, adaptive_audio_normalization_inplace, transformer_layer, lstm_tensor_minibatched,
 *   decoder_tensor (stft.c:226, misc.c:1, transformer.c:237, lstm.c:228, silero_v3.c:305)
 *                                                            silero_b200_stage_* (parity taps)
 *
 * include/vadc_dropin/silero.h implements the reference's three backend_* functions on top of this
 * ABI so that the reference's vadc.c builds against it unchanged.
 *
 * Threading: calls on one handle must not overlap; one handle drives one GPU. Use one handle per GPU
 * (one process per GPU, or several handles in one process).
 */
#ifndef SILERO_B200_H
#define SILERO_B200_H

#include <stddef.h>
#include <stdint.h>

#include "vadc_segmenter.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SILERO_B200_CHUNK_SAMPLES 1536   /* silero.h:41-42: input_size_min == input_size_max == 1536 */
#define SILERO_B200_SAMPLE_RATE 16000
#define SILERO_B200_STATE_FLOATS 128     /* h or c: [2 layers][64] (tensor.h:93-94) */

enum
{
   SILERO_B200_OK = 0,
   SILERO_B200_ERR_ARG = -1,      /* bad argument (NULL, negative count, stream out of range) */
   SILERO_B200_ERR_WEIGHTS = -2,  /* .testtensor blob malformed or not the 99-tensor v3.1 layout */
   SILERO_B200_ERR_CUDA = -3,     /* CUDA runtime error or no device; see silero_b200_last_error */
   SILERO_B200_ERR_NOMEM = -4
};

typedef struct silero_b200 silero_b200; /* opaque engine handle */

/* Kernel families. The family is a property of the ENGINE, decided when it is created (never per call: a persistent stream must not
   change arithmetic when the caller's batch shape changes):
     exact  every mode AUTO (the default), or LAYERS_FAITHFUL / LSTM_FAITHFUL: the reference's own rounding sequence from the STFT to the
            probability (stft_sym_kernel, exact_front/layer kernels, exact_lstm_kernel; few streams: CTA-per-chunk encoder and LSTM
            wavefront -- same bits). Results are bit-identical to the reference build for ANY number of streams and any stream length.
     fast   any explicit STFT_HYBRID / LSTM_FP32 / LSTM_TENSOR / LAYERS_FP32 / LAYERS_TENSOR: FFT-hybrid STFT, FMA chains or tcgen05
            tensor-core contractions with split operands, SFU nonlinearities. ~7x the throughput, within 1e-4 of the reference chunk by
            chunk on short streams; on long streams the decoder LSTM integrates the one-ulp differences. Opt-in. */
/* STFT evaluation. HYBRID: fp32 FFT everywhere + the reference's exact rounding sequence (stft.c:108-184) for every bin whose magnitude is
   below stft_k_rel * ||windowed frame||_2 (stft_fft8_kernel.cuh); EXACT: the reference's sequence for every bin (bit-identical
   magnitudes; stft_sym_kernel.cuh, or stft_kernel.cuh for a basis without the mirror property). */
#define SILERO_B200_STFT_AUTO 0            /* default: _EXACT, except in a fast-family engine whose LSTM runs on the tensor cores (_HYBRID) */
#define SILERO_B200_STFT_HYBRID 4
#define SILERO_B200_STFT_EXACT 1
#define SILERO_B200_STFT_K_REL_DEFAULT 0.004f

/* Decoder LSTM of a fast-family engine. FP32: gate contractions as fp32 FMA chains on the CUDA cores. TENSOR: tcgen05 tensor-core GEMM
   over tiles of 32 streams with the bf16x2 split (3 partial products, fp32 accumulation). In a fast-family engine AUTO resolves, at
   creation, to TENSOR when max_streams >= SILERO_B200_LSTM_TENSOR_MIN_STREAMS, else FP32. FAITHFUL selects the exact path. */
#define SILERO_B200_LSTM_AUTO 0
#define SILERO_B200_LSTM_FP32 1
#define SILERO_B200_LSTM_TENSOR 2
#define SILERO_B200_LSTM_TENSOR_MIN_STREAMS 1024
#define SILERO_B200_LSTM_FAITHFUL 3

/* Encoder layers of a fast-family engine. FP32: every contraction as in-thread fp32 FMA chains on the CUDA cores. TENSOR: the dense
   contractions as tcgen05 tensor-core GEMMs over tiles of 128 tokens with the fp16x2 split (hi/lo, 3 partial products, 22 significant
   bits per operand, fp32 accumulation). AUTO resolves like the LSTM's, from max_streams, at creation. FAITHFUL selects the exact path
   (either FAITHFUL flag selects all of it: exact STFT, encoder, LSTM, decoder). */
#define SILERO_B200_LAYERS_AUTO 0
#define SILERO_B200_LAYERS_FP32 1
#define SILERO_B200_LAYERS_TENSOR 2
#define SILERO_B200_LAYERS_FAITHFUL 3
#define SILERO_B200_EXACT_TOKEN_MIN_CHUNKS 384    /* exact path: windows of at least this many chunks run the thread-per-token encoder
                                                     (below: a CTA per chunk -- more parallelism for small windows; identical bits).
                                                     Measured crossover (scripts/gpu_encoder_crossover.sh): 192 chunks 0.15 vs 0.18 ms,
                                                     384 chunks 0.18 vs 0.18 ms, 768 chunks 0.31 vs 0.21 ms */

typedef struct silero_b200_opts
{
   int device;          /* CUDA device ordinal (default 0) */
   int max_streams;     /* number of independent streams whose LSTM state is kept on device (default 1) */
   int window_chunks;   /* chunks per stream processed per internal pass; 0 = choose from memory budget */
   int stft_mode;       /* SILERO_B200_STFT_AUTO (default), _HYBRID or _EXACT */
   float stft_k_rel;    /* hybrid threshold; 0 = SILERO_B200_STFT_K_REL_DEFAULT */
   int lstm_mode;       /* SILERO_B200_LSTM_AUTO (default), _FP32 (CUDA-core kernel), _TENSOR (tcgen05 kernel) or _FAITHFUL */
   int layer_mode;      /* SILERO_B200_LAYERS_AUTO (default), _FP32 (CUDA-core kernels), _TENSOR (tcgen05 kernel) or _FAITHFUL */
   int reserved[1];
} silero_b200_opts;

void silero_b200_default_opts( silero_b200_opts *opts );

/* What backend_init reports through Silero_Config (silero.h:39-43) */
typedef struct silero_b200_info
{
   int batch_size_restriction; /* -1: any batch */
   int is_silero_v5;           /* 0 */
   int input_size_min;         /* 1536 */
   int input_size_max;         /* 1536 */
   int output_dims;            /* 3  => probability index 1, stride 2 (vadc.c:704-708) */
   int sm_count;
   int max_streams;
   int window_chunks;
} silero_b200_info;

/* Create an engine from a .testtensor weights blob (tensor.h:201-253; 99 tensors, silero.h:31-33). */
int silero_b200_create( const void *testtensor_bytes, size_t nbytes, const silero_b200_opts *opts, silero_b200 **out );
int silero_b200_create_from_file( const char *path, const silero_b200_opts *opts, silero_b200 **out );
void silero_b200_destroy( silero_b200 *h );
int silero_b200_get_info( const silero_b200 *h, silero_b200_info *info );
/* thread-local description of the last failure in the calling thread */
const char *silero_b200_last_error( void );

/* ---- single stream, stateful: backend_run / silero_run_one_batch_with_context ----------------
   samples: host f32 [nchunks][1536] in [-1,1) (consecutive chunks of stream `stream`);
   out: host f32 [nchunks][2] (index 1 = speech probability). State of `stream` advances. */
int silero_b200_run_chunks( silero_b200 *h, int stream, const float *samples, int nchunks, float *out );

/* ---- many streams: the multi-stream chunk scheduler ------------------------------------------
   pcm: HOST s16le, stream s starts at pcm + s*stream_stride (in samples), each holding at least
   nchunks*1536 samples; streams first_stream..first_stream+nstreams-1 of the handle are advanced.
   probs: host f32 [nstreams][nchunks] speech probabilities (may be NULL);
   out2:  host f32 [nstreams][nchunks][2] both decoder heads (may be NULL).
   s16 -> f32 is x/32768.0f exactly as vadc.c:884,898. Copies are pipelined with compute. */
int silero_b200_run_streams( silero_b200 *h, const int16_t *pcm, long long stream_stride,
                             int first_stream, int nstreams, int nchunks, float *probs, float *out2 );

/* Same, with pcm/probs/out2 already resident in DEVICE memory (no copies inside the call).
   Asynchronous on the engine's stream; silero_b200_sync waits for completion. */
int silero_b200_run_streams_device( silero_b200 *h, const int16_t *d_pcm, long long stream_stride,
                                    int first_stream, int nstreams, int nchunks, float *d_probs, float *d_out2 );
int silero_b200_sync( silero_b200 *h );

/* ---- on-device segmenter: feed_probability / combine_or_emit_speech_segment / end-of-stream logic
   (vadc.c:165-299, 1005-1027) as a per-stream scan on the GPU (vadc_b200/csrc/segment_kernel.cuh). Each of the
   handle's streams carries its own FeedState + buffered candidate across calls, like its LSTM state, so only
   finished (start_chunk, end_chunk) pairs leave the device instead of the [streams][chunks] probabilities.
   Pairs are bit-identical to vadc_segmenter_feed/finish on the same probabilities; format them with
   vadc_segment_format. configure(NULL) = the reference's option defaults; configure resets every stream's
   segmenter state (not its LSTM state). */
int silero_b200_segments_configure( silero_b200 *h, const vadc_seg_params *params );
int silero_b200_segments_reset( silero_b200 *h, int first_stream, int nstreams );
/* run_streams + segmentation. segs: host [nstreams][cap] pairs finished by THIS call, counts: host [nstreams]
   (a count above cap means the excess pairs were dropped; nchunks/(min_speech+min_silence chunks)+2 always suffices).
   end_of_stream != 0 additionally closes an open segment and flushes the buffered one (nchunks may then be 0).
   probs (optional, may be NULL): host f32 [nstreams][nchunks]. */
int silero_b200_run_streams_segments( silero_b200 *h, const int16_t *pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                      int end_of_stream, vadc_segment *segs, int cap, int *counts, float *probs );
/* Asynchronous form of silero_b200_run_streams_segments: enqueues the copies and kernels and returns a ticket; the outputs are
   valid, and pcm may be reused, after silero_b200_wait(h, ticket). Up to 4 calls may be in flight; they execute in submission
   order, and the H2D copies of a call overlap the compute of the previous one (this is what lets a steady stream of calls run at
   the PCIe bound). segs/counts may be NULL (probabilities only). pcm and the outputs should be pinned host memory. */
int silero_b200_submit_streams_segments( silero_b200 *h, const int16_t *pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                         int end_of_stream, vadc_segment *segs, int cap, int *counts, float *probs, unsigned long long *ticket );
int silero_b200_wait( silero_b200 *h, unsigned long long ticket );
/* same with every buffer in DEVICE memory; asynchronous (silero_b200_sync). d_probs may be NULL (internal scratch). */
int silero_b200_run_streams_segments_device( silero_b200 *h, const int16_t *d_pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                             int end_of_stream, float *d_probs, vadc_segment *d_segs, int cap, int *d_counts );
/* the segmenter alone on device-resident probabilities: stream s, chunk n at d_probs[s*stride + n] */
int silero_b200_segment_probs_device( silero_b200 *h, const float *d_probs, long long stride, int first_stream, int nstreams, int nchunks,
                                      int end_of_stream, vadc_segment *d_segs, int cap, int *d_counts );

/* ---- several GPUs, one host process: the stream scheduler across the devices of a box (vadc_b200/csrc/group.c) ----------------
   A group owns one engine per device and one host thread per engine. Global stream s lives on device s / streams_per_device
   (= ceil(max_streams / ndevices)) for its whole life, LSTM and segmenter state included. A call fans the caller's stream range out
   to the devices that own a part of it; every device reads and writes ITS slice of the caller's host buffers in place (same
   layouts as silero_b200_run_streams_segments), so the per-stream results arrive gathered in the caller's arrays. No collective,
   no device-to-device traffic. opts->max_streams is the TOTAL over all devices; opts->device is ignored.
   Calls on one group must not overlap. Errors: silero_b200_group_last_error (thread-local). */
#define SILERO_B200_GROUP_MAX_DEVICES 16
typedef struct silero_b200_group silero_b200_group;
int silero_b200_group_create( const void *testtensor_bytes, size_t nbytes, const int *devices, int ndevices, const silero_b200_opts *opts, silero_b200_group **out );
int silero_b200_group_create_from_file( const char *path, const int *devices, int ndevices, const silero_b200_opts *opts, silero_b200_group **out );
void silero_b200_group_destroy( silero_b200_group *g );
int silero_b200_group_get_info( const silero_b200_group *g, int *ndevices, int *streams_per_device, int *max_streams );
int silero_b200_group_run_streams_segments( silero_b200_group *g, const int16_t *pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                            int end_of_stream, vadc_segment *segs, int cap, int *counts, float *probs );
int silero_b200_group_reset( silero_b200_group *g, int first_stream, int nstreams );            /* LSTM and segmenter state */
int silero_b200_group_segments_configure( silero_b200_group *g, const vadc_seg_params *params ); /* NULL: the reference's defaults */
const char *silero_b200_group_last_error( void );

/* per-stream LSTM state (zero after create / reset); h_out,c_out: host f32 [128] = [2][64] */
int silero_b200_reset( silero_b200 *h, int first_stream, int nstreams );
int silero_b200_get_state( silero_b200 *h, int stream, float *h_out, float *c_out );
int silero_b200_set_state( silero_b200 *h, int stream, const float *h_in, const float *c_in );

/* ---- device helpers for callers that keep audio in HBM (bench, pipelines) --------------------- */
int silero_b200_device_alloc( silero_b200 *h, size_t nbytes, void **d_ptr );
int silero_b200_device_free( silero_b200 *h, void *d_ptr );
int silero_b200_memcpy_h2d( silero_b200 *h, void *d_dst, const void *src, size_t nbytes );
int silero_b200_memcpy_d2h( silero_b200 *h, void *dst, const void *d_src, size_t nbytes );
int silero_b200_host_alloc_pinned( size_t nbytes, void **ptr );
int silero_b200_host_free_pinned( void *ptr );
/* device time (ms) of the most recent run_streams[_device] call, and of its kernels by stage:
   ms[0]=total, [1]=stft, [2]=layer1, [3]=layer2, [4]=layer3, [5]=layer4, [6]=lstm0, [7]=lstm1+decoder;
   kernel_launches = number of kernels launched by that call. Valid after silero_b200_sync. */
int silero_b200_last_timing( silero_b200 *h, float ms[8], long long *kernel_launches );
/* hybrid STFT statistics since the last reset: spectrogram bins produced and bins that took the exact path */
int silero_b200_stft_stats( silero_b200 *h, unsigned long long *bins_total, unsigned long long *bins_exact, int reset );
/* parity tap: the engine's expf / tanhf / log1pf(|x|) (csrc/libm_exact.cuh: glibc's algorithms, used by the fp32 LSTM path, lstm.c:64-88
   via maths.h:302-334, and by the exact STFT path, misc.c:40-46) of n host floats; must equal the C library's results bit for bit */
int silero_b200_stage_libm( silero_b200 *h, const float *x, int n, float *out_expf, float *out_tanhf, float *out_log1pf_abs );
/* test hook for the failure path of the few-streams LSTM wavefront (exact_lstm_kernel<true>): stall_producer != 0 makes the layer-0
   tasks never publish their progress, spin_limit (> 0) bounds the consumers' polls. A consumer that gives up raises the engine's
   error word; the next synchronizing call (run_streams, sync, wait) returns SILERO_B200_ERR_CUDA and the streams' state is untouched.
   (0, 0) restores normal operation. */
int silero_b200_debug_wavefront( silero_b200 *h, int stall_producer, int spin_limit );
/* enable (1) / disable (0) per-stage CUDA-event timing (adds events between kernels) */
int silero_b200_set_profiling( silero_b200 *h, int enabled );

/* device-side stopwatch on the engine's stream (CUDA events): start, run any number of calls, stop */
int silero_b200_timer_start( silero_b200 *h );
int silero_b200_timer_stop( silero_b200 *h, float *ms );
/* FP32 FMA-pipe throughput this device sustains (TFLOP/s, independent FFMA chains): the roofline
   denominator for the CUDA-core kernels of this engine */
int silero_b200_measure_fp32_peak( silero_b200 *h, float *tflops );
/* the same with separately rounded multiplies and adds (independent FMUL -> FADD chains, one FLOP per instruction): the roofline
   denominator of the exact path, whose arithmetic may not be contracted into FMAs */
int silero_b200_measure_fp32_unfused_peak( silero_b200 *h, float *tflops );

/* ---- parity taps (the reference's per-stage functions; host pointers; stateless unless noted) -
   Layouts are the reference's: spectrogram [B,129,25]; layer outputs [B,16,13] [B,32,7] [B,32,7]
   [B,64,7]; lstm sequence [B,7,64]. Any output pointer may be NULL. */
/* my_stft + adaptive_audio_normalization_inplace: samples [B,1536] -> log1p spectrogram minus mean.
   If logmag_out != NULL it receives log1p(mag*2^20) before the mean subtraction. */
int silero_b200_stage_stft_norm( silero_b200 *h, const float *samples, int batch, float *norm_out, float *logmag_out );
/* my_stft alone (stft.c:226): samples [B,1536] -> magnitude [B,129,25] */
int silero_b200_stage_stft_magnitude( silero_b200 *h, const float *samples, int batch, float *mag_out );
/* the production kernels end to end from samples (STFT -> first layer with in-kernel normalization
   -> layers 2..4), every layer output tapped in the reference layout */
int silero_b200_stage_pipeline( silero_b200 *h, const float *samples, int batch, float *l1, float *l2, float *l3, float *l4 );
/* the same taps on the exact path's kernels (exact STFT -> exact_front_kernel -> exact_layer_kernel x 4, the reference's rounding
   sequence): y1 [B,16,25] = conv_block output of the first layer (conv.c:761-814), l1..l4 as above */
int silero_b200_stage_exact_pipeline( silero_b200 *h, const float *samples, int batch, float *y1, float *l1, float *l2, float *l3, float *l4 );
/* one transformer_layer on the exact path's kernels, layouts as silero_b200_stage_layer; layer 0 takes the NORMALIZED spectrogram and
   can also return its conv_block output y1 [B,16,25] (may be NULL) */
int silero_b200_stage_exact_layer( silero_b200 *h, int layer, const float *in, int batch, float *out, float *y1 );
/* encoder on the exact path's kernels from a spectrogram [B,129,25]; kind 0: log1p spectrogram (normalized inside, as in production),
   1: already normalized, 2: raw magnitude (log1p(m * 2^20) applied first, misc.c:40-46) */
int silero_b200_stage_exact_encoder( silero_b200 *h, const float *spec, int batch, int kind, float *l1, float *l2, float *l3, float *l4 );
/* softmax_inplace_stable (tensor.h:751-784) over the rows of x [rows][cols] in the exact path's arithmetic (the reference's softmax fixture) */
int silero_b200_stage_exact_softmax( silero_b200 *h, const float *x, int rows, int cols, float *out );
/* lstm_tensor_minibatched on the exact path's kernels, arguments as silero_b200_stage_lstm; wave 0: the multi-stream kernel
   (exact_lstm_kernel.cuh), 1: its two-layer wavefront launch, which serves few streams -- identical bits */
int silero_b200_stage_exact_lstm( silero_b200 *h, const float *x, int batch, const float *h0, const float *c0, float *out, float *hn, float *cn, int wave );
/* adaptive_audio_normalization_inplace on a caller-supplied magnitude spectrogram [B,129,25] */
int silero_b200_stage_norm( silero_b200 *h, const float *magnitude, int batch, float *norm_out );
/* encoder (silero_v3.c:4-64) from a normalized spectrogram, every layer's output tapped */
int silero_b200_stage_encoder( silero_b200 *h, const float *norm, int batch,
                               float *l1, float *l2, float *l3, float *l4 );
/* one transformer_layer (transformer.c:237-295), layer = 0..3, input in the reference layout
   [B,cin,T] with (cin,T) = (129,25) (16,13) (32,7) (32,7) */
int silero_b200_stage_layer( silero_b200 *h, int layer, const float *in, int batch, float *out );
/* sub-stage taps of one layer for the reference's op/block-level fixtures (conv_block conv.c:761,
   dual_head_attention transformer.c:13, layer_norm misc.c:143, transformer_block transformer.c:160,
   conv+batch_norm1d transformer.c:280-290). layer 0..3.
   entry: 0 layer input in the reference layout [B,cin,T]; 1 conv_block output [B,T,C]; 2 transformer_block output [B,T,C]
   tap:   0 layer output [B,C,TOUT]; 1 conv_block output; 2 attention output; 3 after norm1; 4 transformer_block output (all [B,T,C]) */
int silero_b200_stage_layer_tap( silero_b200 *h, int layer, int entry, int tap, const float *in, int batch, float *out );
/* lstm_tensor_minibatched (lstm.c:228) over batch*7 steps of ONE sequence from explicit state:
   x [B,7,64]; h0,c0 [2,64]; out [B,7,64]; hn,cn [2,64] */
int silero_b200_stage_lstm( silero_b200 *h, const float *x, int batch, const float *h0, const float *c0,
                            float *out, float *hn, float *cn );
/* decoder_tensor (silero_v3.c:305): in [B,64,7] -> out [B,2] */
int silero_b200_stage_decoder( silero_b200 *h, const float *in, int batch, float *out );
#ifdef __cplusplus
}
#endif
#endif /* SILERO_B200_H */
