/* include/vadc_segmenter.h -- probability -> speech-segment contract of vadc, host C.
 *
 * Replaces (same thresholds, same integer/float arithmetic, same text):
 *   feed_probability                vadc.c:165-221
 *   combine_or_emit_speech_segment  vadc.c:262-299
 *   emit_speech_segment             vadc.c:223-260   ("%.2f,%.2f\n" or centiseconds)
 *   end-of-stream logic             vadc.c:1005-1027
 *   chunk-duration maths            vadc.c:756-768, 846; option defaults vadc.c:1110-1124, 1244
 * It is a streaming state machine: feed any number of per-chunk probabilities at a time (one
 * instance per audio stream), collect finished segments, call finish at end of stream.
 * Implemented in vadc_b200/csrc/segmenter.c, exported from libsilero_b200.so.
 */
#ifndef VADC_SEGMENTER_H
#define VADC_SEGMENTER_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vadc_seg_params
{
   float min_silence_ms;          /* --min_silence            200  */
   float min_speech_ms;           /* --min_speech             250  */
   float threshold;               /* --threshold              0.5  */
   float neg_threshold_relative;  /* --neg_threshold_relative 0.15 */
   float speech_pad_ms;           /* --speech_pad             30   */
   int chunk_samples;             /* --sequence_count, clamped to 1536 by the C backend (silero.h:41-42) */
   int centiseconds;              /* --output_centi_seconds */
} vadc_seg_params;

typedef struct vadc_segment
{
   int start_chunk; /* first speech chunk (global chunk index) */
   int end_chunk;   /* chunk index where silence began */
} vadc_segment;

typedef struct vadc_segmenter
{
   vadc_seg_params p;
   /* derived (vadc.c:756-768, 846, 1244) */
   int min_speech_chunks;
   int min_silence_chunks;
   float neg_threshold;
   float seconds_per_chunk;
   /* FeedState (vadc.h:110-115) */
   int temp_end;
   int current_speech_start;
   int triggered;
   /* one buffered candidate (vadc.c:831) */
   vadc_segment buffered;
   int buffered_valid;
   int global_chunk_index;
} vadc_segmenter;

void vadc_seg_params_default( vadc_seg_params *p );
void vadc_segmenter_init( vadc_segmenter *s, const vadc_seg_params *p );

/* Feed `n` consecutive probabilities. Finished (merged) segments are appended to out[0..cap).
   Returns the number of segments produced by this call (may exceed cap; only cap are stored). */
long long vadc_segmenter_feed( vadc_segmenter *s, const float *prob, long long n, vadc_segment *out, long long cap );
/* End of stream: closes a still-open segment per vadc.c:1005-1021 and flushes the buffered one. */
long long vadc_segmenter_finish( vadc_segmenter *s, vadc_segment *out, long long cap );

/* Padded times in seconds exactly as emit_speech_segment computes them (fp32). */
void vadc_segment_times( const vadc_segmenter *s, vadc_segment seg, float *start_s, float *end_s );
/* The exact stdout line vadc prints for `seg`; returns the byte count (excluding NUL). */
int vadc_segment_format( const vadc_segmenter *s, vadc_segment seg, char *buf, size_t cap );

/* Convenience: whole stream -> stdout text of the reference CLI. Returns bytes written (excl. NUL). */
size_t vadc_segments_text( const float *prob, long long nchunks, const vadc_seg_params *p, char *text, size_t cap );

/* Deterministic synthetic 16 kHz s16le test audio (bench/test input):
   speech-like harmonic bursts with vibrato and syllabic amplitude modulation, pauses, and a
   Gaussian noise floor (sigma ~ 0.003 FS). kind: 0 = bursts+noise, 1 = all zero, 2 = full-scale white noise. */
void vadc_synth_pcm( unsigned long long seed, int kind, long long nsamples, short *out );

#ifdef __cplusplus
}
#endif
#endif /* VADC_SEGMENTER_H */
