/* oracle/silero_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the Silero VAD v3.1 (16 kHz) hot path of IntendedConsequence/vadc.
 * It restates the reference's arithmetic (including its summation orders) so that, built with
 * -ffp-contract=off against the same libm, it is bit-identical to the reference C backend
 * compiled by oracle/Makefile (tests/test_oracle_vs_ref.py pins that) and matches every golden
 * fixture the reference checks in (tests/test_oracle_fixtures.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use this library. The product (vadc_b200/) never links, imports or calls it.
 */
#ifndef SILERO_ORACLE_H
#define SILERO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct so_model so_model;

/* .testtensor container (reference tensor.h:201-253, utils.py:7-53) holding the 99 weight tensors */
so_model *so_model_load( const void *bytes, size_t nbytes );
so_model *so_model_load_file( const char *path );
void so_model_free( so_model *m );
/* access tensor i of the loaded container: returns data pointer, fills ndim/dims (up to 8) */
const float *so_model_tensor( const so_model *m, int index, int *ndim, int *dims );
int so_model_tensor_count( const so_model *m );

/* ---- op level (one batch item unless stated); shapes follow the reference ------------------- */
void so_reflect_pad( const float *x, int n, int pad_l, int pad_r, float *out );           /* tensor.h:912-958 */
void so_stft( const so_model *m, const float *x, int batch, float *out );                 /* stft.c:15-229: [B,1536]->[B,129,25] */
void so_adaptive_norm( float *x, int batch, int channels, int frames );                   /* misc.c:1-124 in place */
void so_dw_conv( const float *in, int channels, int T, const float *w, const float *b, float *out ); /* conv.c:17-113 */
void so_pw_conv( const float *in, int cin, int T, const float *w, const float *b, int cout, float *out ); /* conv.c:532-589 (variant E), out must be zeroed or hold the accumulator start */
void so_conv1x1_strided( const float *in, int cin, int T, const float *w, const float *b, int cout, int stride, float *out ); /* conv.c:597-709 */
void so_conv_block( const float *in, int cin, int T, int has_proj,
                    const float *dw_w, const float *dw_b, const float *pw_w, const float *pw_b,
                    const float *proj_w, const float *proj_b, int cout, float *out );     /* conv.c:761-814 */
void so_linear( const float *in, int rows, int k, const float *w, const float *b, int n, float *out ); /* tensor.h:675-723 */
void so_softmax_rows( float *x, int rows, int cols );                                     /* tensor.h:751-784 */
void so_layer_norm( const float *in, int rows, int features, const float *w, const float *b, float *out ); /* misc.c:143-210 */
void so_batch_norm( const float *in, int batch, int channels, int T, const float *mean, const float *var,
                    const float *w, const float *b, float *out );                         /* misc.c:221-258 */
void so_attention( const float *in, int T, int C, const float *qkv_w, const float *qkv_b,
                   const float *proj_w, const float *proj_b, float *out );                /* transformer.c:13-153, in/out [T,C] */
void so_transformer_block( const float *in, int C, int T, const float *const *w12, float *out ); /* transformer.c:160-234, in/out [C,T]; w12 order = fill_transformer_weights */
void so_lstm_seq( const float *x, int steps, int hidden, const float *h0, const float *c0,
                  const float *w, const float *b, int layers, float *out );               /* lstm.c:156-218: out = [steps,hidden] + h[layers,hidden] + c[layers,hidden] */
void so_decoder( const float *in, int batch, int channels, int T, const float *w, const float *b, int nout, float *out ); /* silero_v3.c:231-303 */

/* ---- layer / model level -------------------------------------------------------------------- */
/* transformer_layer (transformer.c:237-295) with explicit weights in fill_transformer_weights order
   (24 pointers, or 22 when has_proj==0): in [cin,T] -> out [cout, 1+(T-1)/stride] */
void so_transformer_layer_w( const float *in, int cin, int T, int cout, int stride, int has_proj,
                             const float *const *w, float *out );
/* layer index 0..3 of the loaded model: in [B,cin,T] -> out [B,cout,Tout] */
void so_transformer_layer( const so_model *m, int layer, const float *in, int batch, float *out );
void so_encoder( const so_model *m, const float *in, int batch, float *out );             /* silero_v3.c:4-64: [B,129,25]->[B,64,7] */

typedef struct so_state
{
   float h[2 * 64];
   float c[2 * 64];
} so_state;

/* silero_run_one_batch_with_context (silero_v3.c:72-215): `batch` consecutive chunks of ONE stream,
   samples [B,1536] f32 in [-1,1), out [B,2] (index 1 = speech probability); updates state. */
void so_run_chunks( const so_model *m, so_state *state, const float *samples, int batch, float *out );
/* same with every intermediate dumped (any pointer may be NULL); layouts as the reference's */
void so_run_chunks_stages( const so_model *m, so_state *state, const float *samples, int batch,
                           float *stft_out, float *norm_out, float *l1, float *l2, float *l3, float *l4,
                           float *lstm_out, float *out );
/* s16le stream -> [nchunks,2]; s16->f32 as vadc.c:873-909; trailing partial chunk dropped (vadc.c:964) */
void so_run_pcm( const so_model *m, so_state *state, const int16_t *pcm, long long nsamples, float *out );

/* ---- timestamp contract (vadc.c:165-299, 756-768, 846, 1005-1027) ----- */
typedef struct so_segment_params
{
   float min_silence_ms;   /* 200 */
   float min_speech_ms;    /* 250 */
   float threshold;        /* 0.5 */
   float neg_threshold_relative; /* 0.15 */
   float speech_pad_ms;    /* 30 */
   int centiseconds;       /* 0: "%.2f,%.2f\n"; 1: centisecond integers */
} so_segment_params;

void so_segment_params_default( so_segment_params *p );
/* probabilities (speech channel) of one stream -> the exact text vadc prints on stdout.
   Returns the number of bytes written (excluding NUL); text is truncated to cap-1 bytes. */
size_t so_segments_text( const float *prob, long long nchunks, const so_segment_params *p, char *text, size_t cap );
/* same state machine, emitting merged (start_chunk,end_chunk) pairs; returns the pair count */
long long so_segments_chunks( const float *prob, long long nchunks, const so_segment_params *p,
                              int *pairs, long long max_pairs );

#ifdef __cplusplus
}
#endif
#endif
