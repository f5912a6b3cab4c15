/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C-ABI wrapper around the UNMODIFIED reference C backend. The reference sources are
 * #included where they lie under /root/reference (unity build, exactly like the reference's own
 * vadc.c:15-19 / silero.h:7-19); nothing is copied into this repo. Built by oracle/Makefile into
 * oracle/_ref/libvadc_ref.so with the pinned flags
 *     -O2 -mavx2 -ffp-contract=off -DNDEBUG -DONNX_INFERENCE_ENABLED=0
 * (closest to upstream MSVC /O2 /arch:AVX2, build_msvc.bat:43,68).
 *
 * Entry points wrap:
 *   vadc_ref_run      -> silero_run_one_batch_with_context   (silero_v3.c:72)
 *   vadc_ref_stages   -> the same stage sequence, dumping every intermediate tensor
 *                        (my_stft stft.c:226, adaptive_audio_normalization_inplace misc.c:1,
 *                         transformer_layer transformer.c:237, lstm_tensor_minibatched lstm.c:228,
 *                         decoder_tensor silero_v3.c:305)
 */
#include <stdlib.h>
#include <tracy/TracyC.h>

#include "vadc.h"
#include "silero.h"

#define MEMORY_IMPLEMENTATION
#include "memory.h"

typedef struct RefHandle
{
   MemoryArena arena;
   Silero_Context *ctx;
   Silero_Config config;
} RefHandle;

#define EXPORT __attribute__((visibility("default")))

EXPORT void *vadc_ref_create( void )
{
   RefHandle *h = calloc( 1, sizeof( RefHandle ) );
   size_t cap = Megabytes( 1024 );
   u8 *base = malloc( cap );
   if ( !h || !base ) return 0;
   initializeMemoryArena( &h->arena, base, cap );
   String8 nopath = {0};
   h->ctx = backend_init( &h->arena, nopath, &h->config );
   return h;
}

EXPORT void vadc_ref_destroy( void *handle )
{
   RefHandle *h = handle;
   if ( !h ) return;
   free( h->arena.base );
   free( h );
}

/* silero.h:39-43 contract values, for the boundary test */
EXPORT void vadc_ref_config( void *handle, int *out5 )
{
   RefHandle *h = handle;
   out5[0] = h->config.batch_size_restriction;
   out5[1] = h->config.is_silero_v5;
   out5[2] = h->config.input_size_min;
   out5[3] = h->config.input_size_max;
   out5[4] = h->config.output_dims;
}

EXPORT void vadc_ref_reset( void *handle )
{
   RefHandle *h = handle;
   memset( h->ctx->state_lstm_h->data, 0, h->ctx->state_lstm_h->nbytes );
   memset( h->ctx->state_lstm_c->data, 0, h->ctx->state_lstm_c->nbytes );
}

EXPORT void vadc_ref_get_state( void *handle, float *h_out, float *c_out )
{
   RefHandle *h = handle;
   memcpy( h_out, h->ctx->state_lstm_h->data, 128 * sizeof( float ) );
   memcpy( c_out, h->ctx->state_lstm_c->data, 128 * sizeof( float ) );
}

EXPORT void vadc_ref_set_state( void *handle, const float *h_in, const float *c_in )
{
   RefHandle *h = handle;
   memcpy( h->ctx->state_lstm_h->data, h_in, 128 * sizeof( float ) );
   memcpy( h->ctx->state_lstm_c->data, c_in, 128 * sizeof( float ) );
}

/* [batch,1536] f32 -> [batch,2]; carries LSTM state exactly like backend_run (silero.h:53-74) */
EXPORT int vadc_ref_run( void *handle, const float *samples, int batch, float *out )
{
   RefHandle *h = handle;
   /* the returned tensor lives above the callee's own mark (silero_v3.c:80-82): bracket the call
      so a long-running oracle does not exhaust the arena (SURVEY.md section 8b "Data ownership") */
   TemporaryMemory mark = beginTemporaryMemory( &h->arena );
   TestTensor *o = silero_run_one_batch_with_context( &h->arena, h->ctx, batch, 1536, (float *)samples );
   memcpy( out, o->data, sizeof( float ) * 2 * batch );
   endTemporaryMemory( mark );
   return 0;
}

/* Whole stream helper: s16le -> probabilities, batch 96 like vadc.c:715,1116, partial trailing
   chunk dropped like vadc.c:964. out is [nchunks,2]. */
EXPORT int vadc_ref_run_pcm( void *handle, const short *pcm, long long nsamples, int batch, float *out )
{
   RefHandle *h = handle;
   long long nchunks = nsamples / 1536;
   float *buf = malloc( sizeof( float ) * 1536 * (size_t)batch );
   float *o = malloc( sizeof( float ) * 2 * (size_t)batch );
   for ( long long c0 = 0; c0 < nchunks; c0 += batch )
   {
      long long n = nchunks - c0 < batch ? nchunks - c0 : batch;
      memset( buf, 0, sizeof( float ) * 1536 * (size_t)batch );
      for ( long long i = 0; i < n * 1536; ++i )
      {
         float v = pcm[c0 * 1536 + i];
         buf[i] = v / 32768.0f; /* vadc.c:884,898 */
      }
      /* only the n valid chunks are run so that state does not advance through padding;
         identical to vadc.c for every chunk the CLI consumes (vadc.c:964) */
      vadc_ref_run( h, buf, (int)n, o );
      memcpy( out + c0 * 2, o, sizeof( float ) * 2 * (size_t)n );
   }
   free( buf );
   free( o );
   return 0;
}

/* Stage-by-stage run of `batch` consecutive chunks. Any output pointer may be NULL.
   Layouts are the reference's: stft/norm [B,129,25]; l1 [B,16,13]; l2 [B,32,7]; l3 [B,32,7];
   l4 [B,64,7]; lstm [B,7,64]; probs [B,2]. Carries and updates the handle's LSTM state. */
EXPORT int vadc_ref_stages( void *handle, const float *samples, int batch,
                            float *stft_out, float *norm_out,
                            float *l1_out, float *l2_out, float *l3_out, float *l4_out,
                            float *lstm_out, float *probs )
{
   RefHandle *h = handle;
   MemoryArena *arena = &h->arena;
   Silero_Context *ctx = h->ctx;
   TemporaryMemory mark = beginTemporaryMemory( arena );

   TestTensor *input = tensor_zeros_2d( arena, batch, 1536 );
   memcpy( input->data, samples, sizeof( float ) * 1536 * batch );

   TestTensor *stft = tensor_zeros_3d( arena, batch, 129, 25 );
   my_stft( arena, input, ctx->weights.forward_basis_buffer, stft, 64, 128 );
   if ( stft_out ) memcpy( stft_out, stft->data, stft->nbytes );

   TestTensor *norm = tensor_copy( arena, stft );
   adaptive_audio_normalization_inplace( arena, norm );
   if ( norm_out ) memcpy( norm_out, norm->data, norm->nbytes );

   Encoder_Weights ew = ctx->weights.encoder_weights;
   TestTensor *l1 = tensor_zeros_3d( arena, batch, 16, 13 );
   TestTensor *l2 = tensor_zeros_3d( arena, batch, 32, 7 );
   TestTensor *l3 = tensor_zeros_3d( arena, batch, 32, 7 );
   TestTensor *l4 = tensor_zeros_3d( arena, batch, 64, 7 );
   transformer_layer( arena, norm, ew.l1, ew.l1_conv_stride, l1 );
   transformer_layer( arena, l1, ew.l2, ew.l2_conv_stride, l2 );
   transformer_layer( arena, l2, ew.l3, ew.l3_conv_stride, l3 );
   transformer_layer( arena, l3, ew.l4, ew.l4_conv_stride, l4 );
   if ( l1_out ) memcpy( l1_out, l1->data, l1->nbytes );
   if ( l2_out ) memcpy( l2_out, l2->data, l2->nbytes );
   if ( l3_out ) memcpy( l3_out, l3->data, l3->nbytes );
   if ( l4_out ) memcpy( l4_out, l4->data, l4->nbytes );

   TestTensor *l4_t = tensor_transpose_last_2d( arena, l4 );
   LSTM_Result lstm_out_t = lstm_tensor_minibatched( arena, l4_t, ctx->weights.lstm_weights, ctx->weights.lstm_biases,
                                                     ctx->state_lstm_h, ctx->state_lstm_c );
   if ( lstm_out ) memcpy( lstm_out, lstm_out_t.output.data, lstm_out_t.output.nbytes );
   TestTensor *lstm_t = tensor_transpose_last_2d( arena, &lstm_out_t.output );
   memmove( ctx->state_lstm_h->data, lstm_out_t.hn.data, lstm_out_t.hn.nbytes );
   memmove( ctx->state_lstm_c->data, lstm_out_t.cn.data, lstm_out_t.cn.nbytes );

   TestTensor *dec = tensor_zeros_3d( arena, batch, 2, 1 );
   decoder_tensor( arena, lstm_t, ctx->weights.decoder_weights, ctx->weights.decoder_biases, dec );
   if ( probs ) memcpy( probs, dec->data, dec->nbytes );

   endTemporaryMemory( mark );
   return 0;
}

/* encoder only (stateless): normalized spectrogram [B,129,25] -> [B,64,7] (silero_v3.c:4-64) */
EXPORT int vadc_ref_encoder( void *handle, const float *norm_in, int batch, float *l4_out )
{
   RefHandle *h = handle;
   MemoryArena *arena = &h->arena;
   TemporaryMemory mark = beginTemporaryMemory( arena );
   TestTensor *in = tensor_zeros_3d( arena, batch, 129, 25 );
   memcpy( in->data, norm_in, in->nbytes );
   TestTensor *out = tensor_zeros_3d( arena, batch, 64, 7 );
   encoder( arena, in, h->ctx->weights.encoder_weights, out );
   memcpy( l4_out, out->data, out->nbytes );
   endTemporaryMemory( mark );
   return 0;
}
