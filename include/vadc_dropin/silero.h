/* include/vadc_dropin/silero.h -- drop-in replacement for the reference's silero.h.
 *
 * The reference selects its inference backend at compile time: vadc.c:15-19 #includes either
 * onnx_helpers.c or "silero.h", and expects three functions (silero.h:48-81):
 *
 *     void *backend_init( MemoryArena *arena, String8 model_path_arg, Silero_Config *config );
 *     void  backend_run( MemoryArena *arena, void *context_ (VADC_Context*), Silero_Config config );
 *     void  backend_create_tensors( Silero_Config config, void *backend, Tensor_Buffers buffers );
 *
 * Put this directory in front of the reference tree on the include path and build the reference's
 * UNMODIFIED vadc.c with -DONNX_INFERENCE_ENABLED=0: it then runs on the B200 engine through the C
 * ABI of include/silero_b200.h (link libsilero_b200.so). See INTEGRATION.md for the exact command.
 *
 * Same contract as the reference backend: batch_size_restriction=-1, is_silero_v5=false,
 * input_size_min=max=1536, output_dims=3 (silero.h:39-43); input in buffers.input_samples
 * (f32 [batch][1536]), output in buffers.output ([batch][2], index 1 = speech, vadc.c:704-708);
 * LSTM state carried inside the backend across calls (silero_v3.c:178-179); backend_init returns
 * NULL on failure (vadc.c:692-695). The caller's MemoryArena is not used.
 *
 * Weights: --model <file.testtensor> if given (the reference's C backend ignores the argument,
 * silero.h:23), else $VADC_B200_WEIGHTS, else VADC_B200_DEFAULT_WEIGHTS (a compile-time path).
 */
#pragma once

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "silero_b200.h"

#ifndef VADC_B200_DEFAULT_WEIGHTS
#define VADC_B200_DEFAULT_WEIGHTS "silero_v31_16k.testtensor"
#endif

static inline void *backend_init( MemoryArena *arena, String8 model_path_arg, Silero_Config *config )
{
   (void)arena;
   char path[4096];
   const char *env = getenv( "VADC_B200_WEIGHTS" );
   if ( model_path_arg.size > 0 && model_path_arg.size < (strSize)sizeof( path ) )
   {
      memcpy( path, model_path_arg.begin, (size_t)model_path_arg.size );
      path[model_path_arg.size] = 0;
   }
   else
      snprintf( path, sizeof( path ), "%s", env ? env : VADC_B200_DEFAULT_WEIGHTS );

   silero_b200_opts opts;
   silero_b200_default_opts( &opts );
   const char *dev = getenv( "VADC_B200_DEVICE" );
   if ( dev ) opts.device = atoi( dev );
   opts.max_streams = 1; /* vadc processes one stream */

   silero_b200 *engine = 0;
   if ( silero_b200_create_from_file( path, &opts, &engine ) != SILERO_B200_OK )
   {
      fprintf( stderr, "silero_b200: %s\n", silero_b200_last_error() );
      return 0;
   }
   silero_b200_info info;
   silero_b200_get_info( engine, &info );
   config->batch_size_restriction = info.batch_size_restriction;
   config->is_silero_v5 = info.is_silero_v5;
   config->input_size_min = info.input_size_min;
   config->input_size_max = info.input_size_max;
   config->output_dims = info.output_dims;
   return engine;
}

static inline void backend_run( MemoryArena *arena, void *context_, Silero_Config config )
{
   (void)arena;
   VADC_Context *context = (VADC_Context *)context_;
   /* backend_run has no error channel (void, silero.h:53); like the reference's other backend
      (onnx_helpers.h:5-14) a failed run aborts */
   if ( silero_b200_run_chunks( (silero_b200 *)context->backend, 0, context->buffers.input_samples, config.batch_size,
                                context->buffers.output ) != SILERO_B200_OK )
   {
      fprintf( stderr, "silero_b200: %s\n", silero_b200_last_error() );
      abort();
   }
}

static inline void backend_create_tensors( Silero_Config config, void *backend, Tensor_Buffers buffers )
{
   (void)config;
   (void)backend;
   (void)buffers;
}
