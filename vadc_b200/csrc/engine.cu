// vadc_b200/csrc/engine.cu -- the C-ABI engine: weight packing, device memory, the multi-stream
// chunk scheduler (windows of chunks x streams), and kernel launches. See include/silero_b200.h.
//
// Replaces the orchestration of silero_run_one_batch_with_context (silero_v3.c:72-215), backend_run
// (silero.h:53-74), process_chunks (vadc.c:56-103) and the s16->f32 step of run_inference
// (vadc.c:873-909). No CPU fallback exists: without a CUDA device every entry point fails.
#include "silero_b200.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "faithful_kernel.cuh"
#include "exact_lstm_kernel.cuh"
#include "exact_encoder_kernel.cuh"
#include "layer_kernel.cuh"
#include "layer_tc_kernel.cuh"
#include "layer0_tc_kernel.cuh"
#include "lstm_kernel.cuh"
#include "lstm_tc_kernel.cuh"
#include "segment_kernel.cuh"
#include "stft_fft8_kernel.cuh"
#include "stft_kernel.cuh"
#include "stft_sym_kernel.cuh"
#include "testtensor.h"
#include "vadc_segmenter.h"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err( int code, const char *fmt, ... )
{
   va_list ap;
   va_start( ap, fmt );
   vsnprintf( g_err, sizeof( g_err ), fmt, ap );
   va_end( ap );
   return code;
}

extern "C" const char *silero_b200_last_error( void ) { return g_err; }

#define CU( call )                                                                                          \
   do                                                                                                       \
   {                                                                                                        \
      cudaError_t e_ = ( call );                                                                            \
      if ( e_ != cudaSuccess )                                                                              \
         return set_err( SILERO_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString( e_ ), __FILE__, __LINE__ ); \
   } while ( 0 )

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
#define N_STAGE_EVENTS 9

struct silero_b200
{
   int device;
   int sm_count;
   int max_streams;
   int window_chunks_opt;
   int stft_mode;            // kernel family: 0 = FFT hybrid (stft_fft8_kernel), SILERO_B200_STFT_EXACT, _HYBRID_FFT, _HYBRID_TENSOR
   int stft_auto;            // SILERO_B200_STFT_AUTO: run_window picks the exact kernel for small stream batches (the fp32 path)
   float stft_k_rel;         // hybrid: exact re-evaluation below k_rel * ||frame||
   int lstm_mode;            // SILERO_B200_LSTM_*
   unsigned char *d_lstm_tc; // [2 layers][LTC_W_BYTES] bf16 hi/lo weight images (lstm_tc_kernel.cuh)
   int layer_mode;           // SILERO_B200_LAYERS_*
   int faithful;             // != 0: the exact path (reference's rounding sequence) for every call; fixed at creation
   L0DwParams l0_dw;         // first layer's depthwise taps, passed to layer0_tc_kernel by value
   unsigned char *d_layer_tc[4]; // fp16 hi/lo weight images + fp32 parameters per layer (layer0_tc_kernel.cuh, layer_tc_kernel.cuh)
   size_t cap_h0_floats;
   unsigned long long *d_flagged; // bins that took the exact path (device counter)
   size_t scratch_budget;    // bytes of window scratch pick_window may plan for (set from the device's free memory at creation)
   unsigned long long bins_total;
   cudaStream_t stream;      // compute
   cudaStream_t copy_stream; // H2D of the next window
   cudaStream_t front_stream; // STFT of window w+1 while layers 2..4 and the LSTM of window w run (run_window)
   cudaEvent_t ev_spec_ready, ev_spec_free; // spectrogram buffer hand-offs between front_stream and stream
   float *d_weights;         // one allocation holding every packed weight
   DeviceWeights w;
   fq::Weights fw;           // the container's 99 tensors as they are (faithful_kernel.cuh)
   int *d_lstm_sync;         // [1 + groups]: role ticket + per-group layer-0 progress of the exact_lstm_kernel wavefront
   int *err_word;            // mapped pinned host word a kernel raises when it gives up (wavefront consumer whose producer is lost)
   int *d_err_word;          // its device address
   int wave_spin_limit;      // polls before a wavefront consumer gives up (debug taps lower it)
   int token_min_chunks;     // windows of at least this many chunks take the thread-per-token encoder (SILERO_B200_EXACT_TOKEN_MIN_CHUNKS; $SILERO_B200_TOKEN_MIN_CHUNKS for measurements)
   int debug_stall_producer; // test hook: layer-0 tasks never publish progress
   float *state_h, *state_c; // [max_streams][2][64]
   // window scratch (grow-only)
   size_t cap_chunks;
   float *spec, *a1, *a2, *a3, *a4, *h0, *mu;
   int stft_sym;             // the basis has the mirror property: the exact STFT runs on stft_sym_kernel (half the arithmetic, same bits)
   const float *basis_sym;   // [64 kquads][128 rows][4] (stft_sym_kernel.cuh)
   float *y1;                // [chunks][16][25] conv_block output of the first layer (exact_front_kernel)
   float *xe_scratch;        // per-CTA token scratch of exact_layer_kernel
   // host-call staging
   int16_t *pcm_stage[2];
   size_t pcm_stage_cap; // samples per buffer
   cudaEvent_t pcm_ready[2], pcm_free[2];
   unsigned stage_seq;       // windows staged so far: window k of the engine's lifetime uses staging buffer k & 1
   int stage_used[2];        // pcm_free[b] has been recorded since the staging buffers were (re)allocated
   cudaEvent_t done_ev[4];   // completion tickets of asynchronous calls (silero_b200_submit_* / silero_b200_wait)
   unsigned long long ticket_seq;
   float *d_probs;
   size_t d_probs_cap;
   float *d_out2;
   size_t d_out2_cap;
   float *d_f32;
   size_t d_f32_cap;
   // on-device segmenter (segment_kernel.cuh): per-stream FeedState + buffered candidate, output staging
   SegStateDev *d_seg_state; // [max_streams]
   SegParamsDev seg_params;
   SegPair *d_segs;
   size_t d_segs_cap;
   int *d_counts;
   size_t d_counts_cap;
   // timing
   int profiling;
   cudaEvent_t ev_begin, ev_end;
   cudaEvent_t ev_stage[N_STAGE_EVENTS];
   float stage_ms[8];
   long long launches;
   int timing_valid;
};

static int use_device( const silero_b200 *h )
{
   CU( cudaSetDevice( h->device ) );
   return 0;
}

// ---------------------------------------------------------------------------------------------
// weight packing (host)
// ---------------------------------------------------------------------------------------------
static void pack_basis( const float *basis /*[258][256]*/, float *out /*[2][64][128][4]*/ )
{
   for ( int half = 0; half < 2; ++half )
      for ( int rho = 0; rho < 128; ++rho )
      {
         int row;
         if ( half == 0 )
            row = rho < 64 ? rho : ( rho == 64 ? 128 : 129 + ( rho - 64 ) );
         else
            row = rho < 64 ? 64 + rho : 129 + rho;
         for ( int l = 0; l < 8; ++l )
            for ( int g = 0; g < 4; ++g )
               for ( int v = 0; v < 8; ++v )
               {
                  int k = 64 * g + 8 * v + l;
                  int kq = l * 8 + g * 2 + ( v >> 2 );
                  out[( ( (size_t)half * 64 + kq ) * 128 + rho ) * 4 + ( v & 3 )] = basis[(size_t)row * 256 + k];
               }
      }
}

// stft_sym_kernel.cuh: does the table have the mirror property the kernel builds on? (value comparisons: +0 == -0)
static bool basis_is_mirrored( const float *basis /*[258][256]*/ )
{
   for ( int k = 0; k < 256; ++k )
   {
      const float s = ( k & 1 ) ? -1.0f : 1.0f;
      for ( int f = 0; f <= 128; ++f )
         if ( basis[(size_t)( 128 - f ) * 256 + k] != s * basis[(size_t)f * 256 + k] ) return false;
      for ( int f = 1; f < 128; ++f )
         if ( basis[(size_t)( 129 + 128 - f ) * 256 + k] != -s * basis[(size_t)( 129 + f ) * 256 + k] ) return false;
      if ( basis[(size_t)129 * 256 + k] != 0.0f || basis[(size_t)257 * 256 + k] != 0.0f ) return false;
      if ( ( k & 1 ) ? basis[(size_t)64 * 256 + k] != 0.0f : basis[(size_t)( 129 + 64 ) * 256 + k] != 0.0f ) return false;
   }
   return true;
}

// rows of stft_sym_kernel.cuh: slot a * 64 + u; u >= 1: (re, im) of bin u; u == 0: re of bin 0 and the merged row of bin 64
static void pack_basis_sym( const float *basis /*[258][256]*/, float *out /*[64][128][4]*/ )
{
   for ( int a = 0; a < 2; ++a )
      for ( int u = 0; u < 64; ++u )
         for ( int l = 0; l < 8; ++l )
            for ( int g = 0; g < 4; ++g )
               for ( int v = 0; v < 8; ++v )
               {
                  const int k = 64 * g + 8 * v + l;
                  const int kq = l * 8 + g * 2 + ( v >> 2 );
                  int row = a == 0 ? u : 129 + u;
                  if ( u == 0 && a == 1 ) row = ( k & 1 ) ? 129 + 64 : 64;
                  out[( (size_t)kq * 128 + a * 64 + u ) * 4 + ( v & 3 )] = basis[(size_t)row * 256 + k];
               }
}

template <int L>
static void pack_layer( const vb_tensor *t, float *blob )
{
   using P = LayerPack<L>;
   constexpr int CIN = P::CIN, C = P::C, D = P::D;
   memset( blob, 0, sizeof( float ) * P::TOTAL );
   int i = 0;
   const float *dw_w = t[i++].data, *dw_b = t[i++].data, *pw_w = t[i++].data, *pw_b = t[i++].data;
   const float *proj_w = 0, *proj_b = 0;
   if ( P::PROJ )
   {
      proj_w = t[i++].data;
      proj_b = t[i++].data;
   }
   const float *qkv_w = t[i++].data, *qkv_b = t[i++].data, *ao_w = t[i++].data, *ao_b = t[i++].data;
   const float *n1w = t[i++].data, *n1b = t[i++].data, *f1w = t[i++].data, *f1b = t[i++].data;
   const float *f2w = t[i++].data, *f2b = t[i++].data, *n2w = t[i++].data, *n2b = t[i++].data;
   const float *cvw = t[i++].data, *cvb = t[i++].data;
   const float *bnw = t[i++].data, *bnb = t[i++].data, *bnm = t[i++].data, *bnv = t[i++].data;

   for ( int c = 0; c < CIN; ++c )
   {
      for ( int k = 0; k < 5; ++k ) blob[P::DW + c * 8 + k] = dw_w[c * 5 + k];
      blob[P::DW + c * 8 + 5] = dw_b[c];
   }
   if ( CIN == VB_BINS )
   {
      for ( int f = 0; f < CIN; ++f )
         for ( int o = 0; o < C; ++o )
         {
            blob[P::PW + f * 2 * C + o] = pw_w[o * CIN + f];
            blob[P::PW + f * 2 * C + C + o] = proj_w[o * CIN + f];
         }
   }
   else
   {
      for ( int o = 0; o < C; ++o )
         for ( int c = 0; c < CIN; ++c )
         {
            blob[P::PW + o * P::KP + c] = pw_w[o * CIN + c];
            if ( P::PROJ ) blob[P::PW + o * P::KP + CIN + c] = proj_w[o * CIN + c];
         }
   }
   for ( int o = 0; o < C; ++o ) blob[P::PWB + o] = P::PROJ ? pw_b[o] + proj_b[o] : pw_b[o];
   for ( int h = 0; h < 2; ++h )
   {
      float *q = blob + P::QKV + h * P::QH;
      for ( int part = 0; part < 3; ++part ) // q, k, v column blocks of the fused QKV (transformer.c:72-99)
         for ( int j = 0; j < D; ++j )
         {
            int src_row = part * C + h * D + j;
            for ( int c = 0; c < C; ++c ) q[( part * D + j ) * C + c] = qkv_w[src_row * C + c];
            q[3 * D * C + part * D + j] = qkv_b[src_row];
         }
   }
   memcpy( blob + P::AO, ao_w, sizeof( float ) * C * C );
   memcpy( blob + P::AOB, ao_b, sizeof( float ) * C );
   memcpy( blob + P::LN1W, n1w, sizeof( float ) * C );
   memcpy( blob + P::LN1B, n1b, sizeof( float ) * C );
   memcpy( blob + P::F1, f1w, sizeof( float ) * C * C );
   memcpy( blob + P::F1B, f1b, sizeof( float ) * C );
   memcpy( blob + P::F2, f2w, sizeof( float ) * C * C );
   memcpy( blob + P::F2B, f2b, sizeof( float ) * C );
   memcpy( blob + P::LN2W, n2w, sizeof( float ) * C );
   memcpy( blob + P::LN2B, n2b, sizeof( float ) * C );
   memcpy( blob + P::CV, cvw, sizeof( float ) * C * C );
   memcpy( blob + P::CVB, cvb, sizeof( float ) * C );
   for ( int o = 0; o < C; ++o )
   {
      blob[P::BNM + o] = bnm[o];
      blob[P::BNS + o] = sqrtf( bnv[o] + 1e-5f ); // misc.c:245
      blob[P::BNW + o] = bnw[o];
      blob[P::BNB + o] = bnb[o];
   }
}

static void pack_lstm( const float *w /*[2][256][128]*/, float *out /*[2][32][256][4]*/ )
{
   for ( int l = 0; l < 2; ++l )
      for ( int r = 0; r < 256; ++r )
         for ( int k = 0; k < 128; ++k ) out[( ( (size_t)l * 32 + ( k >> 2 ) ) * 256 + r ) * 4 + ( k & 3 )] = w[( (size_t)l * 256 + r ) * 128 + k];
}

// bf16 hi/lo image of one LSTM layer in the shared-memory order of lstm_tc_kernel.cuh:
// [split][K chunk of 8][row'][8], rows permuted so that a warp's 32 TMEM lanes hold (i|f) resp. (g|o) of 16 units
static void pack_lstm_tc( const float *w /*[2][256][128]*/, unsigned char *img /*[2][LTC_W_BYTES]*/ )
{
   for ( int l = 0; l < 2; ++l )
      for ( int rp = 0; rp < 256; ++rp )
      {
         const int m = rp >> 7, L = rp & 127, wq = L >> 5, q = L & 31;
         const int u = 16 * wq + ( q & 15 );
         const int gate = m == 0 ? ( q < 16 ? 0 : 1 ) : ( q < 16 ? 2 : 3 );
         const float *src = w + ( (size_t)l * 256 + gate * 64 + u ) * 128;
         for ( int k = 0; k < 128; ++k )
         {
            const __nv_bfloat16 hi = __float2bfloat16_rn( src[k] );
            const __nv_bfloat16 lo = __float2bfloat16_rn( src[k] - __bfloat162float( hi ) );
            const size_t off = (size_t)l * LTC_W_BYTES + (size_t)( k >> 3 ) * LTC_W_LBO + (size_t)rp * 16 + (size_t)( k & 7 ) * 2;
            memcpy( img + off, &hi, 2 );
            memcpy( img + off + LTC_W_SPLIT_BYTES, &lo, 2 );
         }
      }
}

// fp16 hi/lo image of one [N][K] weight matrix in the operand order of tc_common.cuh: [split][K/8][N][8]
static void pack_f16_split( const float *w, int ldw, int N, int K, unsigned char *img )
{
   const size_t split_bytes = (size_t)( K / 8 ) * N * 16;
   for ( int n = 0; n < N; ++n )
      for ( int k = 0; k < K; ++k )
      {
         const float v = w[(size_t)n * ldw + k];
         const __half hi = __float2half_rn( v );
         const __half lo = __float2half_rn( v - __half2float( hi ) );
         const size_t off = (size_t)( k >> 3 ) * N * 16 + (size_t)n * 16 + (size_t)( k & 7 ) * 2;
         memcpy( img + off, &hi, 2 );
         memcpy( img + split_bytes + off, &lo, 2 );
      }
}

// image of one layer for layer_tc_kernel<L>, built from the validated fp32 LayerPack<L> blob
template <int L>
static void pack_layer_tc( const float *blob, unsigned char *img )
{
   using P = LayerPack<L>;
   using Cfg = LtcCfg<L>;
   constexpr int C = P::C, D = P::D, CIN = P::CIN;
   memset( img, 0, Cfg::IMG_BYTES );
   pack_f16_split( blob + P::PW, P::KP, C, P::KP, img + Cfg::W_PW ); // row o = [pw_w[o][:], proj_w[o][:]]
   // fused QKV: rows in head-major order [h][q(D) k(D) v(D)], exactly the blob's per-head blocks
   {
      float *tmp = (float *)malloc( sizeof( float ) * 3 * C * C );
      for ( int h = 0; h < 2; ++h ) memcpy( tmp + (size_t)h * 3 * D * C, blob + P::QKV + h * P::QH, sizeof( float ) * 3 * D * C );
      pack_f16_split( tmp, C, 3 * C, C, img + Cfg::W_QKV );
      free( tmp );
   }
   pack_f16_split( blob + P::AO, C, C, C, img + Cfg::W_AO );
   pack_f16_split( blob + P::F1, C, C, C, img + Cfg::W_F1 );
   pack_f16_split( blob + P::F2, C, C, C, img + Cfg::W_F2 );
   pack_f16_split( blob + P::CV, C, C, C, img + Cfg::W_CV );
   float *f = reinterpret_cast<float *>( img + Cfg::W_END );
   memcpy( f + Cfg::F_DW, blob + P::DW, sizeof( float ) * CIN * 8 );
   memcpy( f + Cfg::F_PWB, blob + P::PWB, sizeof( float ) * C );
   for ( int h = 0; h < 2; ++h ) memcpy( f + Cfg::F_QKVB + h * 3 * D, blob + P::QKV + h * P::QH + 3 * D * C, sizeof( float ) * 3 * D );
   memcpy( f + Cfg::F_AOB, blob + P::AOB, sizeof( float ) * C );
   memcpy( f + Cfg::F_LN1W, blob + P::LN1W, sizeof( float ) * C );
   memcpy( f + Cfg::F_LN1B, blob + P::LN1B, sizeof( float ) * C );
   memcpy( f + Cfg::F_F1B, blob + P::F1B, sizeof( float ) * C );
   memcpy( f + Cfg::F_F2B, blob + P::F2B, sizeof( float ) * C );
   memcpy( f + Cfg::F_LN2W, blob + P::LN2W, sizeof( float ) * C );
   memcpy( f + Cfg::F_LN2B, blob + P::LN2B, sizeof( float ) * C );
   memcpy( f + Cfg::F_CVB, blob + P::CVB, sizeof( float ) * C );
   memcpy( f + Cfg::F_BNM, blob + P::BNM, sizeof( float ) * C );
   memcpy( f + Cfg::F_BNS, blob + P::BNS, sizeof( float ) * C );
   memcpy( f + Cfg::F_BNW, blob + P::BNW, sizeof( float ) * C );
   memcpy( f + Cfg::F_BNB, blob + P::BNB, sizeof( float ) * C );
}

// image of the first layer for layer0_tc_kernel, from the LayerPack<0> blob
static void pack_layer0_tc( const float *blob, unsigned char *img )
{
   using P = LayerPack<0>;
   using Cfg = L0tc;
   constexpr int C = 16, D = 8;
   memset( img, 0, Cfg::IMG_BYTES );
   float tmp[48 * 16 > 16 * 256 ? 48 * 16 : 16 * 256];
   // conv block: k = 2*f + {0: pw_w[o][f] (applied to relu(dw(x))), 1: proj_w[o][f] (applied to x)}, f < 128
   for ( int o = 0; o < C; ++o )
      for ( int f = 0; f < 128; ++f )
      {
         tmp[o * 256 + 2 * f] = blob[P::PW + f * 2 * C + o];
         tmp[o * 256 + 2 * f + 1] = blob[P::PW + f * 2 * C + C + o];
      }
   pack_f16_split( tmp, 256, C, 256, img + Cfg::W_PW );
   for ( int h = 0; h < 2; ++h ) memcpy( tmp + (size_t)h * 3 * D * C, blob + P::QKV + h * P::QH, sizeof( float ) * 3 * D * C );
   pack_f16_split( tmp, C, 3 * C, C, img + Cfg::W_QKV );
   pack_f16_split( blob + P::AO, C, C, C, img + Cfg::W_AO );
   pack_f16_split( blob + P::F1, C, C, C, img + Cfg::W_F1 );
   pack_f16_split( blob + P::F2, C, C, C, img + Cfg::W_F2 );
   pack_f16_split( blob + P::CV, C, C, C, img + Cfg::W_CV );
   float *f = reinterpret_cast<float *>( img + Cfg::W_END );
   memcpy( f + Cfg::F_DW, blob + P::DW, sizeof( float ) * 129 * 8 );
   for ( int o = 0; o < C; ++o )
   {
      f[Cfg::F_WL + o] = blob[P::PW + 128 * 2 * C + o];
      f[Cfg::F_WL + C + o] = blob[P::PW + 128 * 2 * C + C + o];
   }
   memcpy( f + Cfg::F_PWB, blob + P::PWB, sizeof( float ) * C );
   for ( int h = 0; h < 2; ++h ) memcpy( f + Cfg::F_QKVB + h * 3 * D, blob + P::QKV + h * P::QH + 3 * D * C, sizeof( float ) * 3 * D );
   memcpy( f + Cfg::F_AOB, blob + P::AOB, sizeof( float ) * C );
   memcpy( f + Cfg::F_LN1W, blob + P::LN1W, sizeof( float ) * C );
   memcpy( f + Cfg::F_LN1B, blob + P::LN1B, sizeof( float ) * C );
   memcpy( f + Cfg::F_F1B, blob + P::F1B, sizeof( float ) * C );
   memcpy( f + Cfg::F_F2B, blob + P::F2B, sizeof( float ) * C );
   memcpy( f + Cfg::F_LN2W, blob + P::LN2W, sizeof( float ) * C );
   memcpy( f + Cfg::F_LN2B, blob + P::LN2B, sizeof( float ) * C );
   memcpy( f + Cfg::F_CVB, blob + P::CVB, sizeof( float ) * C );
   memcpy( f + Cfg::F_BNM, blob + P::BNM, sizeof( float ) * C );
   memcpy( f + Cfg::F_BNS, blob + P::BNS, sizeof( float ) * C );
   memcpy( f + Cfg::F_BNW, blob + P::BNW, sizeof( float ) * C );
   memcpy( f + Cfg::F_BNB, blob + P::BNB, sizeof( float ) * C );
}

// ---------------------------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------------------------
extern "C" void silero_b200_default_opts( silero_b200_opts *o )
{
   memset( o, 0, sizeof( *o ) );
   o->device = 0;
   o->max_streams = 1;
   o->window_chunks = 0;
   o->stft_mode = SILERO_B200_STFT_AUTO;
   o->stft_k_rel = 0.0f; /* 0 = default (SILERO_B200_STFT_K_REL_DEFAULT) */
   o->lstm_mode = SILERO_B200_LSTM_AUTO;
   o->layer_mode = SILERO_B200_LAYERS_AUTO;
}

template <typename K>
static cudaError_t allow_smem( K kernel, int bytes )
{
   return cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes );
}

static int configure_kernels()
{
   CU( allow_smem( stft_logmag_kernel<false>, STFT_SMEM_BYTES ) );
   CU( allow_smem( stft_logmag_kernel<true>, STFT_SMEM_BYTES ) );
   CU( allow_smem( stft_sym_kernel<false>, SSYM_SMEM_BYTES ) );
   CU( allow_smem( stft_sym_kernel<true>, SSYM_SMEM_BYTES ) );
   CU( allow_smem( stft_fft8_kernel<false, false>, F8_SMEM_BYTES ) );
   CU( allow_smem( stft_fft8_kernel<true, false>, F8_SMEM_BYTES ) );
   CU( allow_smem( stft_fft8_kernel<false, true>, F8_SMEM_BYTES ) );
   CU( allow_smem( stft_fft8_kernel<true, true>, F8_SMEM_BYTES ) );
   CU( allow_smem( layer_kernel<0, true>, LayerCfg<0>::SMEM_BYTES ) );
   CU( allow_smem( layer_kernel<0, false>, LayerCfg<0>::SMEM_BYTES ) );
   CU( allow_smem( layer_kernel<1, false>, LayerCfg<1>::SMEM_BYTES ) );
   CU( allow_smem( layer_kernel<2, false>, LayerCfg<2>::SMEM_BYTES ) );
   CU( allow_smem( layer_kernel<3, false>, LayerCfg<3>::SMEM_BYTES ) );
   CU( allow_smem( lstm_layer_kernel<0, 4>, LstmSmem<4>::BYTES ) );
   CU( allow_smem( lstm_layer_kernel<1, 4>, LstmSmem<4>::BYTES ) );
   CU( allow_smem( lstm_layer_kernel<0, 1>, LstmSmem<1>::BYTES ) );
   CU( allow_smem( lstm_layer_kernel<1, 1>, LstmSmem<1>::BYTES ) );
   CU( allow_smem( faithful_encoder_kernel, FAITHFUL_SMEM_BYTES ) );
   CU( allow_smem( exact_front_kernel<true>, XF_SMEM_BYTES ) );
   CU( allow_smem( exact_front_kernel<false>, XF_SMEM_BYTES ) );
   CU( allow_smem( exact_layer_kernel<0>, XeCfg<0>::SMEM_BYTES ) );
   CU( allow_smem( exact_layer_kernel<1>, XeCfg<1>::SMEM_BYTES ) );
   CU( allow_smem( exact_layer_kernel<2>, XeCfg<2>::SMEM_BYTES ) );
   CU( allow_smem( exact_layer_kernel<3>, XeCfg<3>::SMEM_BYTES ) );
   CU( allow_smem( exact_lstm_kernel<false>, XL_SMEM_BYTES ) );
   CU( allow_smem( exact_lstm_kernel<true>, XL_SMEM_BYTES ) );
   CU( allow_smem( lstm_tc_kernel<0>, LTC_SMEM_BYTES ) );
   CU( allow_smem( lstm_tc_kernel<1>, LTC_SMEM_BYTES ) );
   CU( allow_smem( layer0_tc_kernel<false>, L0tc::SMEM_BYTES ) );
   CU( allow_smem( layer0_tc_kernel<true>, L0tc::SMEM_BYTES ) );
   CU( allow_smem( layer_tc_kernel<1>, LtcCfg<1>::SMEM_BYTES ) );
   CU( allow_smem( layer_tc_kernel<2>, LtcCfg<2>::SMEM_BYTES ) );
   CU( allow_smem( layer_tc_kernel<3>, LtcCfg<3>::SMEM_BYTES ) );
   return 0;
}

extern "C" void silero_b200_destroy( silero_b200 *h )
{
   if ( !h ) return;
   cudaSetDevice( h->device );
   if ( h->stream ) cudaStreamSynchronize( h->stream );
   cudaFree( h->d_weights );
   cudaFree( h->state_h );
   cudaFree( h->state_c );
   cudaFree( h->spec );
   cudaFree( h->mu );
   cudaFree( h->a1 );
   cudaFree( h->a2 );
   cudaFree( h->a3 );
   cudaFree( h->a4 );
   cudaFree( h->y1 );
   cudaFree( h->xe_scratch );
   cudaFree( h->h0 );
   cudaFree( h->d_probs );
   cudaFree( h->d_out2 );
   cudaFree( h->d_f32 );
   cudaFree( h->d_flagged );
   cudaFree( h->d_lstm_sync );
   if ( h->err_word ) cudaFreeHost( h->err_word );
   cudaFree( h->d_lstm_tc );
   for ( int i = 0; i < 4; ++i ) cudaFree( h->d_layer_tc[i] );
   cudaFree( h->d_seg_state );
   cudaFree( h->d_segs );
   cudaFree( h->d_counts );
   for ( int i = 0; i < 2; ++i )
   {
      cudaFree( h->pcm_stage[i] );
      if ( h->pcm_ready[i] ) cudaEventDestroy( h->pcm_ready[i] );
      if ( h->pcm_free[i] ) cudaEventDestroy( h->pcm_free[i] );
   }
   for ( int i = 0; i < 4; ++i )
      if ( h->done_ev[i] ) cudaEventDestroy( h->done_ev[i] );
   if ( h->ev_begin ) cudaEventDestroy( h->ev_begin );
   if ( h->ev_end ) cudaEventDestroy( h->ev_end );
   for ( int i = 0; i < N_STAGE_EVENTS; ++i )
      if ( h->ev_stage[i] ) cudaEventDestroy( h->ev_stage[i] );
   if ( h->stream ) cudaStreamDestroy( h->stream );
   if ( h->copy_stream ) cudaStreamDestroy( h->copy_stream );
   if ( h->front_stream ) cudaStreamDestroy( h->front_stream );
   if ( h->ev_spec_ready ) cudaEventDestroy( h->ev_spec_ready );
   if ( h->ev_spec_free ) cudaEventDestroy( h->ev_spec_free );
   free( h );
}

// derived segmenter constants exactly as the host state machine computes them (segmenter.c: vadc_segmenter_init)
static void set_seg_params( silero_b200 *h, const vadc_seg_params *p_in )
{
   vadc_seg_params p;
   if ( p_in )
      p = *p_in;
   else
      vadc_seg_params_default( &p );
   vadc_segmenter s;
   vadc_segmenter_init( &s, &p );
   h->seg_params.threshold = p.threshold;
   h->seg_params.neg_threshold = s.neg_threshold;
   h->seg_params.seconds_per_chunk = s.seconds_per_chunk;
   h->seg_params.pad_seconds = p.speech_pad_ms / 1000.0f; // vadc.c:231
   h->seg_params.min_speech_chunks = s.min_speech_chunks;
   h->seg_params.min_silence_chunks = s.min_silence_chunks;
   h->seg_params.chunk_samples = p.chunk_samples;
}

static int create_impl( const void *bytes, size_t nbytes, const silero_b200_opts *opts_in, silero_b200 **out )
{
   silero_b200_opts opts;
   if ( opts_in )
      opts = *opts_in;
   else
      silero_b200_default_opts( &opts );
   if ( opts.max_streams < 1 ) opts.max_streams = 1;

   vb_tensor_file tf;
   char perr[160];
   if ( vb_testtensor_parse( bytes, nbytes, &tf, perr, sizeof( perr ) ) ) return set_err( SILERO_B200_ERR_WEIGHTS, "weights: %s", perr );
   if ( vb_silero_v31_check( &tf, perr, sizeof( perr ) ) )
   {
      vb_testtensor_free( &tf );
      return set_err( SILERO_B200_ERR_WEIGHTS, "weights: %s", perr );
   }

   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount( &ndev );
   if ( e != cudaSuccess || ndev <= 0 || opts.device >= ndev )
   {
      vb_testtensor_free( &tf );
      return set_err( SILERO_B200_ERR_CUDA, "no usable CUDA device (count=%d, requested=%d): %s; this engine has no CPU fallback", ndev,
                      opts.device, cudaGetErrorString( e ) );
   }

   silero_b200 *h = (silero_b200 *)calloc( 1, sizeof( silero_b200 ) );
   if ( !h )
   {
      vb_testtensor_free( &tf );
      return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
   }
   h->device = opts.device;
   h->max_streams = opts.max_streams;
   h->window_chunks_opt = opts.window_chunks;
   h->stft_mode = opts.stft_mode == SILERO_B200_STFT_EXACT ? SILERO_B200_STFT_EXACT : 0;
   h->stft_auto = opts.stft_mode != SILERO_B200_STFT_HYBRID && h->stft_mode == 0;
   h->stft_k_rel = opts.stft_k_rel > 0.0f ? opts.stft_k_rel : SILERO_B200_STFT_K_REL_DEFAULT;
   h->lstm_mode = ( opts.lstm_mode == SILERO_B200_LSTM_FP32 || opts.lstm_mode == SILERO_B200_LSTM_TENSOR ) ? opts.lstm_mode : SILERO_B200_LSTM_AUTO;
   h->layer_mode = ( opts.layer_mode == SILERO_B200_LAYERS_FP32 || opts.layer_mode == SILERO_B200_LAYERS_TENSOR ) ? opts.layer_mode : SILERO_B200_LAYERS_AUTO;
   // The kernel family is decided HERE, once per engine, never per call: a persistent stream must not change arithmetic when the
   // caller's batch shape changes (a stream's LSTM state integrates one-ulp differences over minutes).
   //   * every mode AUTO (the default), or FAITHFUL requested: the exact path -- the reference's own rounding sequence from the STFT to
   //     the probability (stft_sym_kernel / exact_encoder_kernel / exact_lstm_kernel), for
   //     ANY number of streams; results are bit-identical to the reference whatever the batch composition;
   //   * any explicit fast mode (STFT_HYBRID*, LSTM_FP32/TENSOR, LAYERS_FP32/TENSOR): the fast family (1e-4 per chunk, drifts on long
   //     streams: DESIGN.md section 2); its remaining AUTO members are resolved from max_streams, also once.
   const bool all_auto = h->stft_auto && h->lstm_mode == SILERO_B200_LSTM_AUTO && h->layer_mode == SILERO_B200_LAYERS_AUTO;
   h->faithful = ( opts.layer_mode == SILERO_B200_LAYERS_FAITHFUL || opts.lstm_mode == SILERO_B200_LSTM_FAITHFUL || all_auto ) ? 2 : 0;
   if ( h->faithful )
   {
      h->stft_mode = SILERO_B200_STFT_EXACT;
      h->stft_auto = 0;
   }
   else
   {
      const bool big = h->max_streams >= SILERO_B200_LSTM_TENSOR_MIN_STREAMS;
      if ( h->lstm_mode == SILERO_B200_LSTM_AUTO ) h->lstm_mode = big ? SILERO_B200_LSTM_TENSOR : SILERO_B200_LSTM_FP32;
      if ( h->layer_mode == SILERO_B200_LAYERS_AUTO ) h->layer_mode = big ? SILERO_B200_LAYERS_TENSOR : SILERO_B200_LAYERS_FP32;
      if ( h->stft_auto && h->lstm_mode != SILERO_B200_LSTM_TENSOR ) h->stft_mode = SILERO_B200_STFT_EXACT;
      h->stft_auto = 0;
   }

#define CU_H( call )                                                                                   \
   do                                                                                                  \
   {                                                                                                   \
      cudaError_t e_ = ( call );                                                                       \
      if ( e_ != cudaSuccess )                                                                         \
      {                                                                                                \
         set_err( SILERO_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString( e_ ), __FILE__, __LINE__ ); \
         vb_testtensor_free( &tf );                                                                    \
         silero_b200_destroy( h );                                                                     \
         return SILERO_B200_ERR_CUDA;                                                                  \
      }                                                                                                \
   } while ( 0 )

   CU_H( cudaSetDevice( h->device ) );
   cudaDeviceProp prop;
   CU_H( cudaGetDeviceProperties( &prop, h->device ) );
   h->sm_count = prop.multiProcessorCount;
   if ( prop.major < 10 )
   {
      set_err( SILERO_B200_ERR_CUDA, "device %d is sm_%d%d; this build targets sm_100a (B200)", h->device, prop.major, prop.minor );
      vb_testtensor_free( &tf );
      silero_b200_destroy( h );
      return SILERO_B200_ERR_CUDA;
   }
   if ( configure_kernels() )
   {
      vb_testtensor_free( &tf );
      silero_b200_destroy( h );
      return SILERO_B200_ERR_CUDA;
   }
   CU_H( cudaStreamCreateWithFlags( &h->stream, cudaStreamNonBlocking ) );
   CU_H( cudaStreamCreateWithFlags( &h->copy_stream, cudaStreamNonBlocking ) );
   CU_H( cudaStreamCreateWithFlags( &h->front_stream, cudaStreamNonBlocking ) );
   CU_H( cudaEventCreateWithFlags( &h->ev_spec_ready, cudaEventDisableTiming ) );
   CU_H( cudaEventCreateWithFlags( &h->ev_spec_free, cudaEventDisableTiming ) );
   CU_H( cudaEventCreate( &h->ev_begin ) );
   CU_H( cudaEventCreate( &h->ev_end ) );
   for ( int i = 0; i < N_STAGE_EVENTS; ++i ) CU_H( cudaEventCreate( &h->ev_stage[i] ) );
   for ( int i = 0; i < 2; ++i )
   {
      CU_H( cudaEventCreateWithFlags( &h->pcm_ready[i], cudaEventDisableTiming ) );
      CU_H( cudaEventCreateWithFlags( &h->pcm_free[i], cudaEventDisableTiming ) );
   }
   for ( int i = 0; i < 4; ++i ) CU_H( cudaEventCreateWithFlags( &h->done_ev[i], cudaEventDisableTiming ) );

   // pack every weight into one host blob, upload once
   const size_t n_basis = 2 * STFT_BS_FLOATS;
   const size_t n_l0 = LayerPack<0>::TOTAL, n_l1 = LayerPack<1>::TOTAL, n_l2 = LayerPack<2>::TOTAL, n_l3 = LayerPack<3>::TOTAL;
   const size_t n_lstm = 2 * LSTM_WS_FLOATS, n_lb = 512, n_dw = 128, n_db = 4;
   const size_t n_raw = 258 * 256;
   size_t n_all = 0; // every tensor of the container as it is, 4-float aligned (faithful_kernel.cuh)
   for ( int i = 0; i < 99; ++i ) n_all += ( (size_t)tf.tensors[i].size + 3 ) & ~(size_t)3;
   for ( int k = 0; k < fq::N_TRANSPOSED; ++k )
   {
      int idx, n_out, n_in;
      fq::transposed_slot( k, &idx, &n_out, &n_in );
      n_all += (size_t)n_out * n_in; // multiples of 4
   }
   const size_t n_bsym = SSYM_BS_FLOATS;
   const size_t total = n_basis + n_l0 + n_l1 + n_l2 + n_l3 + n_lstm + n_lb + n_dw + n_db + n_raw + n_all + n_bsym;
   float *host = (float *)calloc( total, sizeof( float ) );
   if ( !host )
   {
      vb_testtensor_free( &tf );
      silero_b200_destroy( h );
      return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
   }
   size_t off = 0;
   size_t o_basis = off; off += n_basis;
   size_t o_l0 = off; off += n_l0;
   size_t o_l1 = off; off += n_l1;
   size_t o_l2 = off; off += n_l2;
   size_t o_l3 = off; off += n_l3;
   size_t o_lstm = off; off += n_lstm;
   size_t o_lb = off; off += n_lb;
   size_t o_dw = off; off += n_dw;
   size_t o_db = off; off += n_db;
   size_t o_raw = off; off += n_raw;
   size_t o_all = off; off += n_all;
   size_t o_bsym = off; off += n_bsym;
   size_t all_off[99], tt_off[fq::N_TRANSPOSED];
   {
      size_t o = o_all;
      for ( int i = 0; i < 99; ++i )
      {
         all_off[i] = o;
         memcpy( host + o, tf.tensors[i].data, sizeof( float ) * (size_t)tf.tensors[i].size );
         o += ( (size_t)tf.tensors[i].size + 3 ) & ~(size_t)3;
      }
      for ( int k = 0; k < fq::N_TRANSPOSED; ++k )
      {
         int idx, n_out, n_in;
         fq::transposed_slot( k, &idx, &n_out, &n_in );
         tt_off[k] = o;
         const float *src = tf.tensors[idx].data; // [n_out][n_in]
         for ( int r = 0; r < n_out; ++r )
            for ( int c = 0; c < n_in; ++c ) host[o + (size_t)c * n_out + r] = src[(size_t)r * n_in + c];
         o += (size_t)n_out * n_in;
      }
   }
   pack_basis( tf.tensors[0].data, host + o_basis );
   pack_basis_sym( tf.tensors[0].data, host + o_bsym );
   h->stft_sym = basis_is_mirrored( tf.tensors[0].data ) && !getenv( "SILERO_B200_STFT_NO_SYM" );
   pack_layer<0>( tf.tensors + 1, host + o_l0 );
   pack_layer<1>( tf.tensors + 25, host + o_l1 );
   pack_layer<2>( tf.tensors + 49, host + o_l2 );
   pack_layer<3>( tf.tensors + 71, host + o_l3 );
   pack_lstm( tf.tensors[95].data, host + o_lstm );
   unsigned char *tc_img = (unsigned char *)calloc( 2, LTC_W_BYTES );
   if ( tc_img ) pack_lstm_tc( tf.tensors[95].data, tc_img );
   unsigned char *ltc_img[4] = { (unsigned char *)malloc( L0tc::IMG_BYTES ), (unsigned char *)malloc( LtcCfg<1>::IMG_BYTES ), (unsigned char *)malloc( LtcCfg<2>::IMG_BYTES ),
                                 (unsigned char *)malloc( LtcCfg<3>::IMG_BYTES ) };
   const size_t ltc_bytes[4] = { L0tc::IMG_BYTES, LtcCfg<1>::IMG_BYTES, LtcCfg<2>::IMG_BYTES, LtcCfg<3>::IMG_BYTES };
   if ( ltc_img[0] ) pack_layer0_tc( host + o_l0, ltc_img[0] );
   for ( int f = 0; f < 129; ++f )
   {
      const float *dw = host + o_l0 + LayerPack<0>::DW + f * 8;
      h->l0_dw.w[f][0] = make_float4( dw[0], dw[1], dw[2], dw[3] );
      h->l0_dw.w[f][1] = make_float4( dw[4], dw[5], 0.0f, 0.0f );
   }
   if ( ltc_img[1] ) pack_layer_tc<1>( host + o_l1, ltc_img[1] );
   if ( ltc_img[2] ) pack_layer_tc<2>( host + o_l2, ltc_img[2] );
   if ( ltc_img[3] ) pack_layer_tc<3>( host + o_l3, ltc_img[3] );
   memcpy( host + o_lb, tf.tensors[96].data, sizeof( float ) * 512 );
   memcpy( host + o_dw, tf.tensors[97].data, sizeof( float ) * 128 );
   memcpy( host + o_db, tf.tensors[98].data, sizeof( float ) * 2 );
   memcpy( host + o_raw, tf.tensors[0].data, sizeof( float ) * n_raw );
   vb_testtensor_free( &tf );
   memset( &tf, 0, sizeof( tf ) );

   cudaError_t ce = cudaMalloc( &h->d_weights, total * sizeof( float ) );
   if ( ce == cudaSuccess ) ce = cudaMemcpy( h->d_weights, host, total * sizeof( float ), cudaMemcpyHostToDevice );
   free( host );
   if ( ce == cudaSuccess && !tc_img ) ce = cudaErrorMemoryAllocation;
   if ( ce == cudaSuccess ) ce = cudaMalloc( &h->d_lstm_tc, 2 * LTC_W_BYTES );
   if ( ce == cudaSuccess ) ce = cudaMemcpy( h->d_lstm_tc, tc_img, 2 * LTC_W_BYTES, cudaMemcpyHostToDevice );
   free( tc_img );
   for ( int l = 0; l < 4; ++l )
   {
      if ( ce == cudaSuccess && !ltc_img[l] ) ce = cudaErrorMemoryAllocation;
      if ( ce == cudaSuccess ) ce = cudaMalloc( &h->d_layer_tc[l], ltc_bytes[l] );
      if ( ce == cudaSuccess ) ce = cudaMemcpy( h->d_layer_tc[l], ltc_img[l], ltc_bytes[l], cudaMemcpyHostToDevice );
      free( ltc_img[l] );
   }
   CU_H( ce );
   h->w.basis_pack = h->d_weights + o_basis;
   h->w.layer[0] = h->d_weights + o_l0;
   h->w.layer[1] = h->d_weights + o_l1;
   h->w.layer[2] = h->d_weights + o_l2;
   h->w.layer[3] = h->d_weights + o_l3;
   h->w.lstm_w = h->d_weights + o_lstm;
   h->w.lstm_b = h->d_weights + o_lb;
   h->w.dec_w = h->d_weights + o_dw;
   h->w.dec_b = h->d_weights + o_db;
   h->w.basis_raw = h->d_weights + o_raw;
   h->basis_sym = h->d_weights + o_bsym;
   for ( int i = 0; i < 99; ++i ) h->fw.t[i] = h->d_weights + all_off[i];
   for ( int k = 0; k < fq::N_TRANSPOSED; ++k ) h->fw.tt[k] = h->d_weights + tt_off[k];
   CU_H( cudaMalloc( &h->d_flagged, sizeof( unsigned long long ) ) );
   CU_H( cudaMalloc( &h->d_lstm_sync, ( (size_t)h->max_streams + 1 ) * sizeof( int ) ) );
   CU_H( cudaHostAlloc( (void **)&h->err_word, sizeof( int ), cudaHostAllocMapped ) );
   *h->err_word = 0;
   CU_H( cudaHostGetDevicePointer( (void **)&h->d_err_word, h->err_word, 0 ) );
   h->wave_spin_limit = 1 << 24;
   h->token_min_chunks = SILERO_B200_EXACT_TOKEN_MIN_CHUNKS;
   if ( const char *e = getenv( "SILERO_B200_TOKEN_MIN_CHUNKS" ) )
      if ( atoi( e ) > 0 ) h->token_min_chunks = atoi( e ); // (both mappings give the same bits: only speed depends on it)
   {
      size_t free_b = 0, total_b = 0;
      h->scratch_budget = (size_t)1536 << 20;
      if ( cudaMemGetInfo( &free_b, &total_b ) == cudaSuccess )
      {
         size_t b = free_b / 10;
         if ( b < ( (size_t)1536 << 20 ) ) b = (size_t)1536 << 20;
         if ( b > ( (size_t)12 << 30 ) ) b = (size_t)12 << 30;
         h->scratch_budget = b;
      }
   }
   CU_H( cudaMemset( h->d_flagged, 0, sizeof( unsigned long long ) ) );

   size_t sbytes = (size_t)h->max_streams * SILERO_B200_STATE_FLOATS * sizeof( float );
   CU_H( cudaMalloc( &h->state_h, sbytes ) );
   CU_H( cudaMalloc( &h->state_c, sbytes ) );
   CU_H( cudaMemset( h->state_h, 0, sbytes ) );
   CU_H( cudaMemset( h->state_c, 0, sbytes ) );
   CU_H( cudaMalloc( &h->d_seg_state, (size_t)h->max_streams * sizeof( SegStateDev ) ) );
   CU_H( cudaMemset( h->d_seg_state, 0, (size_t)h->max_streams * sizeof( SegStateDev ) ) );
   set_seg_params( h, 0 );
#undef CU_H
   *out = h;
   return SILERO_B200_OK;
}

extern "C" int silero_b200_create( const void *bytes, size_t nbytes, const silero_b200_opts *opts, silero_b200 **out )
{
   if ( !bytes || !out ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   *out = 0;
   return create_impl( bytes, nbytes, opts, out );
}

extern "C" int silero_b200_create_from_file( const char *path, const silero_b200_opts *opts, silero_b200 **out )
{
   if ( !path || !out ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   *out = 0;
   FILE *f = fopen( path, "rb" );
   if ( !f ) return set_err( SILERO_B200_ERR_WEIGHTS, "cannot open %s", path );
   fseek( f, 0, SEEK_END );
   long n = ftell( f );
   fseek( f, 0, SEEK_SET );
   void *b = malloc( n > 0 ? (size_t)n : 1 );
   size_t got = b ? fread( b, 1, (size_t)n, f ) : 0;
   fclose( f );
   int rc = ( b && got == (size_t)n ) ? create_impl( b, (size_t)n, opts, out ) : set_err( SILERO_B200_ERR_WEIGHTS, "cannot read %s", path );
   free( b );
   return rc;
}

extern "C" int silero_b200_get_info( const silero_b200 *h, silero_b200_info *info )
{
   if ( !h || !info ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   info->batch_size_restriction = -1; // silero.h:39
   info->is_silero_v5 = 0;            // silero.h:40
   info->input_size_min = 1536;       // silero.h:41
   info->input_size_max = 1536;       // silero.h:42
   info->output_dims = 3;             // silero.h:43
   info->sm_count = h->sm_count;
   info->max_streams = h->max_streams;
   info->window_chunks = h->window_chunks_opt;
   return SILERO_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// scratch
// ---------------------------------------------------------------------------------------------
template <typename T>
static int grow( T **p, size_t *cap, size_t need )
{
   if ( need <= *cap ) return 0;
   if ( *p ) CU( cudaFree( *p ) );
   *p = 0;
   *cap = 0;
   CU( cudaMalloc( p, need * sizeof( T ) ) );
   *cap = need;
   return 0;
}

static int ensure_scratch( silero_b200 *h, size_t chunks, size_t h0_floats )
{
   if ( h0_floats > h->cap_h0_floats )
   {
      CU( cudaStreamSynchronize( h->stream ) );
      cudaFree( h->h0 );
      h->h0 = 0;
      h->cap_h0_floats = 0;
      CU( cudaMalloc( &h->h0, h0_floats * sizeof( float ) ) );
      h->cap_h0_floats = h0_floats;
   }
   if ( chunks <= h->cap_chunks ) return 0;
   CU( cudaStreamSynchronize( h->stream ) );
   cudaFree( h->spec ); cudaFree( h->a1 ); cudaFree( h->a2 ); cudaFree( h->a3 ); cudaFree( h->a4 ); cudaFree( h->mu ); cudaFree( h->y1 );
   h->spec = h->a1 = h->a2 = h->a3 = h->a4 = h->mu = h->y1 = 0;
   h->cap_chunks = 0;
   CU( cudaMalloc( &h->spec, chunks * VB_BINS * VB_FRAMES * sizeof( float ) ) );
   CU( cudaMalloc( &h->a1, chunks * 13 * 16 * sizeof( float ) ) );
   CU( cudaMalloc( &h->a2, chunks * 7 * 32 * sizeof( float ) ) );
   CU( cudaMalloc( &h->a3, chunks * 7 * 32 * sizeof( float ) ) );
   CU( cudaMalloc( &h->a4, chunks * 7 * 64 * sizeof( float ) ) );
   CU( cudaMalloc( &h->mu, chunks * sizeof( float ) ) );
   CU( cudaMalloc( &h->y1, chunks * 16 * VB_FRAMES * sizeof( float ) ) );
   h->cap_chunks = chunks;
   return 0;
}

// bytes of window scratch per chunk: spec + a1..a4 + h0
static const size_t kScratchPerChunk = ( VB_BINS * VB_FRAMES + 16 * VB_FRAMES + 13 * 16 + 7 * 32 * 2 + 7 * 64 * 2 ) * sizeof( float );

// host_path: windows are also the granularity at which the H2D copy of the next window overlaps compute, so calls that start from
// host memory keep the small (1.5 GB) windows; device-resident input takes the large ones
static int pick_window( const silero_b200 *h, int nstreams, int nchunks, bool host_path = false )
{
   if ( h->window_chunks_opt > 0 ) return h->window_chunks_opt < nchunks ? h->window_chunks_opt : nchunks;
   // Window scratch: a tenth of the memory that was free at creation, between 1.5 and 12 GB. Every window boundary drains the GPU
   // seven times (persistent kernels with static tile assignment have a tail) and reloads weights into shared / tensor memory:
   // measured on 4096 streams x 125 chunks, 18-chunk windows (1.5 GB) 19.6 ms per step, 63-chunk windows 18.9 ms, one window 18.7 ms.
   const size_t budget = ( h->scratch_budget && !host_path ) ? h->scratch_budget : ( (size_t)1536 << 20 );
   long long per_stream = (long long)( budget / kScratchPerChunk ) / ( nstreams > 0 ? nstreams : 1 );
   if ( per_stream < 1 ) per_stream = 1;
   if ( per_stream > nchunks ) per_stream = nchunks;
   // equal windows: 125 chunks under a cap of 20 run as 7 x 18 (17) instead of 6 x 20 + 5 -- a 5-chunk window pays the same
   // per-launch costs (weights to shared / tensor memory, partial waves) for a quarter of the work
   const long long nwin = ( nchunks + per_stream - 1 ) / per_stream;
   per_stream = ( nchunks + nwin - 1 ) / nwin;
   return (int)per_stream;
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
static inline int imin( int a, int b ) { return a < b ? a : b; }
static inline int imax( int a, int b ) { return a > b ? a : b; }

// mu (optional): receives the adaptive-normalization scalar per chunk when the selected kernel produces it (the FFT
// kernel); the tensor-core and exact kernels leave it to the first layer (the caller checks stft_produces_mu)
static bool stft_produces_mu( const silero_b200 *h, int nchunks )
{
   (void)nchunks;
   return h->stft_mode != SILERO_B200_STFT_EXACT;
}

static int launch_stft( silero_b200 *h, const void *d_in, int in_f32, long long stream_stride, int nw, int nchunks, float *spec, int out_mode, float *mu = 0,
                        cudaStream_t st = 0 )
{
   if ( !st ) st = h->stream;
   if ( h->stft_mode == SILERO_B200_STFT_EXACT && h->stft_sym )
   {
      const int grid = imin( h->sm_count, nchunks );
      if ( in_f32 )
         stft_sym_kernel<true><<<grid, SSYM_THREADS, SSYM_SMEM_BYTES, st>>>( d_in, stream_stride, nw, nchunks, h->basis_sym, spec, out_mode );
      else
         stft_sym_kernel<false><<<grid, SSYM_THREADS, SSYM_SMEM_BYTES, st>>>( d_in, stream_stride, nw, nchunks, h->basis_sym, spec, out_mode );
   }
   else if ( h->stft_mode == SILERO_B200_STFT_EXACT )
   {
      int npairs = imin( h->sm_count / 2, ( nchunks + STFT_GROUPS - 1 ) / STFT_GROUPS );
      if ( npairs < 1 ) npairs = 1;
      if ( in_f32 )
         stft_logmag_kernel<true><<<npairs * 2, STFT_THREADS, STFT_SMEM_BYTES, st>>>( d_in, stream_stride, nw, nchunks, h->w.basis_pack, spec, out_mode );
      else
         stft_logmag_kernel<false><<<npairs * 2, STFT_THREADS, STFT_SMEM_BYTES, st>>>( d_in, stream_stride, nw, nchunks, h->w.basis_pack, spec, out_mode );
   }
   else if ( h->stft_mode == 0 )
   {
      static int per_sm8 = 0;
      if ( !per_sm8 )
      {
         CU( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm8, stft_fft8_kernel<false, false>, F8_THREADS, F8_SMEM_BYTES ) );
         if ( per_sm8 < 1 ) per_sm8 = 1;
      }
      int grid = imin( nchunks, h->sm_count * per_sm8 );
      if ( in_f32 )
      {
         if ( out_mode )
            stft_fft8_kernel<true, true><<<grid, F8_THREADS, F8_SMEM_BYTES, st>>>( d_in, stream_stride, nw, nchunks, h->w.basis_raw, spec, mu, h->stft_k_rel, h->d_flagged );
         else
            stft_fft8_kernel<true, false><<<grid, F8_THREADS, F8_SMEM_BYTES, st>>>( d_in, stream_stride, nw, nchunks, h->w.basis_raw, spec, mu, h->stft_k_rel, h->d_flagged );
      }
      else
      {
         if ( out_mode )
            stft_fft8_kernel<false, true><<<grid, F8_THREADS, F8_SMEM_BYTES, st>>>( d_in, stream_stride, nw, nchunks, h->w.basis_raw, spec, mu, h->stft_k_rel, h->d_flagged );
         else
            stft_fft8_kernel<false, false><<<grid, F8_THREADS, F8_SMEM_BYTES, st>>>( d_in, stream_stride, nw, nchunks, h->w.basis_raw, spec, mu, h->stft_k_rel, h->d_flagged );
      }
      h->bins_total += (unsigned long long)nchunks * VB_BINS * VB_FRAMES;
   }
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

template <int L, bool NORM>
static int launch_layer( silero_b200 *h, const float *in, float *out, int nchunks, int entry = ENTRY_LAYER, int tap = TAP_LAYER, const float *mu = 0 )
{
   using Cfg = LayerCfg<L>;
   int ntiles = ( nchunks + Cfg::G - 1 ) / Cfg::G;
   int per_sm = ( 227 * 1024 ) / ( Cfg::SMEM_BYTES + 1024 );
   if ( per_sm < 1 ) per_sm = 1;
   if ( per_sm > 8 ) per_sm = 8;
   int grid = imin( ntiles, h->sm_count * per_sm );
   layer_kernel<L, NORM><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, h->stream>>>( in, out, h->w.layer[L], nchunks, entry, tap, mu );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

// tensor-core layer (layer_tc_kernel.cuh), layers 2..4 (L = 1..3); same in/out layouts as launch_layer
template <int L>
static int launch_layer_tc( silero_b200 *h, const float *in, float *out, int nchunks )
{
   using Cfg = LtcCfg<L>;
   const int ntiles = ( nchunks + Cfg::CPT - 1 ) / Cfg::CPT;
   const int grid = imin( ( ntiles + Cfg::NGROUPS - 1 ) / Cfg::NGROUPS, h->sm_count );
   layer_tc_kernel<L><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, h->stream>>>( in, out, h->d_layer_tc[L], nchunks );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

// first layer on the tensor cores; in: log spectrogram [chunk][129][25], mu: per-chunk normalization scalar (NULL: input already normalized)
static int launch_layer0_tc( silero_b200 *h, const float *in, float *out, int nchunks, const float *mu, int compute_mu = 0 )
{
   const int ntiles = ( nchunks + 3 ) / 4;
   const int grid = imin( ( ntiles + L0tc::NGROUPS - 1 ) / L0tc::NGROUPS, h->sm_count );
   if ( compute_mu )
      layer0_tc_kernel<true><<<grid, L0tc::THREADS, L0tc::SMEM_BYTES, h->stream>>>( in, out, h->d_layer_tc[0], nchunks, mu, h->l0_dw );
   else
      layer0_tc_kernel<false><<<grid, L0tc::THREADS, L0tc::SMEM_BYTES, h->stream>>>( in, out, h->d_layer_tc[0], nchunks, mu, h->l0_dw );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

static bool layers_use_tensor( const silero_b200 *h, int nchunks )
{
   (void)nchunks; // decided once per engine (create_impl), not per call
   return h->layer_mode == SILERO_B200_LAYERS_TENSOR;
}

// layers 2..4 of the encoder on whichever kernel family the engine is configured for
template <int L>
static int launch_layer_any( silero_b200 *h, const float *in, float *out, int nchunks )
{
   return layers_use_tensor( h, nchunks ) ? launch_layer_tc<L>( h, in, out, nchunks ) : launch_layer<L, false>( h, in, out, nchunks );
}

template <int LAYER>
static int launch_lstm( silero_b200 *h, const float *x, float *hseq, int first_stream, int nstreams, int nw, float *d_out2, float *d_probs,
                        long long out_stride, long long out_off )
{
   float *sh = h->state_h + (size_t)first_stream * SILERO_B200_STATE_FLOATS;
   float *sc = h->state_c + (size_t)first_stream * SILERO_B200_STATE_FLOATS;
   const bool wide = nstreams >= 4 * h->sm_count / 2;
   const int st = wide ? 4 : 1;
   int tiles = ( nstreams + st - 1 ) / st;
   int ng = ( tiles + h->sm_count - 1 ) / h->sm_count;
   if ( ng > LSTM_MAX_GROUPS ) ng = LSTM_MAX_GROUPS;
   if ( ng < 1 ) ng = 1;
   int grid = imin( ( tiles + ng - 1 ) / ng, h->sm_count );
   if ( wide )
      lstm_layer_kernel<LAYER, 4><<<grid, 64 * ng, LstmSmem<4>::BYTES, h->stream>>>( x, hseq, sh, sc, h->w.lstm_w, h->w.lstm_b, h->w.dec_w, h->w.dec_b, nstreams,
                                                                                     nw, d_out2, d_probs, out_stride, out_off );
   else
      lstm_layer_kernel<LAYER, 1><<<grid, 64 * ng, LstmSmem<1>::BYTES, h->stream>>>( x, hseq, sh, sc, h->w.lstm_w, h->w.lstm_b, h->w.dec_w, h->w.dec_b, nstreams,
                                                                                     nw, d_out2, d_probs, out_stride, out_off );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

static bool lstm_use_tensor( const silero_b200 *h, int nstreams )
{
   (void)nstreams; // decided once per engine (create_impl), not per call
   return h->lstm_mode == SILERO_B200_LSTM_TENSOR;
}

// tensor-core LSTM (lstm_tc_kernel.cuh): layer 0 consumes a4 and leaves its packed h sequence in h->h0
template <int LAYER>
static int launch_lstm_tc( silero_b200 *h, const float *a4, int first_stream, int nstreams, int nw, float *d_out2, float *d_probs, long long out_stride,
                           long long out_off )
{
   float *sh = h->state_h + (size_t)first_stream * SILERO_B200_STATE_FLOATS;
   float *sc = h->state_c + (size_t)first_stream * SILERO_B200_STATE_FLOATS;
   const int ntiles = ( nstreams + LTC_N - 1 ) / LTC_N;
   const int grid = imin( ntiles, h->sm_count );
   unsigned char *hp = reinterpret_cast<unsigned char *>( h->h0 );
   lstm_tc_kernel<LAYER><<<grid, LTC_THREADS, LTC_SMEM_BYTES, h->stream>>>( a4, hp, hp, sh, sc, h->d_lstm_tc, h->w.lstm_b, h->w.dec_w, h->w.dec_b, nstreams, nw,
                                                                            d_out2, d_probs, out_stride, out_off );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

// the whole path in the reference's rounding sequence (faithful_kernel.cuh)
static bool use_faithful( const silero_b200 *h, int nstreams )
{
   (void)nstreams; // decided once per engine (create_impl), not per call
   return h->faithful != 0;
}

static int launch_faithful_encoder( silero_b200 *h, const float *spec, float *a4, int nchunks )
{
   const int grid = imin( nchunks, h->sm_count * 4 );
   faithful_encoder_kernel<<<grid, FAITHFUL_THREADS, FAITHFUL_SMEM_BYTES, h->stream>>>( spec, a4, h->fw, nchunks );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

// the encoder in the reference's rounding sequence, thread = token (exact_encoder_kernel.cuh): front (normalization scalar, depthwise
// conv, the two K = 129 contractions) + one launch per layer
template <int L>
static int launch_exact_layer( silero_b200 *h, const float *in, float *out, int nchunks )
{
   using Cfg = XeCfg<L>;
   // chunks per batch: a full CTA's worth (GB) when there is work for every SM, fewer when the window is small -- a window of 4 736
   // chunks is 65 full batches of layer 4 (65 of 148 SMs busy) or 148 batches of 32 chunks
   const int gb = imax( 1, imin( Cfg::GB, ( nchunks + h->sm_count - 1 ) / h->sm_count ) );
   const int grid = imin( ( nchunks + gb - 1 ) / gb, h->sm_count );
   exact_layer_kernel<L><<<grid, XE_THREADS, Cfg::SMEM_BYTES, h->stream>>>( in, out, h->fw.t[L == 0 ? 1 : ( L == 1 ? 25 : ( L == 2 ? 49 : 71 ) )], h->xe_scratch, nchunks, gb );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

static int launch_exact_front( silero_b200 *h, const float *spec, float *y1, int nchunks, bool normalize = true )
{
   if ( !h->xe_scratch ) CU( cudaMalloc( &h->xe_scratch, XeCfg<3>::SCRATCH_FLOATS * sizeof( float ) * (size_t)h->sm_count ) );
   const int grid = imin( ( nchunks + XF_G - 1 ) / XF_G, h->sm_count );
   if ( normalize )
      exact_front_kernel<true><<<grid, XF_THREADS, XF_SMEM_BYTES, h->stream>>>( spec, y1, h->fw.t[1], nchunks );
   else
      exact_front_kernel<false><<<grid, XF_THREADS, XF_SMEM_BYTES, h->stream>>>( spec, y1, h->fw.t[1], nchunks );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

static int launch_exact_encoder( silero_b200 *h, const float *spec, float *a4, int nchunks, bool normalize = true )
{
   if ( launch_exact_front( h, spec, h->y1, nchunks, normalize ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_exact_layer<0>( h, h->y1, h->a1, nchunks ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_exact_layer<1>( h, h->a1, h->a2, nchunks ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_exact_layer<2>( h, h->a2, h->a3, nchunks ) ) return SILERO_B200_ERR_CUDA;
   return launch_exact_layer<3>( h, h->a3, a4, nchunks );
}

// the decoder LSTM in the reference's rounding sequence (exact_lstm_kernel.cuh: weights in registers, a CTA walks a set of streams
// together). x0: encoder output [S][nw*7][64] on entry, top layer's output sequence on exit; h0: the first layer's sequence.
// Few streams (2 x groups <= SMs): both layers in one wavefront launch; else one launch per layer.
#define XL_WAVE_MAX_PER_CTA 10 // measured: the wavefront wins up to ~10 streams per CTA (297 streams: 1.09 vs 1.59 ms), two launches win from ~1000 streams on (4096: 9.4 vs 9.9 ms)
static void stage_mark( silero_b200 *h, int i );
static int launch_lstm_exact( silero_b200 *h, float *x0, float *h0, int first_stream, int nstreams, int nw, int force_mode = 0, bool mark_layers = false )
{
   float *sh = h->state_h + (size_t)first_stream * SILERO_B200_STATE_FLOATS;
   float *sc = h->state_c + (size_t)first_stream * SILERO_B200_STATE_FLOATS;
   const int half_sms = h->sm_count / 2;
   const bool wave = force_mode ? force_mode == 1 : nstreams <= half_sms * XL_WAVE_MAX_PER_CTA;
   if ( wave )
   {
      const int groups = imin( nstreams, half_sms );
      CU( cudaMemsetAsync( h->d_lstm_sync, 0, ( (size_t)groups + 1 ) * sizeof( int ), h->stream ) );
      exact_lstm_kernel<true><<<2 * groups, XL_THREADS, XL_SMEM_BYTES, h->stream>>>( x0, h0, x0, sh, sc, h->w.lstm_w, h->w.lstm_b, nstreams, nw, 0, h->d_lstm_sync,
                                                                                     h->d_err_word, h->wave_spin_limit, h->debug_stall_producer );
      h->launches++;
   }
   else
   {
      const int grid = imin( nstreams, h->sm_count );
      for ( int layer = 0; layer < 2; ++layer )
      {
         exact_lstm_kernel<false><<<grid, XL_THREADS, XL_SMEM_BYTES, h->stream>>>( x0, h0, x0, sh, sc, h->w.lstm_w, h->w.lstm_b, nstreams, nw, layer, 0, h->d_err_word,
                                                                                   0, 0 );
         h->launches++;
         if ( mark_layers && layer == 0 ) stage_mark( h, 6 ); // stage [6] = layer 0, [7] = layer 1 + decoder head
      }
   }
   if ( mark_layers && wave ) stage_mark( h, 6 );             // one launch: stage [6] = both layers, [7] = decoder head
   CU( cudaGetLastError() );
   return 0;
}

// first encoder layer from the log spectrogram: mu = per-chunk normalization scalar if the STFT kernel produced it, else NULL
// (the layer computes it itself, misc.c:48-121)
static int first_layer_from_logspec( silero_b200 *h, const float *spec, float *a1, int nchunks, const float *mu )
{
   if ( layers_use_tensor( h, nchunks ) ) return launch_layer0_tc( h, spec, a1, nchunks, mu, mu ? 0 : 1 );
   return mu ? launch_layer<0, false>( h, spec, a1, nchunks, ENTRY_LAYER, TAP_LAYER, mu ) : launch_layer<0, true>( h, spec, a1, nchunks );
}

// Stage boundaries of a window: a CUDA event when per-stage profiling is on, and always an NVTX range (header-only nvtx3: a no-op
// unless a profiler is attached) named after the reference function(s) the stage replaces -- the counterpart of the reference's Tracy
// zones (TracyCZone in stft.c, conv.c, transformer.c, lstm.c, silero_v3.c).
static const char *const kStageNames[8] = { "window", "my_stft+log1p (stft.c:15, misc.c:40)", "transformer_layer 1 (transformer.c:237)", "transformer_layer 2",
                                            "transformer_layer 3", "transformer_layer 4", "lstm layer 0 (lstm.c:228)", "lstm layer 1 + decoder (silero_v3.c:231)" };
static void stage_mark( silero_b200 *h, int i )
{
   if ( h->profiling ) cudaEventRecord( h->ev_stage[i], h->stream );
   if ( i > 0 ) nvtxRangePop();
   if ( i < 7 ) nvtxRangePushA( kStageNames[i + 1] );
}

// one window: nstreams x nw chunks; input chunk (s, n) at d_in + s*stream_stride + n*1536
// in_ready (optional): event the input samples become valid on; in_free (optional): recorded once the input has been consumed.
//
// The STFT runs on front_stream: the spectrogram (+ normalization scalars) is the only thing it shares with the rest of the window,
// and it is free again as soon as the first layer of the previous window has read it. Consecutive windows therefore overlap: the
// STFT of window w+1 (CUDA-core kernel, 66 KB of shared memory and 80 registers per thread) runs beside layers 2..4 and above all
// beside the two LSTM kernels of window w, which leave 20 SMs empty and the issue slots of the others half idle.
// Per-stage profiling keeps everything on one stream, so that the stage times stay those of the kernels alone.
static int run_window( silero_b200 *h, const void *d_in, int in_f32, long long stream_stride, int first_stream, int nstreams, int nw, float *d_out2,
                       float *d_probs, long long out_stride, long long out_off, int accumulate_timing, cudaEvent_t in_ready = 0, cudaEvent_t in_free = 0,
                       bool input_is_ordered_on_stream = false )
{
   const int nchunks = nstreams * nw;
   const size_t h0_floats = (size_t)( ( nstreams + LTC_N - 1 ) / LTC_N ) * LTC_N * nw * 7 * 64;
   if ( ensure_scratch( h, (size_t)nchunks, h0_floats ) ) return SILERO_B200_ERR_CUDA;
   // (which kernels run is a property of the engine, fixed at creation: create_impl)
   const bool faithful = use_faithful( h, nstreams );
   stage_mark( h, 0 );
   const bool have_mu = stft_produces_mu( h, nchunks );
   // (an input that was produced by earlier work on h->stream itself, like run_chunks' own upload, keeps the STFT on that stream)
   const bool overlap = !h->profiling && !input_is_ordered_on_stream;
   cudaStream_t fs = overlap ? h->front_stream : h->stream;
   if ( overlap )
   {
      // first layer of the previous window -- of this call or of the previous one: back-to-back calls overlap the same way, so a
      // call's own ev_begin..ev_end time can miss its first STFT; timer_start/timer_stop around several calls see all of it
      // (a never-recorded event counts as complete)
      CU( cudaStreamWaitEvent( fs, h->ev_spec_free, 0 ) );
   }
   if ( in_ready ) CU( cudaStreamWaitEvent( fs, in_ready, 0 ) );
   if ( launch_stft( h, d_in, in_f32, stream_stride, nw, nchunks, h->spec, 0, have_mu ? h->mu : 0, fs ) ) return SILERO_B200_ERR_CUDA;
   if ( in_free ) CU( cudaEventRecord( in_free, fs ) );
   if ( overlap )
   {
      CU( cudaEventRecord( h->ev_spec_ready, fs ) );
      CU( cudaStreamWaitEvent( h->stream, h->ev_spec_ready, 0 ) );
   }
   stage_mark( h, 1 );
   if ( faithful )
   {
      // log spectrogram (bit-identical to the reference's) -> encoder -> LSTM -> decoder, every step in the reference's rounding
      // sequence. Two mappings of the same arithmetic per stage, chosen per window by its shape (identical bits, so the choice is
      // free): the encoder as a CTA per chunk for small windows (more parallelism: latency) or a thread per token from
      // SILERO_B200_EXACT_TOKEN_MIN_CHUNKS chunks up (no partial warps, one barrier per layer: throughput); the LSTM as a wavefront
      // of (stream, layer) CTAs while there are SM pairs for every stream, else CTAs that walk sets of streams together.
      // Stage times (profiling): [2] front + rest of layer 1, [3..5] layers 2..4, [6] LSTM layer 0, [7] LSTM layer 1 + decoder head
      // (small windows: [2] the whole encoder, [6] both LSTM layers).
      if ( nchunks < h->token_min_chunks )
      {
         if ( launch_faithful_encoder( h, h->spec, h->a4, nchunks ) ) return SILERO_B200_ERR_CUDA;
         CU( cudaEventRecord( h->ev_spec_free, h->stream ) );
         stage_mark( h, 2 );
         stage_mark( h, 3 );
         stage_mark( h, 4 );
         stage_mark( h, 5 );
      }
      else
      {
         if ( launch_exact_front( h, h->spec, h->y1, nchunks ) ) return SILERO_B200_ERR_CUDA;
         CU( cudaEventRecord( h->ev_spec_free, h->stream ) );
         if ( launch_exact_layer<0>( h, h->y1, h->a1, nchunks ) ) return SILERO_B200_ERR_CUDA;
         stage_mark( h, 2 );
         if ( launch_exact_layer<1>( h, h->a1, h->a2, nchunks ) ) return SILERO_B200_ERR_CUDA;
         stage_mark( h, 3 );
         if ( launch_exact_layer<2>( h, h->a2, h->a3, nchunks ) ) return SILERO_B200_ERR_CUDA;
         stage_mark( h, 4 );
         if ( launch_exact_layer<3>( h, h->a3, h->a4, nchunks ) ) return SILERO_B200_ERR_CUDA;
         stage_mark( h, 5 );
      }
      if ( launch_lstm_exact( h, h->a4, h->h0, first_stream, nstreams, nw, 0, true ) ) return SILERO_B200_ERR_CUDA;
      {
         const long long n = (long long)nchunks * 2;
         faithful_decoder_kernel<<<(unsigned)( ( n + 127 ) / 128 ), 128, 0, h->stream>>>( h->a4, h->w.dec_w, h->w.dec_b, nstreams, nw, d_out2, d_probs, out_stride, out_off );
         h->launches++;
         CU( cudaGetLastError() );
      }
      stage_mark( h, 7 );
   }
   else
   {
   // the FFT STFT kernel also produces the normalization scalar; the others leave it to the first layer
   if ( first_layer_from_logspec( h, h->spec, h->a1, nchunks, have_mu ? h->mu : 0 ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaEventRecord( h->ev_spec_free, h->stream ) );
   stage_mark( h, 2 );
   if ( launch_layer_any<1>( h, h->a1, h->a2, nchunks ) ) return SILERO_B200_ERR_CUDA;
   stage_mark( h, 3 );
   if ( launch_layer_any<2>( h, h->a2, h->a3, nchunks ) ) return SILERO_B200_ERR_CUDA;
   stage_mark( h, 4 );
   if ( launch_layer_any<3>( h, h->a3, h->a4, nchunks ) ) return SILERO_B200_ERR_CUDA;
   stage_mark( h, 5 );
   const bool tensor = lstm_use_tensor( h, nstreams );
   if ( tensor ? launch_lstm_tc<0>( h, h->a4, first_stream, nstreams, nw, 0, 0, 0, 0 ) : launch_lstm<0>( h, h->a4, h->h0, first_stream, nstreams, nw, 0, 0, 0, 0 ) )
      return SILERO_B200_ERR_CUDA;
   stage_mark( h, 6 );
   if ( tensor ? launch_lstm_tc<1>( h, 0, first_stream, nstreams, nw, d_out2, d_probs, out_stride, out_off )
               : launch_lstm<1>( h, h->h0, 0, first_stream, nstreams, nw, d_out2, d_probs, out_stride, out_off ) )
      return SILERO_B200_ERR_CUDA;
   stage_mark( h, 7 );
   }
   if ( h->profiling && accumulate_timing )
   {
      // per-window stage times are accumulated on the host after a sync (profiling mode only)
      CU( cudaEventSynchronize( h->ev_stage[7] ) );
      for ( int i = 0; i < 7; ++i )
      {
         float ms = 0.0f;
         cudaEventElapsedTime( &ms, h->ev_stage[i], h->ev_stage[i + 1] );
         h->stage_ms[1 + i] += ms;
      }
   }
   return 0;
}

static int check_streams( const silero_b200 *h, int first_stream, int nstreams, int nchunks )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null handle" );
   if ( nstreams < 0 || nchunks < 0 || first_stream < 0 || first_stream + nstreams > h->max_streams )
      return set_err( SILERO_B200_ERR_ARG, "streams [%d,%d) outside [0,%d) or negative count", first_stream, first_stream + nstreams, h->max_streams );
   if ( (long long)nstreams * nchunks > 2000000000ll ) return set_err( SILERO_B200_ERR_ARG, "too many chunks in one call" );
   return 0;
}

static void timing_begin( silero_b200 *h )
{
   memset( h->stage_ms, 0, sizeof( h->stage_ms ) );
   h->launches = 0;
   h->timing_valid = 0;
   cudaEventRecord( h->ev_begin, h->stream );
}

static void timing_end( silero_b200 *h )
{
   cudaEventRecord( h->ev_end, h->stream );
   h->timing_valid = 1;
}

extern "C" int silero_b200_run_streams_device( silero_b200 *h, const int16_t *d_pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                               float *d_probs, float *d_out2 )
{
   int rc = check_streams( h, first_stream, nstreams, nchunks );
   if ( rc ) return rc;
   if ( nstreams == 0 || nchunks == 0 ) return SILERO_B200_OK;
   if ( !d_pcm ) return set_err( SILERO_B200_ERR_ARG, "null pcm" );
   if ( ( stream_stride % 8 ) != 0 || ( (uintptr_t)d_pcm % 16 ) != 0 ) return set_err( SILERO_B200_ERR_ARG, "device pcm must be 16-byte aligned with stream_stride %% 8 == 0" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   timing_begin( h );
   const int nw_max = pick_window( h, nstreams, nchunks );
   for ( int n0 = 0; n0 < nchunks; n0 += nw_max )
   {
      int nw = imin( nw_max, nchunks - n0 );
      rc = run_window( h, d_pcm + (long long)n0 * VB_CHUNK, 0, stream_stride, first_stream, nstreams, nw, d_out2, d_probs, nchunks, n0, 1 );
      if ( rc ) return rc;
   }
   timing_end( h );
   return SILERO_B200_OK;
}

// after a synchronization point: did a kernel give up? (the consumers of exact_lstm_kernel's wavefront raise the word instead of hanging)
static int check_err_word( silero_b200 *h )
{
   const int e = *(volatile int *)h->err_word;
   if ( !e ) return 0;
   *(volatile int *)h->err_word = 0;
   return set_err( SILERO_B200_ERR_CUDA, "LSTM wavefront: the layer-1 task of stream %d lost its layer-0 producer (no progress within the poll limit); "
                                         "the stream's state was left untouched, the call's results are invalid", e - 1 );
}

extern "C" int silero_b200_sync( silero_b200 *h )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null handle" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaStreamSynchronize( h->stream ) );
   return check_err_word( h );
}

extern "C" int silero_b200_debug_wavefront( silero_b200 *h, int stall_producer, int spin_limit )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null handle" );
   h->debug_stall_producer = stall_producer;
   h->wave_spin_limit = spin_limit > 0 ? spin_limit : ( 1 << 24 );
   return SILERO_B200_OK;
}

// ---- on-device segmenter launches -------------------------------------------------------------
static int launch_segments( silero_b200 *h, const float *d_probs, long long stride, long long off, int first_stream, int nstreams, int nchunks, int finish,
                            SegPair *d_segs, int cap, int *d_counts )
{
   const int threads = 128;
   segment_scan_kernel<<<( nstreams + threads - 1 ) / threads, threads, 0, h->stream>>>( d_probs, stride, off, nchunks, h->d_seg_state + first_stream, nstreams,
                                                                                        h->seg_params, finish, d_segs, cap, d_counts );
   h->launches++;
   CU( cudaGetLastError() );
   return 0;
}

struct SegRequest
{
   int finish, cap;
   vadc_segment *segs; // host [nstreams][cap]
   int *counts;        // host [nstreams]
};

static int run_streams_host( silero_b200 *h, const int16_t *pcm, long long stream_stride, int first_stream, int nstreams, int nchunks, float *probs, float *out2,
                             const SegRequest *seg, unsigned long long *ticket = 0 )
{
   int rc = check_streams( h, first_stream, nstreams, nchunks );
   if ( rc ) return rc;
   if ( nstreams == 0 ) return SILERO_B200_OK;
   if ( nchunks == 0 && !( seg && seg->finish ) ) return SILERO_B200_OK;
   if ( nchunks > 0 && !pcm ) return set_err( SILERO_B200_ERR_ARG, "null pcm" );
   if ( seg && ( !seg->segs || !seg->counts || seg->cap < 1 ) ) return set_err( SILERO_B200_ERR_ARG, "segments: null output or cap < 1" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;

   const size_t nout = (size_t)nstreams * nchunks;
   if ( ( probs || seg ) && grow( &h->d_probs, &h->d_probs_cap, nout ? nout : 1 ) ) return SILERO_B200_ERR_CUDA;
   if ( out2 && grow( &h->d_out2, &h->d_out2_cap, nout * 2 ) ) return SILERO_B200_ERR_CUDA;
   if ( seg )
   {
      if ( grow( &h->d_segs, &h->d_segs_cap, (size_t)nstreams * seg->cap ) ) return SILERO_B200_ERR_CUDA;
      if ( grow( &h->d_counts, &h->d_counts_cap, (size_t)nstreams ) ) return SILERO_B200_ERR_CUDA;
   }

   timing_begin( h );
   if ( nchunks > 0 )
   {
      const int nw_max = pick_window( h, nstreams, nchunks, true );
      const size_t win_samples = (size_t)nstreams * nw_max * VB_CHUNK;
      if ( win_samples > h->pcm_stage_cap )
      {
         CU( cudaStreamSynchronize( h->stream ) );
         CU( cudaStreamSynchronize( h->copy_stream ) );
         for ( int i = 0; i < 2; ++i )
         {
            if ( h->pcm_stage[i] ) CU( cudaFree( h->pcm_stage[i] ) );
            h->pcm_stage[i] = 0;
         }
         h->pcm_stage_cap = 0;
         for ( int i = 0; i < 2; ++i ) CU( cudaMalloc( &h->pcm_stage[i], win_samples * sizeof( int16_t ) ) );
         h->pcm_stage_cap = win_samples;
         h->stage_used[0] = h->stage_used[1] = 0;
         cudaEventRecord( h->ev_begin, h->stream );
      }
      // Window plan: short windows at both ends so that neither the first copy (nothing to overlap with) nor the
      // last compute pass (no copy left to hide behind) costs a full window; full-size windows in between.
      int wsize[64 + 8], nwin = 0;
      {
         const int q = nw_max / 4 > 0 ? nw_max / 4 : 1, hh = nw_max / 2 > 0 ? nw_max / 2 : 1;
         const long long mid = (long long)nchunks - 2ll * ( q + hh );
         // (asynchronous calls are pipelined against their neighbours instead: there the copy of window w+1 hides behind the
         // compute of window w only if consecutive windows have the same size, so they keep uniform windows)
         if ( ticket || mid < nw_max || ( mid + nw_max - 1 ) / nw_max > 64 )
            nwin = -1; // short call (or very long one): uniform windows
         else
         {
            const int nmid = (int)( ( mid + nw_max - 1 ) / nw_max );
            wsize[nwin++] = q;
            wsize[nwin++] = hh;
            for ( int i = 0; i < nmid; ++i ) wsize[nwin++] = (int)( mid / nmid ) + ( i < mid % nmid ? 1 : 0 );
            wsize[nwin++] = hh;
            wsize[nwin++] = q;
         }
      }
      const bool planned = nwin > 0;
      // otherwise: equal windows (sizes differ by at most one chunk), so that copy and compute of neighbouring windows pair up
      if ( !planned ) nwin = ( nchunks + nw_max - 1 ) / nw_max;
      const int wbase = nchunks / nwin, wextra = nchunks % nwin;
      auto win_begin = [&]( int w ) -> int {
         if ( !planned ) return w * wbase + imin( w, wextra );
         int n0 = 0;
         for ( int i = 0; i < w; ++i ) n0 += wsize[i];
         return n0;
      };
      auto win_size = [&]( int w ) -> int { return planned ? wsize[w] : wbase + ( w < wextra ? 1 : 0 ); };
      // window w is copied on copy_stream into one staging buffer while window w-1 computes out of the other. The alternation
      // continues across calls (stage0), so the first copy of an asynchronous call never waits for the previous call's last window.
      const int stage0 = h->stage_seq & 1;
      h->stage_seq += nwin;
      auto issue_copy = [&]( int w ) -> int {
         int n0 = win_begin( w ), nw = win_size( w ), b = ( w + stage0 ) & 1;
         // the staging buffer may still be read by an earlier window -- of this call or of a previous asynchronous one
         if ( h->stage_used[b] ) CU( cudaStreamWaitEvent( h->copy_stream, h->pcm_free[b], 0 ) );
         CU( cudaMemcpy2DAsync( h->pcm_stage[b], (size_t)nw * VB_CHUNK * sizeof( int16_t ), pcm + (long long)n0 * VB_CHUNK,
                                (size_t)stream_stride * sizeof( int16_t ), (size_t)nw * VB_CHUNK * sizeof( int16_t ), (size_t)nstreams,
                                cudaMemcpyHostToDevice, h->copy_stream ) );
         CU( cudaEventRecord( h->pcm_ready[b], h->copy_stream ) );
         return 0;
      };
      if ( issue_copy( 0 ) ) return SILERO_B200_ERR_CUDA;
      const bool need_probs = probs || seg;
      for ( int w = 0; w < nwin; ++w )
      {
         int n0 = win_begin( w ), nw = win_size( w ), b = ( w + stage0 ) & 1;
         if ( w + 1 < nwin && issue_copy( w + 1 ) ) return SILERO_B200_ERR_CUDA;
         // the staging buffer is free again as soon as the STFT has consumed it (run_window records pcm_free there)
         rc = run_window( h, h->pcm_stage[b], 0, (long long)nw * VB_CHUNK, first_stream, nstreams, nw, out2 ? h->d_out2 : 0, need_probs ? h->d_probs : 0, nchunks, n0, 1,
                          h->pcm_ready[b], h->pcm_free[b] );
         if ( rc ) return rc;
         h->stage_used[b] = 1;
      }
   }
   if ( seg )
   {
      if ( launch_segments( h, h->d_probs, nchunks, 0, first_stream, nstreams, nchunks, seg->finish, h->d_segs, seg->cap, h->d_counts ) ) return SILERO_B200_ERR_CUDA;
      CU( cudaMemcpyAsync( seg->counts, h->d_counts, (size_t)nstreams * sizeof( int ), cudaMemcpyDeviceToHost, h->stream ) );
      CU( cudaMemcpyAsync( seg->segs, h->d_segs, (size_t)nstreams * seg->cap * sizeof( SegPair ), cudaMemcpyDeviceToHost, h->stream ) );
   }
   if ( probs && nout ) CU( cudaMemcpyAsync( probs, h->d_probs, nout * sizeof( float ), cudaMemcpyDeviceToHost, h->stream ) );
   if ( out2 && nout ) CU( cudaMemcpyAsync( out2, h->d_out2, nout * 2 * sizeof( float ), cudaMemcpyDeviceToHost, h->stream ) );
   timing_end( h );
   if ( ticket )
   {
      // asynchronous: hand out a completion ticket instead of waiting
      CU( cudaEventRecord( h->done_ev[h->ticket_seq % 4], h->stream ) );
      *ticket = h->ticket_seq++;
      return SILERO_B200_OK;
   }
   CU( cudaStreamSynchronize( h->stream ) );
   return check_err_word( h );
}

extern "C" int silero_b200_submit_streams_segments( silero_b200 *h, const int16_t *pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                                    int end_of_stream, vadc_segment *segs, int cap, int *counts, float *probs, unsigned long long *ticket )
{
   if ( !ticket ) return set_err( SILERO_B200_ERR_ARG, "null ticket" );
   if ( h && h->ticket_seq >= 4 )
   {
      // at most 4 calls in flight: the event about to be reused must have completed
      if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
      CU( cudaEventSynchronize( h->done_ev[h->ticket_seq % 4] ) );
   }
   *ticket = ~0ull;
   if ( segs )
   {
      SegRequest rq = { end_of_stream, cap, segs, counts };
      return run_streams_host( h, pcm, stream_stride, first_stream, nstreams, nchunks, probs, 0, &rq, ticket );
   }
   return run_streams_host( h, pcm, stream_stride, first_stream, nstreams, nchunks, probs, 0, 0, ticket );
}

extern "C" int silero_b200_wait( silero_b200 *h, unsigned long long ticket )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null handle" );
   if ( ticket == ~0ull ) return SILERO_B200_OK; // the submit had nothing to do
   if ( ticket >= h->ticket_seq ) return set_err( SILERO_B200_ERR_ARG, "unknown ticket" );
   if ( h->ticket_seq > ticket + 4 ) return SILERO_B200_OK; // its event slot has been reused, and submit waited on it before reusing it
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaEventSynchronize( h->done_ev[ticket % 4] ) );
   return check_err_word( h );
}

extern "C" int silero_b200_run_streams( silero_b200 *h, const int16_t *pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                        float *probs, float *out2 )
{
   return run_streams_host( h, pcm, stream_stride, first_stream, nstreams, nchunks, probs, out2, 0 );
}

static_assert( sizeof( SegPair ) == sizeof( vadc_segment ), "device and host segment records must have the same layout" );

extern "C" int silero_b200_run_streams_segments( silero_b200 *h, const int16_t *pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                                 int end_of_stream, vadc_segment *segs, int cap, int *counts, float *probs )
{
   SegRequest rq = { end_of_stream, cap, segs, counts };
   return run_streams_host( h, pcm, stream_stride, first_stream, nstreams, nchunks, probs, 0, &rq );
}

extern "C" int silero_b200_segment_probs_device( silero_b200 *h, const float *d_probs, long long stride, int first_stream, int nstreams, int nchunks,
                                                 int end_of_stream, vadc_segment *d_segs, int cap, int *d_counts )
{
   int rc = check_streams( h, first_stream, nstreams, nchunks );
   if ( rc ) return rc;
   if ( nstreams == 0 ) return SILERO_B200_OK;
   if ( ( nchunks > 0 && !d_probs ) || !d_segs || !d_counts || cap < 1 ) return set_err( SILERO_B200_ERR_ARG, "segments: null buffer or cap < 1" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   return launch_segments( h, d_probs, stride, 0, first_stream, nstreams, nchunks, end_of_stream, reinterpret_cast<SegPair *>( d_segs ), cap, d_counts );
}

extern "C" int silero_b200_run_streams_segments_device( silero_b200 *h, const int16_t *d_pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                                        int end_of_stream, float *d_probs, vadc_segment *d_segs, int cap, int *d_counts )
{
   int rc = check_streams( h, first_stream, nstreams, nchunks );
   if ( rc ) return rc;
   if ( nstreams == 0 ) return SILERO_B200_OK;
   if ( !d_segs || !d_counts || cap < 1 ) return set_err( SILERO_B200_ERR_ARG, "segments: null buffer or cap < 1" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   if ( !d_probs )
   {
      if ( grow( &h->d_probs, &h->d_probs_cap, (size_t)nstreams * ( nchunks ? nchunks : 1 ) ) ) return SILERO_B200_ERR_CUDA;
      d_probs = h->d_probs;
   }
   if ( nchunks > 0 )
   {
      rc = silero_b200_run_streams_device( h, d_pcm, stream_stride, first_stream, nstreams, nchunks, d_probs, 0 );
      if ( rc ) return rc;
   }
   if ( launch_segments( h, d_probs, nchunks, 0, first_stream, nstreams, nchunks, end_of_stream, reinterpret_cast<SegPair *>( d_segs ), cap, d_counts ) )
      return SILERO_B200_ERR_CUDA;
   timing_end( h );
   return SILERO_B200_OK;
}

extern "C" int silero_b200_segments_configure( silero_b200 *h, const vadc_seg_params *params )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null handle" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   set_seg_params( h, params );
   CU( cudaMemsetAsync( h->d_seg_state, 0, (size_t)h->max_streams * sizeof( SegStateDev ), h->stream ) );
   CU( cudaStreamSynchronize( h->stream ) );
   return SILERO_B200_OK;
}

extern "C" int silero_b200_segments_reset( silero_b200 *h, int first_stream, int nstreams )
{
   int rc = check_streams( h, first_stream, nstreams, 0 );
   if ( rc ) return rc;
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaMemsetAsync( h->d_seg_state + first_stream, 0, (size_t)nstreams * sizeof( SegStateDev ), h->stream ) );
   CU( cudaStreamSynchronize( h->stream ) );
   return SILERO_B200_OK;
}

extern "C" int silero_b200_run_chunks( silero_b200 *h, int stream, const float *samples, int nchunks, float *out )
{
   int rc = check_streams( h, stream, 1, nchunks );
   if ( rc ) return rc;
   if ( nchunks == 0 ) return SILERO_B200_OK;
   if ( !samples || !out ) return set_err( SILERO_B200_ERR_ARG, "null buffer" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t n = (size_t)nchunks * VB_CHUNK;
   if ( grow( &h->d_f32, &h->d_f32_cap, n ) ) return SILERO_B200_ERR_CUDA;
   if ( grow( &h->d_out2, &h->d_out2_cap, (size_t)nchunks * 2 ) ) return SILERO_B200_ERR_CUDA;
   timing_begin( h );
   CU( cudaMemcpyAsync( h->d_f32, samples, n * sizeof( float ), cudaMemcpyHostToDevice, h->stream ) );
   const int nw_max = pick_window( h, 1, nchunks );
   for ( int n0 = 0; n0 < nchunks; n0 += nw_max )
   {
      int nw = imin( nw_max, nchunks - n0 );
      rc = run_window( h, h->d_f32 + (size_t)n0 * VB_CHUNK, 1, 0, stream, 1, nw, h->d_out2, 0, nchunks, n0, 1, 0, 0, true );
      if ( rc ) return rc;
   }
   CU( cudaMemcpyAsync( out, h->d_out2, (size_t)nchunks * 2 * sizeof( float ), cudaMemcpyDeviceToHost, h->stream ) );
   timing_end( h );
   CU( cudaStreamSynchronize( h->stream ) );
   return SILERO_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------------------------
extern "C" int silero_b200_reset( silero_b200 *h, int first_stream, int nstreams )
{
   int rc = check_streams( h, first_stream, nstreams, 0 );
   if ( rc ) return rc;
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   size_t off = (size_t)first_stream * SILERO_B200_STATE_FLOATS, n = (size_t)nstreams * SILERO_B200_STATE_FLOATS * sizeof( float );
   CU( cudaMemsetAsync( h->state_h + off, 0, n, h->stream ) );
   CU( cudaMemsetAsync( h->state_c + off, 0, n, h->stream ) );
   CU( cudaStreamSynchronize( h->stream ) );
   return SILERO_B200_OK;
}

extern "C" int silero_b200_get_state( silero_b200 *h, int stream, float *h_out, float *c_out )
{
   int rc = check_streams( h, stream, 1, 0 );
   if ( rc ) return rc;
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaStreamSynchronize( h->stream ) );
   size_t off = (size_t)stream * SILERO_B200_STATE_FLOATS;
   if ( h_out ) CU( cudaMemcpy( h_out, h->state_h + off, SILERO_B200_STATE_FLOATS * sizeof( float ), cudaMemcpyDeviceToHost ) );
   if ( c_out ) CU( cudaMemcpy( c_out, h->state_c + off, SILERO_B200_STATE_FLOATS * sizeof( float ), cudaMemcpyDeviceToHost ) );
   return SILERO_B200_OK;
}

extern "C" int silero_b200_set_state( silero_b200 *h, int stream, const float *h_in, const float *c_in )
{
   int rc = check_streams( h, stream, 1, 0 );
   if ( rc ) return rc;
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaStreamSynchronize( h->stream ) );
   size_t off = (size_t)stream * SILERO_B200_STATE_FLOATS;
   if ( h_in ) CU( cudaMemcpy( h->state_h + off, h_in, SILERO_B200_STATE_FLOATS * sizeof( float ), cudaMemcpyHostToDevice ) );
   if ( c_in ) CU( cudaMemcpy( h->state_c + off, c_in, SILERO_B200_STATE_FLOATS * sizeof( float ), cudaMemcpyHostToDevice ) );
   return SILERO_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
extern "C" int silero_b200_device_alloc( silero_b200 *h, size_t nbytes, void **d_ptr )
{
   if ( !h || !d_ptr ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaMalloc( d_ptr, nbytes ) );
   return SILERO_B200_OK;
}
extern "C" int silero_b200_device_free( silero_b200 *h, void *d_ptr )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaFree( d_ptr ) );
   return SILERO_B200_OK;
}
extern "C" int silero_b200_memcpy_h2d( silero_b200 *h, void *d_dst, const void *src, size_t nbytes )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaMemcpyAsync( d_dst, src, nbytes, cudaMemcpyHostToDevice, h->stream ) );
   CU( cudaStreamSynchronize( h->stream ) );
   return SILERO_B200_OK;
}
extern "C" int silero_b200_memcpy_d2h( silero_b200 *h, void *dst, const void *d_src, size_t nbytes )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaMemcpyAsync( dst, d_src, nbytes, cudaMemcpyDeviceToHost, h->stream ) );
   CU( cudaStreamSynchronize( h->stream ) );
   return SILERO_B200_OK;
}
extern "C" int silero_b200_host_alloc_pinned( size_t nbytes, void **ptr )
{
   if ( !ptr ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   CU( cudaHostAlloc( ptr, nbytes, cudaHostAllocDefault ) );
   return SILERO_B200_OK;
}
extern "C" int silero_b200_host_free_pinned( void *ptr )
{
   CU( cudaFreeHost( ptr ) );
   return SILERO_B200_OK;
}

// parity tap for libm_exact.cuh: expf_ref / tanhf_ref of n host floats
__global__ void libm_exact_kernel( const float *__restrict__ x, float *__restrict__ e, float *__restrict__ t, float *__restrict__ l, int n )
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if ( i >= n ) return;
   e[i] = lme::expf_ref( x[i] );
   t[i] = lme::tanhf_ref( x[i] );
   l[i] = lme::log1pf_ref( fabsf( x[i] ) );
}
extern "C" int silero_b200_stage_libm( silero_b200 *h, const float *x, int n, float *out_expf, float *out_tanhf, float *out_log1pf_abs )
{
   if ( !h || !x || !out_expf || !out_tanhf || !out_log1pf_abs || n < 0 ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( n == 0 ) return SILERO_B200_OK;
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   float *d = 0;
   CU( cudaMalloc( &d, (size_t)n * 4 * sizeof( float ) ) );
   cudaError_t ce = cudaMemcpyAsync( d, x, (size_t)n * sizeof( float ), cudaMemcpyHostToDevice, h->stream );
   if ( ce == cudaSuccess )
   {
      libm_exact_kernel<<<( n + 255 ) / 256, 256, 0, h->stream>>>( d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, n );
      ce = cudaGetLastError();
   }
   if ( ce == cudaSuccess ) ce = cudaMemcpyAsync( out_expf, d + n, (size_t)n * sizeof( float ), cudaMemcpyDeviceToHost, h->stream );
   if ( ce == cudaSuccess ) ce = cudaMemcpyAsync( out_tanhf, d + 2 * (size_t)n, (size_t)n * sizeof( float ), cudaMemcpyDeviceToHost, h->stream );
   if ( ce == cudaSuccess ) ce = cudaMemcpyAsync( out_log1pf_abs, d + 3 * (size_t)n, (size_t)n * sizeof( float ), cudaMemcpyDeviceToHost, h->stream );
   if ( ce == cudaSuccess ) ce = cudaStreamSynchronize( h->stream );
   cudaFree( d );
   CU( ce );
   return SILERO_B200_OK;
}

extern "C" int silero_b200_stft_stats( silero_b200 *h, unsigned long long *bins_total, unsigned long long *bins_exact, int reset )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null handle" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaStreamSynchronize( h->stream ) );
   unsigned long long n = 0;
   CU( cudaMemcpy( &n, h->d_flagged, sizeof( n ), cudaMemcpyDeviceToHost ) );
   if ( bins_total ) *bins_total = h->bins_total;
   if ( bins_exact ) *bins_exact = n;
   if ( reset )
   {
      CU( cudaMemset( h->d_flagged, 0, sizeof( n ) ) );
      h->bins_total = 0;
   }
   return SILERO_B200_OK;
}

extern "C" int silero_b200_set_profiling( silero_b200 *h, int enabled )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null handle" );
   h->profiling = enabled ? 1 : 0;
   return SILERO_B200_OK;
}

extern "C" int silero_b200_last_timing( silero_b200 *h, float ms[8], long long *kernel_launches )
{
   if ( !h || !ms ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( !h->timing_valid ) return set_err( SILERO_B200_ERR_ARG, "no completed run to report" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaEventSynchronize( h->ev_end ) );
   float total = 0.0f;
   CU( cudaEventElapsedTime( &total, h->ev_begin, h->ev_end ) );
   h->stage_ms[0] = total;
   for ( int i = 0; i < 8; ++i ) ms[i] = h->stage_ms[i];
   if ( kernel_launches ) *kernel_launches = h->launches;
   return SILERO_B200_OK;
}

struct DevBuf
{
   float *p = 0;
   ~DevBuf() { cudaFree( p ); }
   int alloc( size_t n )
   {
      CU( cudaMalloc( &p, ( n ? n : 1 ) * sizeof( float ) ) );
      return 0;
   }
};

// ---------------------------------------------------------------------------------------------
// measurement helpers (bench.py): device-side timer on the engine's stream, FP32 pipe peak
// ---------------------------------------------------------------------------------------------
extern "C" int silero_b200_timer_start( silero_b200 *h )
{
   if ( !h ) return set_err( SILERO_B200_ERR_ARG, "null handle" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaEventRecord( h->ev_stage[N_STAGE_EVENTS - 1], h->stream ) );
   return SILERO_B200_OK;
}

extern "C" int silero_b200_timer_stop( silero_b200 *h, float *ms )
{
   if ( !h || !ms ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaEventRecord( h->ev_stage[N_STAGE_EVENTS - 2], h->stream ) );
   CU( cudaEventSynchronize( h->ev_stage[N_STAGE_EVENTS - 2] ) );
   CU( cudaEventElapsedTime( ms, h->ev_stage[N_STAGE_EVENTS - 1], h->ev_stage[N_STAGE_EVENTS - 2] ) );
   return SILERO_B200_OK;
}

// 16 independent FFMA chains per thread: the FP32 FMA-pipe roofline this part actually delivers
__global__ void __launch_bounds__( 256 ) fp32_peak_kernel( float *out, int iters, float a, float b )
{
   float acc[16];
#pragma unroll
   for ( int i = 0; i < 16; ++i ) acc[i] = (float)( threadIdx.x + i );
   for ( int it = 0; it < iters; ++it )
   {
#pragma unroll
      for ( int i = 0; i < 16; ++i ) acc[i] = fmaf( acc[i], a, b );
   }
   float s = 0.0f;
#pragma unroll
   for ( int i = 0; i < 16; ++i ) s += acc[i];
   if ( s == 123.456f ) out[0] = s; // never true; keeps the chains alive
}

// the same with the multiply and the add as separately rounded instructions (FMUL, FADD): what the exact path's kernels may use.
// One FLOP per instruction: half the FMA figure is the roofline of arithmetic that must not be contracted.
__global__ void __launch_bounds__( 256 ) fp32_unfused_peak_kernel( float *out, int iters, float a, float b )
{
   float acc[16];
#pragma unroll
   for ( int i = 0; i < 16; ++i ) acc[i] = (float)( threadIdx.x + i );
   for ( int it = 0; it < iters; ++it )
   {
#pragma unroll
      for ( int i = 0; i < 16; ++i ) acc[i] = __fadd_rn( __fmul_rn( acc[i], a ), b );
   }
   float s = 0.0f;
#pragma unroll
   for ( int i = 0; i < 16; ++i ) s += acc[i];
   if ( s == 123.456f ) out[0] = s;
}

extern "C" int silero_b200_measure_fp32_unfused_peak( silero_b200 *h, float *tflops )
{
   if ( !h || !tflops ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   DevBuf o;
   if ( o.alloc( 4 ) ) return SILERO_B200_ERR_CUDA;
   const int iters = 1 << 14, blocks = h->sm_count * 8, threads = 256;
   float best = 0.0f;
   for ( int rep = 0; rep < 5; ++rep )
   {
      CU( cudaEventRecord( h->ev_stage[0], h->stream ) );
      fp32_unfused_peak_kernel<<<blocks, threads, 0, h->stream>>>( o.p, iters, 0.999f, 0.001f );
      CU( cudaEventRecord( h->ev_stage[1], h->stream ) );
      CU( cudaEventSynchronize( h->ev_stage[1] ) );
      float ms = 0.0f;
      CU( cudaEventElapsedTime( &ms, h->ev_stage[0], h->ev_stage[1] ) );
      float tf = 2.0f * 16.0f * (float)iters * (float)blocks * (float)threads / ( ms * 1e-3f ) / 1e12f;
      if ( rep > 0 && tf > best ) best = tf;
   }
   *tflops = best;
   return SILERO_B200_OK;
}

extern "C" int silero_b200_measure_fp32_peak( silero_b200 *h, float *tflops )
{
   if ( !h || !tflops ) return set_err( SILERO_B200_ERR_ARG, "null argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   DevBuf o;
   if ( o.alloc( 4 ) ) return SILERO_B200_ERR_CUDA;
   const int iters = 1 << 15, blocks = h->sm_count * 8, threads = 256;
   float best = 0.0f;
   for ( int rep = 0; rep < 5; ++rep )
   {
      CU( cudaEventRecord( h->ev_stage[0], h->stream ) );
      fp32_peak_kernel<<<blocks, threads, 0, h->stream>>>( o.p, iters, 0.999f, 0.001f );
      CU( cudaEventRecord( h->ev_stage[1], h->stream ) );
      CU( cudaEventSynchronize( h->ev_stage[1] ) );
      float ms = 0.0f;
      CU( cudaEventElapsedTime( &ms, h->ev_stage[0], h->ev_stage[1] ) );
      float tf = 2.0f * 16.0f * (float)iters * (float)blocks * (float)threads / ( ms * 1e-3f ) / 1e12f;
      if ( rep > 0 && tf > best ) best = tf;
   }
   *tflops = best;
   return SILERO_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// parity taps (test-facing; layouts converted on the host, compute on the device)
// ---------------------------------------------------------------------------------------------

static int up( silero_b200 *h, float *d, const float *src, size_t n )
{
   CU( cudaMemcpyAsync( d, src, n * sizeof( float ), cudaMemcpyHostToDevice, h->stream ) );
   return 0;
}
static int down( silero_b200 *h, float *dst, const float *d, size_t n )
{
   CU( cudaMemcpyAsync( dst, d, n * sizeof( float ), cudaMemcpyDeviceToHost, h->stream ) );
   CU( cudaStreamSynchronize( h->stream ) );
   return 0;
}

// [B][T][C] (token-major, engine) <-> [B][C][T] (reference)
static void tok_to_ref( const float *tok, int B, int T, int C, float *ref )
{
   for ( int b = 0; b < B; ++b )
      for ( int t = 0; t < T; ++t )
         for ( int c = 0; c < C; ++c ) ref[( (size_t)b * C + c ) * T + t] = tok[( (size_t)b * T + t ) * C + c];
}
static void ref_to_tok( const float *ref, int B, int C, int T, float *tok )
{
   for ( int b = 0; b < B; ++b )
      for ( int c = 0; c < C; ++c )
         for ( int t = 0; t < T; ++t ) tok[( (size_t)b * T + t ) * C + c] = ref[( (size_t)b * C + c ) * T + t];
}

__global__ void subtract_chunk_mean_kernel( const float *__restrict__ logmag, float *__restrict__ out, int nchunks );
__global__ void log1p_scale_kernel( const float *__restrict__ mag, float *__restrict__ out, size_t n );

extern "C" int silero_b200_stage_stft_norm( silero_b200 *h, const float *samples, int batch, float *norm_out, float *logmag_out )
{
   if ( !h || !samples || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t nin = (size_t)batch * VB_CHUNK, nsp = (size_t)batch * VB_BINS * VB_FRAMES;
   DevBuf in, sp, nm;
   if ( in.alloc( nin ) || sp.alloc( nsp ) || nm.alloc( nsp ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, in.p, samples, nin ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_stft( h, in.p, 1, 0, batch, batch, sp.p, 0 ) ) return SILERO_B200_ERR_CUDA;
   if ( logmag_out && down( h, logmag_out, sp.p, nsp ) ) return SILERO_B200_ERR_CUDA;
   if ( norm_out )
   {
      subtract_chunk_mean_kernel<<<batch, 32, 0, h->stream>>>( sp.p, nm.p, batch );
      CU( cudaGetLastError() );
      if ( down( h, norm_out, nm.p, nsp ) ) return SILERO_B200_ERR_CUDA;
   }
   return SILERO_B200_OK;
}

// tap for stft.c alone: raw magnitudes
extern "C" int silero_b200_stage_stft_magnitude( silero_b200 *h, const float *samples, int batch, float *mag_out )
{
   if ( !h || !samples || !mag_out || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t nin = (size_t)batch * VB_CHUNK, nsp = (size_t)batch * VB_BINS * VB_FRAMES;
   DevBuf in, sp;
   if ( in.alloc( nin ) || sp.alloc( nsp ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, in.p, samples, nin ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_stft( h, in.p, 1, 0, batch, batch, sp.p, 1 ) ) return SILERO_B200_ERR_CUDA;
   return down( h, mag_out, sp.p, nsp ) ? SILERO_B200_ERR_CUDA : SILERO_B200_OK;
}

__global__ void log1p_scale_kernel( const float *__restrict__ mag, float *__restrict__ out, size_t n )
{
   size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if ( i < n ) out[i] = lme::log1pf_ref( __fmul_rn( mag[i], 1048576.0f ) ); // misc.c:40-46 with the C library's bits
}

// the mean part of adaptive_audio_normalization_inplace (misc.c:48-121), one warp per chunk; the
// production path computes the same quantity inside layer_kernel<0,true>
__global__ void subtract_chunk_mean_kernel( const float *__restrict__ logmag, float *__restrict__ out, int nchunks )
{
   __shared__ float m[VB_FRAMES], s[VB_FRAMES];
   const int ci = blockIdx.x, t = threadIdx.x;
   if ( ci >= nchunks ) return;
   const float *sp = logmag + (size_t)ci * VB_BINS * VB_FRAMES;
   if ( t < VB_FRAMES )
   {
      float a = 0.0f;
      for ( int f = 0; f < VB_BINS; ++f ) a = __fadd_rn( a, sp[f * VB_FRAMES + t] );
      m[t] = a / (float)VB_BINS;
   }
   __syncwarp();
   if ( t < VB_FRAMES )
   {
      const float gk[7] = { 0.03663284704089164733887f, 0.11128076165914535522461f, 0.21674531698226928710938f, 0.27068215608596801757812f,
                            0.21674531698226928710938f, 0.11128076165914535522461f, 0.03663284704089164733887f };
      float v = 0.0f;
      for ( int k = 0; k < 7; ++k )
      {
         int idx = t + k - 3;
         if ( idx < 0 ) idx = -idx;
         if ( idx >= VB_FRAMES ) idx = 2 * ( VB_FRAMES - 1 ) - idx;
         v = __fadd_rn( v, __fmul_rn( m[idx], gk[k] ) );
      }
      s[t] = v;
   }
   __syncwarp();
   float mu = 0.0f;
   for ( int i = 0; i < VB_FRAMES; ++i ) mu = __fadd_rn( mu, s[i] );
   mu = mu / (float)VB_FRAMES;
   for ( int i = t; i < VB_BINS * VB_FRAMES; i += 32 ) out[(size_t)ci * VB_BINS * VB_FRAMES + i] = sp[i] - mu;
}

extern "C" int silero_b200_stage_norm( silero_b200 *h, const float *magnitude, int batch, float *norm_out )
{
   if ( !h || !magnitude || !norm_out || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t nsp = (size_t)batch * VB_BINS * VB_FRAMES;
   DevBuf mg, lg, nm;
   if ( mg.alloc( nsp ) || lg.alloc( nsp ) || nm.alloc( nsp ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, mg.p, magnitude, nsp ) ) return SILERO_B200_ERR_CUDA;
   log1p_scale_kernel<<<(unsigned)( ( nsp + 255 ) / 256 ), 256, 0, h->stream>>>( mg.p, lg.p, nsp );
   subtract_chunk_mean_kernel<<<batch, 32, 0, h->stream>>>( lg.p, nm.p, batch );
   CU( cudaGetLastError() );
   return down( h, norm_out, nm.p, nsp ) ? SILERO_B200_ERR_CUDA : SILERO_B200_OK;
}

// production kernels from raw samples to every encoder layer output (stft -> layer<0,NORM> -> ...)
extern "C" int silero_b200_stage_pipeline( silero_b200 *h, const float *samples, int batch, float *l1, float *l2, float *l3, float *l4 )
{
   if ( !h || !samples || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t B = (size_t)batch;
   DevBuf in, sp, d1, d2, d3, d4;
   if ( in.alloc( B * VB_CHUNK ) || sp.alloc( B * 3225 ) || d1.alloc( B * 208 ) || d2.alloc( B * 224 ) || d3.alloc( B * 224 ) || d4.alloc( B * 448 ) )
      return SILERO_B200_ERR_CUDA;
   if ( up( h, in.p, samples, B * VB_CHUNK ) ) return SILERO_B200_ERR_CUDA;
   const bool have_mu = stft_produces_mu( h, batch );
   DevBuf mu;
   if ( mu.alloc( B ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_stft( h, in.p, 1, 0, batch, batch, sp.p, 0, have_mu ? mu.p : 0 ) ) return SILERO_B200_ERR_CUDA;
   if ( first_layer_from_logspec( h, sp.p, d1.p, batch, have_mu ? mu.p : 0 ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_layer_any<1>( h, d1.p, d2.p, batch ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_layer_any<2>( h, d2.p, d3.p, batch ) ) return SILERO_B200_ERR_CUDA;
   if ( launch_layer_any<3>( h, d3.p, d4.p, batch ) ) return SILERO_B200_ERR_CUDA;
   float *tmp = (float *)malloc( B * 448 * sizeof( float ) );
   if ( !tmp ) return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
   int rc = 0;
   if ( l1 && !( rc = down( h, tmp, d1.p, B * 208 ) ) ) tok_to_ref( tmp, batch, 13, 16, l1 );
   if ( !rc && l2 && !( rc = down( h, tmp, d2.p, B * 224 ) ) ) tok_to_ref( tmp, batch, 7, 32, l2 );
   if ( !rc && l3 && !( rc = down( h, tmp, d3.p, B * 224 ) ) ) tok_to_ref( tmp, batch, 7, 32, l3 );
   if ( !rc && l4 && !( rc = down( h, tmp, d4.p, B * 448 ) ) ) tok_to_ref( tmp, batch, 7, 64, l4 );
   free( tmp );
   return rc ? SILERO_B200_ERR_CUDA : SILERO_B200_OK;
}

// the exact path's encoder from samples (exact STFT -> exact_front_kernel -> exact_layer_kernel x 4), every stage tapped in the
// reference layout: y1 [B,16,25] = conv_block output of the first layer (conv.c:761), l1..l4 as in silero_b200_stage_pipeline
extern "C" int silero_b200_stage_exact_pipeline( silero_b200 *h, const float *samples, int batch, float *y1, float *l1, float *l2, float *l3, float *l4 )
{
   if ( !h || !samples || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t B = (size_t)batch;
   DevBuf in;
   if ( in.alloc( B * VB_CHUNK ) ) return SILERO_B200_ERR_CUDA;
   if ( ensure_scratch( h, B, 0 ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, in.p, samples, B * VB_CHUNK ) ) return SILERO_B200_ERR_CUDA;
   const int saved = h->stft_mode;
   h->stft_mode = SILERO_B200_STFT_EXACT;
   int rc = launch_stft( h, in.p, 1, 0, batch, batch, h->spec, 0, 0 );
   h->stft_mode = saved;
   if ( rc ) return SILERO_B200_ERR_CUDA;
   if ( launch_exact_encoder( h, h->spec, h->a4, batch ) ) return SILERO_B200_ERR_CUDA;
   if ( y1 && down( h, y1, h->y1, B * 400 ) ) return SILERO_B200_ERR_CUDA;
   if ( l1 && down( h, l1, h->a1, B * 208 ) ) return SILERO_B200_ERR_CUDA;
   if ( l2 && down( h, l2, h->a2, B * 224 ) ) return SILERO_B200_ERR_CUDA;
   if ( l3 && down( h, l3, h->a3, B * 224 ) ) return SILERO_B200_ERR_CUDA;
   if ( l4 )
   {
      float *tmp = (float *)malloc( B * 448 * sizeof( float ) );
      if ( !tmp ) return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
      rc = down( h, tmp, h->a4, B * 448 );
      if ( !rc ) tok_to_ref( tmp, batch, 7, 64, l4 );
      free( tmp );
      if ( rc ) return SILERO_B200_ERR_CUDA;
   }
   return SILERO_B200_OK;
}

// one transformer_layer on the exact path's kernels; layer 0 takes the NORMALIZED spectrogram [B,129,25] (exact_front_kernel without
// its normalization + exact_layer_kernel<0>), y1 (optional) its conv_block output [B,16,25]; layouts as silero_b200_stage_layer
extern "C" int silero_b200_stage_exact_layer( silero_b200 *h, int layer, const float *in, int batch, float *out, float *y1 )
{
   if ( !h || !in || batch <= 0 || layer < 0 || layer > 3 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t B = (size_t)batch;
   const LayerDims d = layer_dims( layer );
   const size_t n_in = B * d.cin * d.t, n_out = B * d.c * ( 1 + ( d.t - 1 ) / d.stride );
   DevBuf din, dy, dout;
   if ( din.alloc( n_in ) || dy.alloc( B * 400 ) || dout.alloc( n_out ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, din.p, in, n_in ) ) return SILERO_B200_ERR_CUDA;
   if ( !h->xe_scratch ) CU( cudaMalloc( &h->xe_scratch, XeCfg<3>::SCRATCH_FLOATS * sizeof( float ) * (size_t)h->sm_count ) );
   int rc = 0;
   if ( layer == 0 )
   {
      rc = launch_exact_front( h, din.p, dy.p, batch, false );
      if ( !rc ) rc = launch_exact_layer<0>( h, dy.p, dout.p, batch );
   }
   else if ( layer == 1 ) rc = launch_exact_layer<1>( h, din.p, dout.p, batch );
   else if ( layer == 2 ) rc = launch_exact_layer<2>( h, din.p, dout.p, batch );
   else rc = launch_exact_layer<3>( h, din.p, dout.p, batch );
   if ( rc ) return SILERO_B200_ERR_CUDA;
   if ( y1 && layer == 0 && down( h, y1, dy.p, B * 400 ) ) return SILERO_B200_ERR_CUDA;
   if ( out )
   {
      if ( layer < 3 ) { if ( down( h, out, dout.p, n_out ) ) return SILERO_B200_ERR_CUDA; }
      else
      {
         float *tmp = (float *)malloc( n_out * sizeof( float ) );
         if ( !tmp ) return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
         rc = down( h, tmp, dout.p, n_out );
         if ( !rc ) tok_to_ref( tmp, batch, 7, 64, out );
         free( tmp );
         if ( rc ) return SILERO_B200_ERR_CUDA;
      }
   }
   return SILERO_B200_OK;
}

// the exact path's encoder from a spectrogram [B,129,25]; kind 0: log1p spectrogram (the front kernel normalizes, as in production),
// 1: already normalized, 2: raw magnitude (log1p(m * 2^20) is applied first, misc.c:40-46)
extern "C" int silero_b200_stage_exact_encoder( silero_b200 *h, const float *spec, int batch, int kind, float *l1, float *l2, float *l3, float *l4 )
{
   if ( !h || !spec || batch <= 0 || kind < 0 || kind > 2 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t B = (size_t)batch;
   if ( ensure_scratch( h, B, 0 ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, h->spec, spec, B * 3225 ) ) return SILERO_B200_ERR_CUDA;
   if ( kind == 2 )
   {
      log1p_scale_kernel<<<(unsigned)( ( B * 3225 + 255 ) / 256 ), 256, 0, h->stream>>>( h->spec, h->spec, B * 3225 );
      CU( cudaGetLastError() );
   }
   if ( launch_exact_encoder( h, h->spec, h->a4, batch, kind != 1 ) ) return SILERO_B200_ERR_CUDA;
   if ( l1 && down( h, l1, h->a1, B * 208 ) ) return SILERO_B200_ERR_CUDA;
   if ( l2 && down( h, l2, h->a2, B * 224 ) ) return SILERO_B200_ERR_CUDA;
   if ( l3 && down( h, l3, h->a3, B * 224 ) ) return SILERO_B200_ERR_CUDA;
   if ( l4 )
   {
      float *tmp = (float *)malloc( B * 448 * sizeof( float ) );
      if ( !tmp ) return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
      const int rc = down( h, tmp, h->a4, B * 448 );
      if ( !rc ) tok_to_ref( tmp, batch, 7, 64, l4 );
      free( tmp );
      if ( rc ) return SILERO_B200_ERR_CUDA;
   }
   return SILERO_B200_OK;
}

static int run_encoder_from( silero_b200 *h, int first_layer, const float *d_in, int batch, float *d1, float *d2, float *d3, float *d4 )
{
   if ( first_layer <= 0 && ( layers_use_tensor( h, batch ) ? launch_layer0_tc( h, d_in, d1, batch, 0 ) : launch_layer<0, false>( h, d_in, d1, batch ) ) )
      return SILERO_B200_ERR_CUDA;
   if ( first_layer <= 1 && launch_layer_any<1>( h, first_layer == 1 ? d_in : d1, d2, batch ) ) return SILERO_B200_ERR_CUDA;
   if ( first_layer <= 2 && launch_layer_any<2>( h, first_layer == 2 ? d_in : d2, d3, batch ) ) return SILERO_B200_ERR_CUDA;
   if ( first_layer <= 3 && launch_layer_any<3>( h, first_layer == 3 ? d_in : d3, d4, batch ) ) return SILERO_B200_ERR_CUDA;
   return 0;
}

extern "C" int silero_b200_stage_encoder( silero_b200 *h, const float *norm, int batch, float *l1, float *l2, float *l3, float *l4 )
{
   if ( !h || !norm || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t B = (size_t)batch;
   DevBuf in, d1, d2, d3, d4;
   if ( in.alloc( B * 3225 ) || d1.alloc( B * 208 ) || d2.alloc( B * 224 ) || d3.alloc( B * 224 ) || d4.alloc( B * 448 ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, in.p, norm, B * 3225 ) ) return SILERO_B200_ERR_CUDA;
   if ( run_encoder_from( h, 0, in.p, batch, d1.p, d2.p, d3.p, d4.p ) ) return SILERO_B200_ERR_CUDA;
   float *tmp = (float *)malloc( B * 448 * sizeof( float ) );
   if ( !tmp ) return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
   int rc = 0;
   if ( l1 && !( rc = down( h, tmp, d1.p, B * 208 ) ) ) tok_to_ref( tmp, batch, 13, 16, l1 );
   if ( !rc && l2 && !( rc = down( h, tmp, d2.p, B * 224 ) ) ) tok_to_ref( tmp, batch, 7, 32, l2 );
   if ( !rc && l3 && !( rc = down( h, tmp, d3.p, B * 224 ) ) ) tok_to_ref( tmp, batch, 7, 32, l3 );
   if ( !rc && l4 && !( rc = down( h, tmp, d4.p, B * 448 ) ) ) tok_to_ref( tmp, batch, 7, 64, l4 );
   free( tmp );
   return rc ? SILERO_B200_ERR_CUDA : SILERO_B200_OK;
}

extern "C" int silero_b200_stage_layer( silero_b200 *h, int layer, const float *in, int batch, float *out )
{
   if ( !h || !in || !out || batch <= 0 || layer < 0 || layer > 3 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const LayerDims d = layer_dims( layer );
   const int tout = 1 + ( d.t - 1 ) / d.stride;
   const size_t nin = (size_t)batch * d.cin * d.t, nout = (size_t)batch * d.c * tout;
   DevBuf din, dout;
   if ( din.alloc( nin ) || dout.alloc( nout ) ) return SILERO_B200_ERR_CUDA;
   float *tmp = (float *)malloc( ( nin > nout ? nin : nout ) * sizeof( float ) );
   if ( !tmp ) return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
   int rc = 0;
   if ( layer == 0 )
      rc = up( h, din.p, in, nin ); // the first layer consumes the reference's [B,129,25] layout directly
   else
   {
      ref_to_tok( in, batch, d.cin, d.t, tmp );
      rc = up( h, din.p, tmp, nin );
      if ( !rc ) rc = cudaStreamSynchronize( h->stream ) != cudaSuccess;
   }
   if ( !rc )
   {
      switch ( layer )
      {
         case 0: rc = layers_use_tensor( h, batch ) ? launch_layer0_tc( h, din.p, dout.p, batch, 0 ) : launch_layer<0, false>( h, din.p, dout.p, batch ); break;
         case 1: rc = launch_layer_any<1>( h, din.p, dout.p, batch ); break;
         case 2: rc = launch_layer_any<2>( h, din.p, dout.p, batch ); break;
         default: rc = launch_layer_any<3>( h, din.p, dout.p, batch ); break;
      }
   }
   if ( !rc ) rc = down( h, tmp, dout.p, nout );
   if ( !rc ) tok_to_ref( tmp, batch, tout, d.c, out );
   free( tmp );
   return rc ? SILERO_B200_ERR_CUDA : SILERO_B200_OK;
}

// Sub-stage taps of one layer for the reference's op/block-level fixtures (layer_kernel.cuh).
// layer 0..3. in: entry 0 -> reference layout [B,cin,T];
// entry 1/2 -> token-major [B,T,C]. out: tap 0 -> reference layout [B,C,TOUT]; else token-major [B,T,C].
extern "C" int silero_b200_stage_layer_tap( silero_b200 *h, int layer, int entry, int tap, const float *in, int batch, float *out )
{
   if ( !h || !in || !out || batch <= 0 || layer < 0 || layer > 3 || entry < 0 || entry > 2 || tap < 0 || tap > 4 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const LayerDims d = layer_dims( layer );
   const int tout = 1 + ( d.t - 1 ) / d.stride;
   const size_t nin = (size_t)batch * ( entry == ENTRY_LAYER ? d.cin : d.c ) * d.t;
   const size_t nout = (size_t)batch * d.c * ( tap == TAP_LAYER ? tout : d.t );
   DevBuf din, dout;
   if ( din.alloc( nin ) || dout.alloc( nout ) ) return SILERO_B200_ERR_CUDA;
   float *tmp = (float *)malloc( ( nin > nout ? nin : nout ) * sizeof( float ) );
   if ( !tmp ) return set_err( SILERO_B200_ERR_NOMEM, "out of host memory" );
   int rc = 0;
   if ( entry == ENTRY_LAYER && d.cin != VB_BINS )
   {
      ref_to_tok( in, batch, d.cin, d.t, tmp );
      rc = up( h, din.p, tmp, nin );
      if ( !rc ) rc = cudaStreamSynchronize( h->stream ) != cudaSuccess;
   }
   else
      rc = up( h, din.p, in, nin );
   if ( !rc )
   {
      switch ( layer )
      {
         case 0: rc = launch_layer<0, false>( h, din.p, dout.p, batch, entry, tap ); break;
         case 1: rc = launch_layer<1, false>( h, din.p, dout.p, batch, entry, tap ); break;
         case 2: rc = launch_layer<2, false>( h, din.p, dout.p, batch, entry, tap ); break;
         default: rc = launch_layer<3, false>( h, din.p, dout.p, batch, entry, tap ); break;
      }
   }
   if ( !rc && tap == TAP_LAYER )
   {
      rc = down( h, tmp, dout.p, nout );
      if ( !rc ) tok_to_ref( tmp, batch, tout, d.c, out );
   }
   else if ( !rc )
      rc = down( h, out, dout.p, nout );
   free( tmp );
   return rc ? SILERO_B200_ERR_CUDA : SILERO_B200_OK;
}

extern "C" int silero_b200_stage_lstm( silero_b200 *h, const float *x, int batch, const float *h0, const float *c0, float *out, float *hn, float *cn )
{
   if ( !h || !x || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t n = (size_t)batch * 7 * 64;
   DevBuf dx, dh0, dtop, sh, sc;
   if ( dx.alloc( n ) || dh0.alloc( n ) || dtop.alloc( n ) || sh.alloc( 128 ) || sc.alloc( 128 ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, dx.p, x, n ) ) return SILERO_B200_ERR_CUDA;
   if ( h0 ) { if ( up( h, sh.p, h0, 128 ) ) return SILERO_B200_ERR_CUDA; } else CU( cudaMemsetAsync( sh.p, 0, 512, h->stream ) );
   if ( c0 ) { if ( up( h, sc.p, c0, 128 ) ) return SILERO_B200_ERR_CUDA; } else CU( cudaMemsetAsync( sc.p, 0, 512, h->stream ) );
   // run on a private state buffer: temporarily swap it in
   float *save_h = h->state_h, *save_c = h->state_c;
   h->state_h = sh.p;
   h->state_c = sc.p;
   int rc = launch_lstm<0>( h, dx.p, dh0.p, 0, 1, batch, 0, 0, 0, 0 );
   if ( !rc ) rc = launch_lstm<1>( h, dh0.p, dtop.p, 0, 1, batch, 0, 0, 0, 0 );
   h->state_h = save_h;
   h->state_c = save_c;
   if ( rc ) return rc;
   if ( out && down( h, out, dtop.p, n ) ) return SILERO_B200_ERR_CUDA;
   if ( hn && down( h, hn, sh.p, 128 ) ) return SILERO_B200_ERR_CUDA;
   if ( cn && down( h, cn, sc.p, 128 ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaStreamSynchronize( h->stream ) );
   return SILERO_B200_OK;
}

// softmax_inplace_stable (tensor.h:751-784) over the rows of a host matrix, in the exact path's arithmetic
extern "C" int silero_b200_stage_exact_softmax( silero_b200 *h, const float *x, int rows, int cols, float *out )
{
   if ( !h || !x || !out || rows <= 0 || cols <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   DevBuf d;
   if ( d.alloc( (size_t)rows * cols ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, d.p, x, (size_t)rows * cols ) ) return SILERO_B200_ERR_CUDA;
   exact_softmax_rows_kernel<<<( rows + 63 ) / 64, 64, 0, h->stream>>>( d.p, rows, cols );
   CU( cudaGetLastError() );
   return down( h, out, d.p, (size_t)rows * cols );
}

// the same on the exact path's kernel (exact_lstm_kernel: weights in registers; here one stream)
extern "C" int silero_b200_stage_exact_lstm( silero_b200 *h, const float *x, int batch, const float *h0, const float *c0, float *out, float *hn, float *cn, int wave )
{
   if ( !h || !x || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   const size_t n = (size_t)batch * 7 * 64;
   DevBuf dx, dh0, dtop, sh, sc;
   if ( dx.alloc( n ) || dh0.alloc( n ) || dtop.alloc( n ) || sh.alloc( 128 ) || sc.alloc( 128 ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, dx.p, x, n ) ) return SILERO_B200_ERR_CUDA;
   if ( h0 ) { if ( up( h, sh.p, h0, 128 ) ) return SILERO_B200_ERR_CUDA; } else CU( cudaMemsetAsync( sh.p, 0, 512, h->stream ) );
   if ( c0 ) { if ( up( h, sc.p, c0, 128 ) ) return SILERO_B200_ERR_CUDA; } else CU( cudaMemsetAsync( sc.p, 0, 512, h->stream ) );
   float *save_h = h->state_h, *save_c = h->state_c;
   h->state_h = sh.p;
   h->state_c = sc.p;
   int rc;
   // (in place: the input buffer receives the top layer's sequence)
   rc = launch_lstm_exact( h, dx.p, dh0.p, 0, 1, batch, wave ? 1 : 2 );
   if ( !rc ) rc = cudaMemcpyAsync( dtop.p, dx.p, n * sizeof( float ), cudaMemcpyDeviceToDevice, h->stream ) == cudaSuccess ? 0 : SILERO_B200_ERR_CUDA;
   h->state_h = save_h;
   h->state_c = save_c;
   if ( rc ) return rc;
   if ( out && down( h, out, dtop.p, n ) ) return SILERO_B200_ERR_CUDA;
   if ( hn && down( h, hn, sh.p, 128 ) ) return SILERO_B200_ERR_CUDA;
   if ( cn && down( h, cn, sc.p, 128 ) ) return SILERO_B200_ERR_CUDA;
   CU( cudaStreamSynchronize( h->stream ) );
   return SILERO_B200_OK;
}

// decoder alone (silero_v3.c:231-303): one thread per (chunk, head); the production path fuses the
// same arithmetic into lstm_layer_kernel<1>
__global__ void decoder_kernel( const float *__restrict__ in /*[B][64][7]*/, const float *__restrict__ w, const float *__restrict__ b, float *__restrict__ out, int batch )
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if ( i >= batch * 2 ) return;
   int n = i >> 1, head = i & 1;
   float sum = 0.0f;
   for ( int t = 0; t < 7; ++t )
   {
      float q = 0.0f;
      for ( int c = 0; c < 64; ++c ) q = fmaf( w[head * 64 + c], fmaxf( in[( (size_t)n * 64 + c ) * 7 + t], 0.0f ), q );
      sum += q + b[head];
   }
   float mean = sum / 7.0f;
   out[i] = 1.0f / ( 1.0f + expf( -mean ) );
}

extern "C" int silero_b200_stage_decoder( silero_b200 *h, const float *in, int batch, float *out )
{
   if ( !h || !in || !out || batch <= 0 ) return set_err( SILERO_B200_ERR_ARG, "bad argument" );
   if ( use_device( h ) ) return SILERO_B200_ERR_CUDA;
   DevBuf din, dout;
   if ( din.alloc( (size_t)batch * 448 ) || dout.alloc( (size_t)batch * 2 ) ) return SILERO_B200_ERR_CUDA;
   if ( up( h, din.p, in, (size_t)batch * 448 ) ) return SILERO_B200_ERR_CUDA;
   decoder_kernel<<<( batch * 2 + 127 ) / 128, 128, 0, h->stream>>>( din.p, h->w.dec_w, h->w.dec_b, dout.p, batch );
   CU( cudaGetLastError() );
   return down( h, out, dout.p, (size_t)batch * 2 ) ? SILERO_B200_ERR_CUDA : SILERO_B200_OK;
}

