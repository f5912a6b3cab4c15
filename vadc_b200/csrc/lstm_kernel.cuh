// vadc_b200/csrc/lstm_kernel.cuh -- the stateful 2-layer decoder LSTM + decoder head.
//
// Replaces lstm_tensor_minibatched / lstm_seq / lstm / lstm_cell (lstm.c:31-341) and
// decoder (silero_v3.c:231-303). Gate order i,f,g,o; W[l] is [256][128] with columns [0,64) on the
// layer input and [64,128) on that layer's previous h; b = b_ih + b_hh (utils.py:93-107).
//
// The recurrence is sequential per stream (7 steps per chunk, state carried for the whole stream),
// so throughput comes from batching streams. The two layers are separable in time: layer 0 is run
// over all steps of the window first (writing its h sequence), then layer 1 + the decoder head.
// One launch = one layer. A CTA keeps that layer's 128 KB of weights resident in shared memory as
// [k-quad][row][4] and owns NG groups of ST streams; thread (j = tid%64, group) owns hidden unit j:
// it accumulates all four gate rows of unit j for its ST streams (full K=128 in-thread, so the
// cell update needs no cross-thread exchange), keeps c in registers, and publishes h through a
// ping-pong [x|h] buffer in shared memory with one 64-thread named barrier per step.
#pragma once
#include "common.cuh"
#include "libm_exact.cuh"

#define LSTM_H 64
#define LSTM_WS_FLOATS ( 32 * 256 * 4 )
#define LSTM_MAX_GROUPS 8

template <int ST>
struct LstmSmem
{
   static constexpr int XH = 2 * LSTM_MAX_GROUPS * ST * 128; // ping-pong [group][stream][x(64)|h(64)]
   static constexpr int DEC = LSTM_MAX_GROUPS * 2 * 2 * ST;  // [group][warp][head][stream]
   static constexpr int FLOATS = LSTM_WS_FLOATS + XH + DEC + 256 + 128 + 4;
   static constexpr int BYTES = FLOATS * 4;
};

// Gate nonlinearities with the bits of the reference's C library and a cell update without fused multiply-adds (lstm.c:64-88 as
// compiled with -ffp-contract=off): this path is the one a single-stream user of the reference runs, and over a long silence a
// nonlinearity that is off by one ulp in the same direction on every step makes the cell state drift (libm_exact.cuh).
__device__ __forceinline__ float sigmoid_acc( float v ) { return lme::sigmoid_ref( v ); }

// x:      layer input sequence [S][steps][64] (stream-major), steps = nw*7
// hseq:   LAYER 0: output sequence [S][steps][64]; LAYER 1: optional tap of the top-layer sequence
// state_h, state_c: [S_total][2][64], rows first_stream.. are read and written
// out2:   LAYER 1: [S][out_stride][2] decoder outputs written at chunk n -> out2[(s*out_stride + out_off + n)*2 + {0,1}]
// probs:  LAYER 1: [S][out_stride] speech probability (head 1); either may be NULL
template <int LAYER, int ST>
__global__ void __launch_bounds__( 64 * LSTM_MAX_GROUPS, 1 )
lstm_layer_kernel( const float *__restrict__ x, float *__restrict__ hseq, float *__restrict__ state_h, float *__restrict__ state_c,
                   const float *__restrict__ wpack, const float *__restrict__ bias, const float *__restrict__ dec_w, const float *__restrict__ dec_b,
                   int nstreams, int nw, float *__restrict__ out2, float *__restrict__ probs, long long out_stride, long long out_off )
{
   extern __shared__ __align__( 16 ) float smem[];
   float *Ws = smem;
   float *xh = Ws + LSTM_WS_FLOATS;
   float *decp = xh + LstmSmem<ST>::XH;
   float *bs = decp + LstmSmem<ST>::DEC; // [256]
   float *dws = bs + 256;                // [2][64]
   float *dbs = dws + 128;               // [2]

   const int tid = threadIdx.x;
   const int j = tid & 63;
   const int grp = tid >> 6;
   const int ngroups = blockDim.x >> 6;
   const int steps = nw * 7;

   {
      const float4 *src = reinterpret_cast<const float4 *>( wpack ) + (size_t)LAYER * ( LSTM_WS_FLOATS / 4 );
      float4 *dst = reinterpret_cast<float4 *>( Ws );
      for ( int i = tid; i < LSTM_WS_FLOATS / 4; i += blockDim.x ) dst[i] = __ldg( src + i );
      for ( int i = tid; i < 256; i += blockDim.x ) bs[i] = bias[LAYER * 256 + i];
      if ( LAYER == 1 )
      {
         for ( int i = tid; i < 128; i += blockDim.x ) dws[i] = dec_w[i];
         if ( tid < 2 ) dbs[tid] = dec_b[tid];
      }
   }
   __syncthreads();

   const float bi = bs[j], bf = bs[64 + j], bg = bs[128 + j], bo = bs[192 + j];
   const float dw0 = ( LAYER == 1 ) ? dws[j] : 0.0f, dw1 = ( LAYER == 1 ) ? dws[64 + j] : 0.0f;

   // stream tiles: each group of 64 threads walks its own ST streams for the whole window
   const int tiles = ( nstreams + ST - 1 ) / ST;
   for ( int tile = blockIdx.x * ngroups + grp; tile < tiles; tile += gridDim.x * ngroups )
   {
      const int s0 = tile * ST;
      float c[ST];
      float *buf0 = xh + ( grp * ST ) * 128;
      float *buf1 = xh + ( ( LSTM_MAX_GROUPS + grp ) * ST ) * 128;
      // initial state + first input
#pragma unroll
      for ( int st = 0; st < ST; ++st )
      {
         int s = s0 + st;
         bool ok = s < nstreams;
         c[st] = ok ? state_c[( (size_t)s * 2 + LAYER ) * 64 + j] : 0.0f;
         buf0[st * 128 + 64 + j] = ok ? state_h[( (size_t)s * 2 + LAYER ) * 64 + j] : 0.0f;
         buf0[st * 128 + j] = ok ? __ldg( x + ( (size_t)s * steps ) * 64 + j ) : 0.0f;
      }
      bar_sync( 1 + grp, 64 );

      float d0[ST], d1[ST];
#pragma unroll
      for ( int st = 0; st < ST; ++st ) d0[st] = d1[st] = 0.0f;

      for ( int step = 0; step < steps; ++step )
      {
         float *cur = ( step & 1 ) ? buf1 : buf0;
         float *nxt = ( step & 1 ) ? buf0 : buf1;
         // prefetch next step's input while the gates are accumulated
         float xn[ST];
#pragma unroll
         for ( int st = 0; st < ST; ++st )
         {
            int s = s0 + st;
            xn[st] = ( s < nstreams && step + 1 < steps ) ? __ldg( x + ( (size_t)s * steps + step + 1 ) * 64 + j ) : 0.0f;
         }

         float zi[ST], zf[ST], zg[ST], zo[ST];
#pragma unroll
         for ( int st = 0; st < ST; ++st )
         {
            zi[st] = bi; zf[st] = bf; zg[st] = bg; zo[st] = bo;
         }
#pragma unroll 4
         for ( int kq = 0; kq < 32; ++kq )
         {
            const float *wq = Ws + ( kq * 256 + j ) * 4;
            float4 wi = ld4( wq ), wf = ld4( wq + 64 * 4 ), wg = ld4( wq + 128 * 4 ), wo = ld4( wq + 192 * 4 );
#pragma unroll
            for ( int st = 0; st < ST; ++st )
            {
               float4 v = ld4( cur + st * 128 + kq * 4 );
               zi[st] = fmaf( wi.x, v.x, zi[st] ); zi[st] = fmaf( wi.y, v.y, zi[st] );
               zi[st] = fmaf( wi.z, v.z, zi[st] ); zi[st] = fmaf( wi.w, v.w, zi[st] );
               zf[st] = fmaf( wf.x, v.x, zf[st] ); zf[st] = fmaf( wf.y, v.y, zf[st] );
               zf[st] = fmaf( wf.z, v.z, zf[st] ); zf[st] = fmaf( wf.w, v.w, zf[st] );
               zg[st] = fmaf( wg.x, v.x, zg[st] ); zg[st] = fmaf( wg.y, v.y, zg[st] );
               zg[st] = fmaf( wg.z, v.z, zg[st] ); zg[st] = fmaf( wg.w, v.w, zg[st] );
               zo[st] = fmaf( wo.x, v.x, zo[st] ); zo[st] = fmaf( wo.y, v.y, zo[st] );
               zo[st] = fmaf( wo.z, v.z, zo[st] ); zo[st] = fmaf( wo.w, v.w, zo[st] );
            }
         }
         // cell update (lstm.c:64-88)
#pragma unroll
         for ( int st = 0; st < ST; ++st )
         {
            float ig = sigmoid_acc( zi[st] ), fg = sigmoid_acc( zf[st] ), gg = lme::tanhf_ref( zg[st] ), og = sigmoid_acc( zo[st] );
            float cn = __fadd_rn( __fmul_rn( fg, c[st] ), __fmul_rn( ig, gg ) );
            c[st] = cn;
            float hn = __fmul_rn( lme::tanhf_ref( cn ), og );
            nxt[st * 128 + 64 + j] = hn;
            nxt[st * 128 + j] = xn[st];
            int s = s0 + st;
            if ( hseq && s < nstreams ) hseq[( (size_t)s * steps + step ) * 64 + j] = hn;
            if ( LAYER == 1 )
            {
               float r = fmaxf( hn, 0.0f );
               d0[st] = fmaf( dw0, r, d0[st] );
               d1[st] = fmaf( dw1, r, d1[st] );
            }
         }
         if ( LAYER == 1 && ( step % 7 ) == 6 )
         {
            // decoder: relu -> 1x1 conv 64->2 -> mean over the chunk's 7 frames -> sigmoid
#pragma unroll
            for ( int st = 0; st < ST; ++st )
            {
               float a = d0[st], b = d1[st];
#pragma unroll
               for ( int off = 16; off > 0; off >>= 1 )
               {
                  a += __shfl_xor_sync( 0xffffffffu, a, off );
                  b += __shfl_xor_sync( 0xffffffffu, b, off );
               }
               if ( ( j & 31 ) == 0 )
               {
                  decp[( ( grp * 2 + ( j >> 5 ) ) * 2 + 0 ) * ST + st] = a;
                  decp[( ( grp * 2 + ( j >> 5 ) ) * 2 + 1 ) * ST + st] = b;
               }
               d0[st] = d1[st] = 0.0f;
            }
         }
         bar_sync( 1 + grp, 64 );
         if ( LAYER == 1 && ( step % 7 ) == 6 && j < 2 * ST )
         {
            int st = j >> 1, head = j & 1, s = s0 + st;
            if ( s < nstreams )
            {
               float sum = decp[( ( grp * 2 + 0 ) * 2 + head ) * ST + st] + decp[( ( grp * 2 + 1 ) * 2 + head ) * ST + st];
               float mean = sum / 7.0f + dbs[head];
               float p = lme::sigmoid_ref( mean );
               long long n = out_off + step / 7;
               if ( out2 ) out2[( (long long)s * out_stride + n ) * 2 + head] = p;
               if ( probs && head == 1 ) probs[(long long)s * out_stride + n] = p;
            }
         }
      }
      // final state lives in the buffer the last step wrote
      {
         float *fin = ( steps & 1 ) ? buf1 : buf0;
#pragma unroll
         for ( int st = 0; st < ST; ++st )
         {
            int s = s0 + st;
            if ( s < nstreams )
            {
               state_c[( (size_t)s * 2 + LAYER ) * 64 + j] = c[st];
               state_h[( (size_t)s * 2 + LAYER ) * 64 + j] = fin[st * 128 + 64 + j];
            }
         }
      }
      bar_sync( 1 + grp, 64 ); // before the next tile reuses the buffers / decp
   }
}
