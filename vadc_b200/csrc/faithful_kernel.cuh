// vadc_b200/csrc/faithful_kernel.cuh -- the encoder and the decoder head in the reference's own rounding sequence.
//
// The fast kernels of this engine (layer_kernel.cuh, layer*_tc_kernel.cuh) evaluate every contraction as an FMA chain or
// on the tensor cores: within 1e-5 of the reference chunk by chunk, but over a long silence the decoder LSTM integrates
// one-ulp differences into its cell state ("long streams"). This file is the other end of the
// trade: every rounding step of the reference's C backend as compiled with -mavx2 -ffp-contract=off, so that small
// stream batches -- the way the reference itself is used -- get the reference's bits.
//
//   adaptive_audio_normalization_inplace  misc.c:48-121   (means as sequential sums and true divisions)
//   conv_block                            conv.c:761-814  (dw_conv_tensor :60-113 scalar tails of dotproduct_simd;
//                                                          pw_conv_tensor = conv_tensor "variant E" :532-589)
//   transformer_block                     transformer.c:160-234 (tensor_linear -> mymatmul -> dotproduct_simd,
//                                                          maths.h:123-158: 16 taps per step into eight lanes in
//                                                          _mm256_hadd_ps order, lanes left to right, scalar tail)
//   dual_head_attention                   transformer.c:13-153, softmax tensor.h:751-784, layer_norm misc.c:143-210
//   conv_tensor_out + batch_norm1d + relu conv.c:715-724 (stride 1: variant E; stride 2: generic path :597-709),
//                                         misc.c:221-258
//   decoder                               silero_v3.c:231-303, maths.h:352-400
//
// No multiply is ever fused with an add (__fmul_rn/__fadd_rn are never contracted), divisions and square roots are the
// IEEE ones, expf is glibc's restated bit for bit (libm_exact.cuh).
//
// The arithmetic lives in functions that also compile for the host (tests/faithful_host_check.cpp runs this very file
// serially -- one "thread", barriers as no-ops -- against the oracle, so the orchestration below is checked bit for bit
// without a GPU); the kernels at the bottom are the only device-only part.
//
// Mapping: one CTA per chunk walks the four layers with the chunk's activations in shared memory in the reference's
// [C][T] layout; each stage spreads its output elements over the CTA's threads, weights come from the container's tensors
// through L1/L2 (764 KB for the whole model; matrices that are read one output per lane also exist as transposed copies so
// that those loads coalesce). The LSTM runs one CTA per stream, the decoder head one thread per (chunk, head). This path serves
// up to ~128 streams at the speed of the fp32 fast kernels (the serial LSTM bounds small batches either way), not thousands.
#pragma once

#if defined( __CUDACC__ )
#include "libm_exact.cuh"
#define FQ __host__ __device__ __forceinline__
#else
#include <math.h>
#define FQ static inline
#endif

namespace fq
{

FQ float mul( float a, float b )
{
#ifdef __CUDA_ARCH__
   return __fmul_rn( a, b );
#else
   return a * b;
#endif
}
FQ float add( float a, float b )
{
#ifdef __CUDA_ARCH__
   return __fadd_rn( a, b );
#else
   return a + b;
#endif
}
FQ float sub( float a, float b )
{
#ifdef __CUDA_ARCH__
   return __fsub_rn( a, b );
#else
   return a - b;
#endif
}
FQ float quot( float a, float b )
{
#ifdef __CUDA_ARCH__
   return __fdiv_rn( a, b );
#else
   return a / b;
#endif
}
FQ float root( float a )
{
#ifdef __CUDA_ARCH__
   return __fsqrt_rn( a );
#else
   return sqrtf( a );
#endif
}
FQ float expo( float a )
{
#ifdef __CUDA_ARCH__
   return lme::expf_ref( a );
#else
   return expf( a );
#endif
}
#if !defined( __CUDACC__ ) && defined( FQ_HOST_BARRIER )
extern "C" void fq_host_barrier( void ); // tests/hostcheck/faithful_host_race.cpp: a real barrier between host threads
#endif
FQ void barrier()
{
#ifdef __CUDA_ARCH__
   __syncthreads();
#elif !defined( __CUDACC__ ) && defined( FQ_HOST_BARRIER )
   fq_host_barrier();
#endif
}
FQ float relu( float v ) { return v < 0.0f ? 0.0f : v; } // `if (v < 0) v = 0` keeps -0.0f, like the reference

struct Quad
{
   float x, y, z, w;
};
FQ Quad load4( const float *p )
{
#ifdef __CUDA_ARCH__
   const float4 v = *reinterpret_cast<const float4 *>( p );
   return Quad{ v.x, v.y, v.z, v.w };
#else
   return Quad{ p[0], p[1], p[2], p[3] };
#endif
}

// maths.h:123-158 dotproduct_simd over strided operands: a[i*sa] * b[i*sb]. ROW: `a` is a contiguous, 16-byte aligned row (sa == 1)
// and is fetched four taps per load (on the device: one LDS.128 that the whole warp shares as a broadcast).
template <bool ROW = false>
FQ float dot_simd( const float *a, int sa, const float *b, int sb, int n )
{
   float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f, r4 = 0.0f, r5 = 0.0f, r6 = 0.0f, r7 = 0.0f;
   const int blocks = ( n / 16 ) * 16;
   for ( int i = 0; i < blocks; i += 16 )
   {
      float p[16];
      if ( ROW )
      {
#pragma unroll
         for ( int q = 0; q < 4; ++q )
         {
            const Quad v = load4( a + i + 4 * q );
            p[4 * q] = mul( v.x, b[( i + 4 * q ) * sb] );
            p[4 * q + 1] = mul( v.y, b[( i + 4 * q + 1 ) * sb] );
            p[4 * q + 2] = mul( v.z, b[( i + 4 * q + 2 ) * sb] );
            p[4 * q + 3] = mul( v.w, b[( i + 4 * q + 3 ) * sb] );
         }
      }
      else
      {
#pragma unroll
         for ( int j = 0; j < 16; ++j ) p[j] = mul( a[( i + j ) * sa], b[( i + j ) * sb] );
      }
      // _mm256_hadd_ps(p[0..7], p[8..15])
      r0 = add( r0, add( p[0], p[1] ) );
      r1 = add( r1, add( p[2], p[3] ) );
      r2 = add( r2, add( p[8], p[9] ) );
      r3 = add( r3, add( p[10], p[11] ) );
      r4 = add( r4, add( p[4], p[5] ) );
      r5 = add( r5, add( p[6], p[7] ) );
      r6 = add( r6, add( p[12], p[13] ) );
      r7 = add( r7, add( p[14], p[15] ) );
   }
   float result = 0.0f;
   result = add( result, r0 );
   result = add( result, r1 );
   result = add( result, r2 );
   result = add( result, r3 );
   result = add( result, r4 );
   result = add( result, r5 );
   result = add( result, r6 );
   result = add( result, r7 );
   for ( int i = blocks; i < n; ++i ) result = add( result, mul( a[i * sa], b[i * sb] ) );
   return result;
}

// conv.c:532-589 ("variant E", kernel size 1, hop 1): output (filter row `w`, position i) over `cin` channels of an
// input addressed as in[c*sc + i*st], filter taps w[c*sw]; the accumulator starts at the zero conv_tensor cleared the output with
FQ float conv1_e( const float *in, int sc, int st, int i, const float *w, int sw, int cin, float bias )
{
   float r1[8], r2[8];
#pragma unroll
   for ( int l = 0; l < 8; ++l ) r1[l] = r2[l] = 0.0f;
   const float *col = in + i * st;
   int j = 0;
   for ( ; j < cin - 15; j += 16 )
   {
#pragma unroll
      for ( int l = 0; l < 8; ++l )
      {
         r1[l] = add( r1[l], mul( col[( j + l ) * sc], w[( j + l ) * sw] ) );
         r2[l] = add( r2[l], mul( col[( j + 8 + l ) * sc], w[( j + 8 + l ) * sw] ) );
      }
   }
   const float h0 = add( r1[0], r1[1] ), h1 = add( r1[2], r1[3] ), h2 = add( r2[0], r2[1] ), h3 = add( r2[2], r2[3] );
   const float h4 = add( r1[4], r1[5] ), h5 = add( r1[6], r1[7] ), h6 = add( r2[4], r2[5] ), h7 = add( r2[6], r2[7] );
   const float q0 = add( add( h0, h1 ), add( h2, h3 ) ), q4 = add( add( h4, h5 ), add( h6, h7 ) );
   float o = 0.0f;
   o = add( o, add( q0, q4 ) );
   for ( ; j < cin; ++j ) o = add( o, mul( col[j * sc], w[j * sw] ) );
   return add( o, bias );
}

// conv.c:597-709 (generic path, kernel size 1, hop `stride`): channel-outer accumulation, bias last
FQ float conv1_generic( const float *in, int sc, int st, int i_in, const float *w, int sw, int cin, float bias )
{
   float o = 0.0f;
   const float *col = in + i_in * st;
   for ( int c = 0; c < cin; ++c )
   {
      float d = 0.0f;
      d = add( d, mul( col[c * sc], w[c * sw] ) );
      o = add( o, d );
   }
   return add( o, bias );
}

// misc.c:143-210, one row of `n` features (stride 1), out may alias in
FQ void layer_norm_row( const float *x, int n, const float *w, const float *b, float *out )
{
   const float inv = quot( 1.0f, (float)n );
   float sum = 0.0f;
   for ( int i = 0; i < n; ++i ) sum = add( sum, x[i] );
   const float mean = mul( sum, inv );
   float vs = 0.0f;
   for ( int i = 0; i < n; ++i )
   {
      const float d = sub( x[i], mean );
      vs = add( vs, mul( d, d ) );
   }
   const float var = mul( vs, inv );
   const float sd = root( add( var, 1e-5f ) );
   const float rstd = quot( 1.0f, sd );
   const float mr = mul( mean, rstd );
   for ( int i = 0; i < n; ++i ) out[i] = add( mul( sub( mul( x[i], rstd ), mr ), w[i] ), b[i] );
}

struct LayerShape
{
   int first, cin, c, t, stride, proj;
};
// tensor.h:154-170 (positional binding of the container's tensors)
FQ LayerShape layer_shape( int l )
{
   const LayerShape s[4] = { { 1, 129, 16, 25, 2, 1 }, { 25, 16, 32, 13, 2, 1 }, { 49, 32, 32, 7, 1, 0 }, { 71, 32, 64, 7, 1, 1 } };
   return s[l];
}

// Matrices that are read with one OUTPUT per lane -- per layer: QKV, attention out-projection, FFN linear1 and linear2, the strided
// 1x1 conv -- are kept transposed ([in][out]) as well: a warp's 32 loads of tap k then fall into one 128-byte line instead of 32
// (ncu on the first version: 18 sectors per global load request, the L1 tag stage was the kernel's bound). The arithmetic and its
// order do not change, only the address of each tap.
enum
{
   N_TRANSPOSED = 20
};
FQ void transposed_slot( int slot, int *index, int *n_out, int *n_in )
{
   const LayerShape s = layer_shape( slot / 5 );
   const int rel[5] = { 0, 2, 6, 8, 12 }; // qkv_w, attn_out_w, linear1_w, linear2_w, conv_w relative to qkv_w (tensor.h:131-152)
   *index = s.first + 4 + 2 * s.proj + rel[slot % 5];
   *n_in = s.c;
   *n_out = ( slot % 5 == 0 ) ? 3 * s.c : s.c;
}

// the 99 tensors of silero_v31_16k.testtensor in container order + the transposed copies
struct Weights
{
   const float *t[99];
   const float *tt[N_TRANSPOSED];
};

// shared-memory plan of one chunk (floats)
enum
{
   SM_X = 0,            // layer input  [cin][T]                      <= 129*25
   SM_DW = 3232,        // relu(dw conv) [cin][T]
   SM_Y = 6464,         // conv_block output [C][T]                   <= 448
   SM_PR = SM_Y + 448,  // projection branch [C][T]
   SM_U = SM_PR + 448,  // token-major rows [T][C]: u -> u + attention
   SM_QKV = SM_U + 448, // [T][3C]                                    <= 7*192
   SM_A = SM_QKV + 1344, // [2][T][T]                                 <= 2*625
   SM_CAT = SM_A + 1252, // [T][C] heads side by side
   SM_N1 = SM_CAT + 448, // after the first layer norm, then + FFN
   SM_L1 = SM_N1 + 448,  // FFN hidden
   SM_N2 = SM_L1 + 448,  // after the second layer norm
   SM_MEAN = SM_N2 + 448, // per-frame means [25], padded means [31], normalization scalar
   SM_FLOATS = SM_MEAN + 64
};

// transformer_layer (transformer.c:237-295) on the chunk in sm[SM_X]; the result goes to sm[SM_X] as the next layer's
// [C][Tout] input, or, for the last layer, token-major [Tout][C] to `a4`
FQ void layer( const Weights &W, int l, float *sm, float *a4, int tid, int nt )
{
   const LayerShape s = layer_shape( l );
   const int cin = s.cin, C = s.c, T = s.t, d = C / 2;
   int wi = s.first;
   const float *dw_w = W.t[wi++], *dw_b = W.t[wi++], *pw_w = W.t[wi++], *pw_b = W.t[wi++];
   const float *proj_w = 0, *proj_b = 0;
   if ( s.proj )
   {
      proj_w = W.t[wi++];
      proj_b = W.t[wi++];
   }
   // (the matrices qkv_w, attn_out_w, linear1_w, linear2_w, conv_w at wi + 0, 2, 6, 8, 12 are read through their transposed copies)
   const float *qkv_b = W.t[wi + 1], *ao_b = W.t[wi + 3], *n1_w = W.t[wi + 4], *n1_b = W.t[wi + 5], *l1_b = W.t[wi + 7], *l2_b = W.t[wi + 9];
   const float *n2_w = W.t[wi + 10], *n2_b = W.t[wi + 11], *cv_b = W.t[wi + 13];
   wi += 14;
   const float *bn_w = W.t[wi++], *bn_b = W.t[wi++], *bn_mean = W.t[wi++], *bn_var = W.t[wi++];
   const float *qkv_wT = W.tt[l * 5], *ao_wT = W.tt[l * 5 + 1], *l1_wT = W.tt[l * 5 + 2], *l2_wT = W.tt[l * 5 + 3], *cv_wT = W.tt[l * 5 + 4];

   float *X = sm + SM_X, *DW = sm + SM_DW, *Y = sm + SM_Y, *U = sm + SM_U, *QKV = sm + SM_QKV, *A = sm + SM_A;
   float *CAT = sm + SM_CAT, *N1 = sm + SM_N1, *L1 = sm + SM_L1, *N2 = sm + SM_N2;

   // depthwise k=5, zero pad 2 (conv.c:17-53, 60-113): the taps that exist, left to right from 0, then bias + sum; ReLU
   for ( int e = tid; e < cin * T; e += nt )
   {
      const int c = e / T, i = e - c * T;
      const float *a = X + c * T, *k = dw_w + c * 5;
      const int k0 = i < 2 ? 2 - i : 0, k1 = i + 2 >= T ? T + 1 - i : 4;
      float r = 0.0f;
      for ( int kk = k0; kk <= k1; ++kk ) r = add( r, mul( a[i + kk - 2], k[kk] ) );
      DW[e] = relu( add( dw_b[c], r ) );
   }
   barrier();
   // pointwise conv of the depthwise branch, projection (or identity) of the block input, sum, ReLU (conv.c:761-814)
   for ( int e = tid; e < C * T; e += nt )
   {
      const int f = e / T, i = e - f * T;
      float y = conv1_e( DW, T, 1, i, pw_w + f * cin, 1, cin, pw_b[f] );
      if ( s.proj )
         y = add( y, conv1_e( X, T, 1, i, proj_w + f * cin, 1, cin, proj_b[f] ) );
      else
         y = add( y, X[e] );
      Y[e] = relu( y );
      U[i * C + f] = Y[e]; // transformer_block works on the transposed tensor (transformer.c:170-176)
   }
   barrier();
   // QKV = u qkv_w^T + b (tensor.h:675-723)
   for ( int e = tid; e < T * 3 * C; e += nt )
   {
      const int t = e / ( 3 * C ), o = e - t * 3 * C;
      QKV[e] = add( dot_simd<true>( U + t * C, 1, qkv_wT + o, 3 * C, C ), qkv_b[o] );
   }
   barrier();
   // A_h[tk][tq] = (k_h[tk] . q_h[tq]) * 1/sqrt(d) (transformer.c:101-117)
   {
      const float scale = quot( 1.0f, root( (float)d ) );
      for ( int e = tid; e < 2 * T * T; e += nt )
      {
         const int h = e / ( T * T ), r = e - h * T * T, tk = r / T, tq = r - tk * T;
         A[e] = mul( dot_simd<true>( QKV + tk * 3 * C + C + h * d, 1, QKV + tq * 3 * C + h * d, 1, d ), scale );
      }
   }
   barrier();
   // softmax over each row (tensor.h:751-784)
   for ( int r = tid; r < 2 * T; r += nt )
   {
      float *row = A + r * T;
      float mx = row[0];
      for ( int i = 0; i < T; ++i )
         if ( row[i] > mx ) mx = row[i];
      float sum = 0.0f;
      for ( int i = 0; i < T; ++i )
      {
         const float ev = expo( sub( row[i], mx ) );
         row[i] = ev;
         sum = add( sum, ev );
      }
      const float inv = quot( 1.0f, sum );
      for ( int i = 0; i < T; ++i ) row[i] = mul( row[i], inv );
   }
   barrier();
   // O_h = A_h V_h, heads side by side (transformer.c:122-143)
   for ( int e = tid; e < T * C; e += nt )
   {
      const int tk = e / C, col = e - tk * C, h = col / d, j = col - h * d;
      CAT[e] = dot_simd( A + ( h * T + tk ) * T, 1, QKV + 2 * C + h * d + j, 3 * C, T );
   }
   barrier();
   // out-projection + residual (transformer.c:178-190)
   for ( int e = tid; e < T * C; e += nt )
   {
      const int t = e / C, o = e - t * C;
      const float att = add( dot_simd<true>( CAT + t * C, 1, ao_wT + o, C, C ), ao_b[o] );
      U[e] = add( U[e], att );
   }
   barrier();
   for ( int t = tid; t < T; t += nt ) layer_norm_row( U + t * C, C, n1_w, n1_b, N1 + t * C );
   barrier();
   for ( int e = tid; e < T * C; e += nt )
   {
      const int t = e / C, o = e - t * C;
      L1[e] = relu( add( dot_simd<true>( N1 + t * C, 1, l1_wT + o, C, C ), l1_b[o] ) );
   }
   barrier();
   for ( int e = tid; e < T * C; e += nt )
   {
      const int t = e / C, o = e - t * C;
      const float f2 = add( dot_simd<true>( L1 + t * C, 1, l2_wT + o, C, C ), l2_b[o] );
      U[e] = add( N1[e], f2 ); // (U is free again)
   }
   barrier();
   for ( int t = tid; t < T; t += nt ) layer_norm_row( U + t * C, C, n2_w, n2_b, N2 + t * C );
   barrier();
   // conv 1x1 with the layer's stride on the [C][T] view of N2, batch norm (eval), ReLU
   const int Tout = 1 + ( T - 1 ) / s.stride;
   for ( int e = tid; e < C * Tout; e += nt )
   {
      const int i = e / C, f = e - i * C; // lanes differ in the filter: transposed taps, the input position is a broadcast
      const float z = s.stride == 1 ? conv1_e( N2, 1, C, i, cv_wT + f, C, C, cv_b[f] )
                                    : conv1_generic( N2, 1, C, i * s.stride, cv_wT + f, C, C, cv_b[f] );
      const float sd = root( add( bn_var[f], 1e-5f ) );
      const float nv = quot( sub( z, bn_mean[f] ), sd );
      const float v = relu( add( mul( nv, bn_w[f] ), bn_b[f] ) );
      if ( l == 3 )
         a4[i * C + f] = v;
      else
         X[f * Tout + i] = v;
   }
   barrier();
}

// one chunk: log spectrogram [129][25] (log1pf(mag * 2^20), stft_kernel.cuh) -> a4 [7][64]
FQ void encoder_chunk( const Weights &W, const float *logspec, float *a4, float *sm, int tid, int nt )
{
   float *X = sm + SM_X, *M = sm + SM_MEAN;
   // misc.c:48-62: per-frame mean over the bins, sequential sum, division
   for ( int t = tid; t < 25; t += nt )
   {
      float s = 0.0f;
      for ( int c = 0; c < 129; ++c ) s = add( s, logspec[c * 25 + t] );
      M[t] = quot( s, 129.0f );
   }
   barrier();
   // misc.c:64-82: reflect pad 3, 7-tap smoothing on the generic conv path (0 + left-to-right dot), mean of the 25
   if ( tid == 0 )
   {
      const float g[7] = { 0.03663284704089164733887f, 0.11128076165914535522461f, 0.21674531698226928710938f, 0.27068215608596801757812f,
                           0.21674531698226928710938f, 0.11128076165914535522461f, 0.03663284704089164733887f };
      float *P = M + 25;
      for ( int j = 0; j < 3; ++j ) P[j] = M[3 - j];
      for ( int i = 0; i < 25; ++i ) P[3 + i] = M[i];
      for ( int j = 0; j < 3; ++j ) P[28 + j] = M[23 - j];
      float total = 0.0f;
      for ( int t = 0; t < 25; ++t )
      {
         float r = 0.0f;
         for ( int k = 0; k < 7; ++k ) r = add( r, mul( P[t + k], g[k] ) );
         float v = 0.0f;
         v = add( v, r );
         total = add( total, v );
      }
      M[60] = quot( total, 25.0f );
   }
   barrier();
   {
      const float mm = M[60];
      for ( int e = tid; e < 129 * 25; e += nt ) X[e] = sub( logspec[e], mm );
   }
   barrier();
   for ( int l = 0; l < 4; ++l ) layer( W, l, sm, a4, tid, nt );
}

// One LSTM gate row over [x|h] (lstm.c:31-62: 128 taps through dotproduct_simd, maths.h:123-158; the bias is added by the caller).
// The row's taps are stored in quads, `kq_stride` floats apart (lstm_kernel.cuh keeps W as [k/4][row][4]); xh is 16-byte aligned.
FQ float gate_dot( const float *xh, const float *wrow, int kq_stride )
{
   float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f, r4 = 0.0f, r5 = 0.0f, r6 = 0.0f, r7 = 0.0f;
   for ( int b = 0; b < 8; ++b )
   {
      // taps 16b..16b+15 = quads 0..3; pair sums in _mm256_hadd_ps( p[0..7], p[8..15] ) order
      const Quad w0 = load4( wrow + ( b * 4 + 0 ) * kq_stride ), w1 = load4( wrow + ( b * 4 + 1 ) * kq_stride );
      const Quad w2 = load4( wrow + ( b * 4 + 2 ) * kq_stride ), w3 = load4( wrow + ( b * 4 + 3 ) * kq_stride );
      const Quad v0 = load4( xh + b * 16 ), v1 = load4( xh + b * 16 + 4 ), v2 = load4( xh + b * 16 + 8 ), v3 = load4( xh + b * 16 + 12 );
      r0 = add( r0, add( mul( v0.x, w0.x ), mul( v0.y, w0.y ) ) );
      r1 = add( r1, add( mul( v0.z, w0.z ), mul( v0.w, w0.w ) ) );
      r2 = add( r2, add( mul( v2.x, w2.x ), mul( v2.y, w2.y ) ) );
      r3 = add( r3, add( mul( v2.z, w2.z ), mul( v2.w, w2.w ) ) );
      r4 = add( r4, add( mul( v1.x, w1.x ), mul( v1.y, w1.y ) ) );
      r5 = add( r5, add( mul( v1.z, w1.z ), mul( v1.w, w1.w ) ) );
      r6 = add( r6, add( mul( v3.x, w3.x ), mul( v3.y, w3.y ) ) );
      r7 = add( r7, add( mul( v3.z, w3.z ), mul( v3.w, w3.w ) ) );
   }
   float res = 0.0f;
   res = add( res, r0 );
   res = add( res, r1 );
   res = add( res, r2 );
   res = add( res, r3 );
   res = add( res, r4 );
   res = add( res, r5 );
   res = add( res, r6 );
   res = add( res, r7 );
   return res;
}

// decoder (silero_v3.c:231-303) for one chunk and one head: hs = the top LSTM layer's 7 outputs [7][64]
FQ float decoder_head( const float *hs, const float *w /*[64]*/, float bias )
{
   float acc[7];
#pragma unroll
   for ( int t = 0; t < 7; ++t ) acc[t] = 0.0f;
   for ( int c = 0; c < 64; ++c )
   {
      const float kv = w[c];
#pragma unroll
      for ( int t = 0; t < 7; ++t ) acc[t] = add( acc[t], mul( kv, relu( hs[t * 64 + c] ) ) );
   }
   float s = 0.0f;
#pragma unroll
   for ( int t = 0; t < 7; ++t ) s = add( s, add( acc[t], bias ) );
   const float mean = quot( s, 7.0f );
   return quot( 1.0f, add( 1.0f, expo( -mean ) ) );
}

} // namespace fq

#if defined( __CUDACC__ )
#define FAITHFUL_THREADS 256
#define FAITHFUL_SMEM_BYTES ( fq::SM_FLOATS * 4 )

// spec: [nchunks][129][25] log spectrogram; a4: [nchunks][7][64]
__global__ void __launch_bounds__( FAITHFUL_THREADS, 4 ) faithful_encoder_kernel( const float *__restrict__ spec, float *__restrict__ a4, fq::Weights W, int nchunks )
{
   extern __shared__ __align__( 16 ) float fsm[];
   for ( int ci = blockIdx.x; ci < nchunks; ci += gridDim.x )
      fq::encoder_chunk( W, spec + (size_t)ci * ( 129 * 25 ), a4 + (size_t)ci * 448, fsm, threadIdx.x, FAITHFUL_THREADS );
}

// hs: top-layer LSTM outputs [S][nw*7][64] (stream-major); one thread per (stream, chunk, head)
__global__ void faithful_decoder_kernel( const float *__restrict__ hs, const float *__restrict__ dec_w, const float *__restrict__ dec_b, int nstreams, int nw,
                                         float *__restrict__ out2, float *__restrict__ probs, long long out_stride, long long out_off )
{
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if ( i >= (long long)nstreams * nw * 2 ) return;
   const int head = (int)( i & 1 );
   const long long ci = i >> 1;
   const long long s = ci / nw, n = ci - s * nw;
   const float p = fq::decoder_head( hs + ci * 448, dec_w + head * 64, dec_b[head] );
   if ( out2 ) out2[( s * out_stride + out_off + n ) * 2 + head] = p;
   if ( probs && head == 1 ) probs[s * out_stride + out_off + n] = p;
}
#endif
