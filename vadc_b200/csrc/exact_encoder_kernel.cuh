// vadc_b200/csrc/exact_encoder_kernel.cuh -- the encoder in the reference's rounding sequence, laid out for throughput.
//
// Same arithmetic as faithful_kernel.cuh (adaptive_audio_normalization_inplace misc.c:48-121, conv_block conv.c:761-814 with
// conv_tensor's "variant E" :532-589 and generic path :597-709, transformer_block transformer.c:160-234 over dotproduct_simd
// maths.h:123-158, dual_head_attention transformer.c:13-153, softmax tensor.h:751-784, layer_norm misc.c:143-210, batch_norm1d
// misc.c:221-258) -- every multiply and add separately rounded, in the reference's order -- but mapped the other way round:
//
//   faithful_encoder_kernel: one CTA per chunk, the threads share each stage's outputs, 13 barriers per layer, weights from L1/L2.
//   here: one THREAD per token. A thread walks its token through a layer exactly like the reference's C code walks it (the
//   token's vector in registers, the loops over output features rolled, the loop over the contraction index unrolled); the 32 lanes
//   of a warp are 32 tokens that read the same weight row from shared memory as broadcast LDS.128. No partial warps in narrow
//   stages, no shuffles, one barrier per layer and batch (attention is the only place where tokens of a chunk meet).
//
// Between the stages of a layer a token's vectors rest in a per-CTA scratch in global memory, feature-major [row][thread], so that
// every access of a warp is one 128-byte line (the rolled feature loops cannot index registers). It is reused for every batch of a
// persistent CTA and stays in L1/L2.
//
// Two kernels:
//   exact_front_kernel     log spectrogram [chunk][129][25] -> normalization scalar, depthwise conv, and the first layer's two K = 129
//                          pointwise contractions -> conv_block output [chunk][16][25]. Five chunks per CTA batch in shared memory;
//                          thread = (token, quarter of the 16 output features).
//   exact_layer_kernel<L>  L = 0: the rest of the first layer from the conv_block output; L = 1..3: layers 2..4 from the previous
//                          layer's output [chunk][cin][T]. Output [chunk][C][Tout], the last layer token-major [chunk][7][64].
#pragma once
#include "common.cuh"
#include "libm_exact.cuh"

namespace xe
{
__device__ __forceinline__ float mul( float a, float b ) { return __fmul_rn( a, b ); }
__device__ __forceinline__ float add( float a, float b ) { return __fadd_rn( a, b ); }
__device__ __forceinline__ float sub( float a, float b ) { return __fsub_rn( a, b ); }
__device__ __forceinline__ float quot( float a, float b ) { return __fdiv_rn( a, b ); }
__device__ __forceinline__ float root( float a ) { return __fsqrt_rn( a ); }
__device__ __forceinline__ float relu( float v ) { return v < 0.0f ? 0.0f : v; } // keeps -0.0f, like `if (v < 0) v = 0`

__host__ __device__ constexpr int al4( int x ) { return ( x + 3 ) & ~3; }

// four products a[0..3] * v.{x,y,z,w}: two FMUL2 (mul.rn.f32x2: each half an IEEE product) on register pairs; the additions that
// consume them stay scalar (common.cuh: a packed product feeding a packed add would be contracted by ptxas)
#ifndef XE_PACKED_MUL
#define XE_PACKED_MUL 1
#endif
__device__ __forceinline__ void mul2f( float a0, float a1, float b0, float b1, float &p0, float &p1 )
{
#if XE_PACKED_MUL
   unpk2( mul2( pk2( a0, a1 ), pk2( b0, b1 ) ), p0, p1 );
#else
   p0 = mul( a0, b0 ); p1 = mul( a1, b1 );
#endif
}
__device__ __forceinline__ void mul4( const float *a, const float4 v, float &p0, float &p1, float &p2, float &p3 )
{
#if XE_PACKED_MUL
   unpk2( mul2( pk2( a[0], a[1] ), pk2( v.x, v.y ) ), p0, p1 );
   unpk2( mul2( pk2( a[2], a[3] ), pk2( v.z, v.w ) ), p2, p3 );
#else
   p0 = mul( a[0], v.x ); p1 = mul( a[1], v.y ); p2 = mul( a[2], v.z ); p3 = mul( a[3], v.w );
#endif
}

// offsets (floats) of a layer's tensors inside the engine's copy of the container: consecutive tensors, each padded to 4 floats
// (tensor.h:114-152 order)
struct LayerOff
{
   int cin, C, T, stride, proj, Tout;
   int dw_w, dw_b, pw_w, pw_b, proj_w, proj_b, qkv_w, qkv_b, ao_w, ao_b, n1_w, n1_b, l1_w, l1_b, l2_w, l2_b, n2_w, n2_b, cv_w, cv_b, bn_w, bn_b, bn_mean, bn_var, total;
};
__host__ __device__ constexpr LayerOff layer_off( int l )
{
   const LayerDims d = layer_dims( l );
   LayerOff o{};
   o.cin = d.cin; o.C = d.c; o.T = d.t; o.stride = d.stride; o.proj = d.proj; o.Tout = 1 + ( d.t - 1 ) / d.stride;
   int p = 0;
   o.dw_w = p; p += al4( d.cin * 5 );
   o.dw_b = p; p += al4( d.cin );
   o.pw_w = p; p += al4( d.c * d.cin );
   o.pw_b = p; p += al4( d.c );
   o.proj_w = p; if ( d.proj ) p += al4( d.c * d.cin );
   o.proj_b = p; if ( d.proj ) p += al4( d.c );
   o.qkv_w = p; p += al4( 3 * d.c * d.c );
   o.qkv_b = p; p += al4( 3 * d.c );
   o.ao_w = p; p += al4( d.c * d.c );
   o.ao_b = p; p += al4( d.c );
   o.n1_w = p; p += al4( d.c );
   o.n1_b = p; p += al4( d.c );
   o.l1_w = p; p += al4( d.c * d.c );
   o.l1_b = p; p += al4( d.c );
   o.l2_w = p; p += al4( d.c * d.c );
   o.l2_b = p; p += al4( d.c );
   o.n2_w = p; p += al4( d.c );
   o.n2_b = p; p += al4( d.c );
   o.cv_w = p; p += al4( d.c * d.c );
   o.cv_b = p; p += al4( d.c );
   o.bn_w = p; p += al4( d.c );
   o.bn_b = p; p += al4( d.c );
   o.bn_mean = p; p += al4( d.c );
   o.bn_var = p; p += al4( d.c );
   o.total = p;
   return o;
}

// maths.h:123-158 over a register vector and a 16-byte aligned weight row in shared memory (K a multiple of 16, or below 16).
// The eight lane accumulators start at +0 in the reference; adding the first pair sum to +0 changes at most the sign of a zero, which
// the final left-to-right sum (it starts at +0 as well) erases again, so the first block's pair sums are taken as they are.
template <int K>
__device__ __forceinline__ float dot_simd_r( const float ( &a )[K], const float *__restrict__ w )
{
   constexpr int NB = K / 16;
   float r[8];
#pragma unroll
   for ( int b = 0; b < NB; ++b )
   {
      float p[16];
#pragma unroll
      for ( int q = 0; q < 4; ++q )
      {
         const float4 v = ld4( w + 16 * b + 4 * q );
         mul4( &a[16 * b + 4 * q], v, p[4 * q + 0], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3] );
      }
      // _mm256_hadd_ps( p[0..7], p[8..15] ): lanes 0 1 | 2 3 <- second vector | 4 5 | 6 7 <- second vector
      const float s0 = add( p[0], p[1] ), s1 = add( p[2], p[3] ), s2 = add( p[8], p[9] ), s3 = add( p[10], p[11] );
      const float s4 = add( p[4], p[5] ), s5 = add( p[6], p[7] ), s6 = add( p[12], p[13] ), s7 = add( p[14], p[15] );
      if ( b == 0 )
      {
         r[0] = s0; r[1] = s1; r[2] = s2; r[3] = s3; r[4] = s4; r[5] = s5; r[6] = s6; r[7] = s7;
      }
      else
      {
         r[0] = add( r[0], s0 ); r[1] = add( r[1], s1 ); r[2] = add( r[2], s2 ); r[3] = add( r[3], s3 );
         r[4] = add( r[4], s4 ); r[5] = add( r[5], s5 ); r[6] = add( r[6], s6 ); r[7] = add( r[7], s7 );
      }
   }
   float res = 0.0f;
   if ( NB > 0 )
   {
#pragma unroll
      for ( int j = 0; j < 8; ++j ) res = add( res, r[j] );
   }
#pragma unroll
   for ( int i = NB * 16; i < K; ++i ) res = add( res, mul( a[i], w[i] ) );
   return res;
}

// conv.c:532-589 ("variant E", kernel 1, hop 1) over a register vector of K = 16 or 32 channels and an aligned weight row + bias
template <int K>
__device__ __forceinline__ float conv1_e_r( const float ( &x )[K], const float *__restrict__ w, float bias )
{
   static_assert( K % 16 == 0, "variant E without a scalar tail" );
   float a[16];
#pragma unroll
   for ( int b = 0; b < K / 16; ++b )
#pragma unroll
      for ( int q = 0; q < 4; ++q )
      {
         const float4 v = ld4( w + 16 * b + 4 * q );
         float p0, p1, p2, p3;
         mul4( &x[16 * b + 4 * q], v, p0, p1, p2, p3 );
         if ( b == 0 )
         {
            a[4 * q + 0] = p0; a[4 * q + 1] = p1; a[4 * q + 2] = p2; a[4 * q + 3] = p3;
         }
         else
         {
            a[4 * q + 0] = add( a[4 * q + 0], p0 ); a[4 * q + 1] = add( a[4 * q + 1], p1 );
            a[4 * q + 2] = add( a[4 * q + 2], p2 ); a[4 * q + 3] = add( a[4 * q + 3], p3 );
         }
      }
   // r1 = a[0..7], r2 = a[8..15]: _mm256_hadd_ps( r1, r2 ) twice, then low + high half
   const float h0 = add( a[0], a[1] ), h1 = add( a[2], a[3] ), h2 = add( a[8], a[9] ), h3 = add( a[10], a[11] );
   const float h4 = add( a[4], a[5] ), h5 = add( a[6], a[7] ), h6 = add( a[12], a[13] ), h7 = add( a[14], a[15] );
   const float q0 = add( add( h0, h1 ), add( h2, h3 ) ), q4 = add( add( h4, h5 ), add( h6, h7 ) );
   const float o = add( 0.0f, add( q0, q4 ) );
   return add( o, bias );
}

// conv.c:597-709 (generic path, kernel 1): channel-outer accumulation, bias last
template <int K>
__device__ __forceinline__ float conv1_generic_r( const float ( &x )[K], const float *__restrict__ w, float bias )
{
   float o = 0.0f;
#pragma unroll
   for ( int q = 0; q < K / 4; ++q )
   {
      const float4 v = ld4( w + 4 * q );
      float p0, p1, p2, p3;
      mul4( &x[4 * q], v, p0, p1, p2, p3 );
      o = add( o, p0 );
      o = add( o, p1 );
      o = add( o, p2 );
      o = add( o, p3 );
   }
   return add( o, bias );
}

// misc.c:143-210 on a register vector; w, b in shared memory
template <int C>
__device__ __forceinline__ void layer_norm_r( float ( &x )[C], const float *__restrict__ w, const float *__restrict__ b )
{
   const float inv = quot( 1.0f, (float)C );
   float sum = 0.0f;
#pragma unroll
   for ( int i = 0; i < C; ++i ) sum = add( sum, x[i] );
   const float mean = mul( sum, inv );
   float vs = 0.0f;
#pragma unroll
   for ( int i = 0; i < C; ++i )
   {
      const float d = sub( x[i], mean );
      vs = add( vs, mul( d, d ) );
   }
   const float var = mul( vs, inv );
   const float rstd = quot( 1.0f, root( add( var, 1e-5f ) ) );
   const float mr = mul( mean, rstd );
#pragma unroll
   for ( int i = 0; i < C; ++i ) x[i] = add( mul( sub( mul( x[i], rstd ), mr ), w[i] ), b[i] );
}
// tensor.h:751-784 softmax_inplace_stable on one row in memory: running max, expf of the differences, sequential sum, one reciprocal,
// one multiply per element -- the arithmetic exact_layer_kernel applies to its attention rows (there the row lives in registers)
__device__ __forceinline__ void softmax_row( float *row, int n )
{
   float mx = row[0];
   for ( int i = 0; i < n; ++i )
      if ( row[i] > mx ) mx = row[i];
   float sum = 0.0f;
   for ( int i = 0; i < n; ++i )
   {
      row[i] = lme::expf_ref( sub( row[i], mx ) );
      sum = add( sum, row[i] );
   }
   const float inv = quot( 1.0f, sum );
   for ( int i = 0; i < n; ++i ) row[i] = mul( row[i], inv );
}
} // namespace xe

// parity tap for the reference's softmax fixture (test.c:900, 100 x 100): thread = row
__global__ void exact_softmax_rows_kernel( float *x, int rows, int cols )
{
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   if ( r < rows ) xe::softmax_row( x + (size_t)r * cols, cols );
}

// ------------------------------------------------------------------------------------------------------------------------------
// front: normalization scalar + depthwise conv + the first layer's pointwise / projection contractions (K = 129)
// ------------------------------------------------------------------------------------------------------------------------------
#define XF_THREADS 512
#define XF_G 5                                   // chunks per batch: 125 tokens x 4 feature quarters = 500 threads
#define XF_TILE 16128                            // XF_G * 129 * 25 = 16125 floats, padded to keep what follows 16-byte aligned
#define XF_W_FLOATS ( VB_BINS * 32 )              // [c][16 pw | 16 proj]
#define XF_SMEM_FLOATS ( 3 * XF_TILE + XF_W_FLOATS + al4c( VB_BINS * 5 ) + al4c( VB_BINS ) + 32 + XF_G * 32 * 2 + 32 )
__host__ __device__ constexpr int al4c( int x ) { return ( x + 3 ) & ~3; }
#define XF_SMEM_BYTES ( XF_SMEM_FLOATS * 4 )

// spec: [nchunks][129][25] log1p spectrogram (NORM: the normalization scalar is computed and subtracted here; !NORM: the input is
// already normalized -- parity taps); y1: [nchunks][16][25] conv_block output of the first layer; wl: the first layer's tensors
template <bool NORM>
__global__ void __launch_bounds__( XF_THREADS, 1 ) exact_front_kernel( const float *__restrict__ spec, float *__restrict__ y1, const float *__restrict__ wl, int nchunks )
{
   constexpr xe::LayerOff O = xe::layer_off( 0 );
   extern __shared__ __align__( 16 ) float fsm[];
   float *Xa = fsm;                      // [2 buffers][g][129][25] log spectrogram as the STFT wrote it (+ up to 3 floats of alignment shift)
   float *Ds = Xa + 2 * XF_TILE;         // [g][129][25] relu(depthwise conv)
   float *Wf = Ds + XF_TILE;             // [129][16 features]{pw, proj}
   float *dww = Wf + XF_W_FLOATS;        // [129][5]
   float *dwb = dww + al4c( VB_BINS * 5 ); // [129]
   float *pb = dwb + al4c( VB_BINS );    // [16 pw_b | 16 proj_b]
   float *Ms = pb + 32;                  // [g][32] per-frame means
   float *Ss = Ms + XF_G * 32;           // [g][32] smoothed means
   float *MU = Ss + XF_G * 32;           // [g]
   const int tid = threadIdx.x;
   for ( int i = tid; i < VB_BINS * 16; i += XF_THREADS )
   {
      const int f = i / VB_BINS, c = i - f * VB_BINS;
      // [c][feature f]{pointwise, projection}: a thread's LDS.128 delivers (pw, proj) of two features -- the partners of (d, x) in an FMUL2
      Wf[c * 32 + 2 * f] = wl[O.pw_w + i];
      Wf[c * 32 + 2 * f + 1] = wl[O.proj_w + i];
   }
   for ( int i = tid; i < VB_BINS * 5; i += XF_THREADS ) dww[i] = wl[O.dw_w + i];
   for ( int i = tid; i < VB_BINS; i += XF_THREADS ) dwb[i] = wl[O.dw_b + i];
   if ( tid < 16 )
   {
      pb[tid] = wl[O.pw_b + tid];
      pb[16 + tid] = wl[O.proj_b + tid];
   }
   const int fq = tid >> 7, m = tid & 127;       // feature quarter, token of the batch (m < 125)
   const int g = m / VB_FRAMES, t = m - g * VB_FRAMES;
   const bool tok = m < XF_G * VB_FRAMES;

   // The batch after the current one is on its way (cp.async) while this one is computed. A batch starts at float c0 * 3225 of the
   // spectrogram -- 4-byte aligned only -- so the tile lands shifted by (address / 4) % 4 floats: 16-byte copies for the aligned body,
   // 4-byte copies for the ragged ends, nothing outside the batch is read. Returns the shift.
   auto fetch = [&]( int c0, float *buf ) -> int {
      const float *src = spec + (size_t)c0 * ( VB_BINS * VB_FRAMES );
      const int n = min( XF_G, nchunks - c0 ) * ( VB_BINS * VB_FRAMES );
      const int shift = (int)( ( reinterpret_cast<size_t>( src ) >> 2 ) & 3 ), head = ( 4 - shift ) & 3;
      float *dst = buf + shift;
      const unsigned d0 = (unsigned)__cvta_generic_to_shared( dst );
      const int nv = ( n - head ) >> 2, tail0 = head + 4 * nv;
      if ( tid < head ) asm volatile( "cp.async.ca.shared.global [%0], [%1], 4;" ::"r"( d0 + 4u * tid ), "l"( src + tid ) : "memory" );
      for ( int i = tid; i < nv; i += XF_THREADS )
         asm volatile( "cp.async.cg.shared.global [%0], [%1], 16;" ::"r"( d0 + 4u * ( head + 4 * i ) ), "l"( src + head + 4 * i ) : "memory" );
      if ( tail0 + tid < n ) asm volatile( "cp.async.ca.shared.global [%0], [%1], 4;" ::"r"( d0 + 4u * ( tail0 + tid ) ), "l"( src + tail0 + tid ) : "memory" );
      return shift;
   };
   int c0 = blockIdx.x * XF_G, buf = 0, shift = 0;
   if ( c0 < nchunks ) shift = fetch( c0, Xa );
   for ( ; c0 < nchunks; c0 += gridDim.x * XF_G, buf ^= 1 )
   {
      const int ng = min( XF_G, nchunks - c0 );
      const float *Xs = Xa + buf * XF_TILE + shift;
      asm volatile( "cp.async.wait_all;" ::: "memory" );
      __syncthreads(); // this batch has landed; the previous one is done with the other tile and with Ds (and the weights are staged)
      {
         const int cn = c0 + gridDim.x * XF_G;
         if ( cn < nchunks ) shift = fetch( cn, Xa + ( buf ^ 1 ) * XF_TILE );
      }
      if ( NORM )
      {
      // misc.c:48-62: per-frame mean over the bins, sequential sum, division
      if ( fq == 0 && tok && g < ng )
      {
         float s = 0.0f;
         const float *col = Xs + g * ( VB_BINS * VB_FRAMES ) + t;
#pragma unroll 4
         for ( int c = 0; c < VB_BINS; ++c ) s = xe::add( s, col[c * VB_FRAMES] );
         Ms[g * 32 + t] = xe::quot( s, 129.0f );
      }
      __syncthreads();
      // misc.c:64-82: reflect pad 3, 7-tap smoothing on the generic conv path (0 + left-to-right dot)
      if ( fq == 0 && tok && g < ng )
      {
         const float gk[7] = { 0.03663284704089164733887f, 0.11128076165914535522461f, 0.21674531698226928710938f, 0.27068215608596801757812f,
                               0.21674531698226928710938f, 0.11128076165914535522461f, 0.03663284704089164733887f };
         float r = 0.0f;
#pragma unroll
         for ( int k = 0; k < 7; ++k )
         {
            const int j = t + k; // index into the padded row: j < 3 -> M[3 - j]; j < 28 -> M[j - 3]; else M[51 - j]
            const int src = j < 3 ? 3 - j : ( j < 28 ? j - 3 : 51 - j );
            r = xe::add( r, xe::mul( Ms[g * 32 + src], gk[k] ) );
         }
         Ss[g * 32 + t] = xe::add( 0.0f, r );
      }
      __syncthreads();
      if ( tid < ng )
      {
         float total = 0.0f;
         for ( int i = 0; i < VB_FRAMES; ++i ) total = xe::add( total, Ss[tid * 32 + i] );
         MU[tid] = xe::quot( total, 25.0f );
      }
      __syncthreads();
      }
      // The normalization scalar is subtracted where a value is used (misc.c:84-124 does it in place; the difference is the same
      // number wherever it is formed), which saves the pass over the tile.
      // depthwise k = 5, zero pad 2 (conv.c:17-53, 60-113): the taps that exist, left to right from 0, then bias + sum; ReLU.
      // task = (row g * 129 + c, five frames): nine inputs, five weights in registers; only the outer two taps of the first and the last
      // segment of a row can fall into the padding.
      {
         const int ntask = ng * VB_BINS * 5;
         for ( int task = tid; task < ntask; task += XF_THREADS )
         {
            const int r = task / 5, seg = task - 5 * r;
            const int gg = r / VB_BINS, c = r - gg * VB_BINS;
            const float mu = NORM ? MU[gg] : 0.0f;
            const float *a = Xs + 5 * task; // row r, frame 5 seg
            float x[9];
#pragma unroll
            for ( int q = 0; q < 9; ++q )
            {
               const bool valid = q < 2 ? seg != 0 : ( q > 6 ? seg != 4 : true );
               const float v = valid ? a[q - 2] : 0.0f;
               x[q] = NORM ? xe::sub( v, mu ) : v;
            }
            const float *k = dww + c * 5;
            const float k0 = k[0], k1 = k[1], k2 = k[2], k3 = k[3], k4 = k[4], bias = dwb[c];
            const float kw[5] = { k0, k1, k2, k3, k4 };
#pragma unroll
            for ( int j = 0; j < 5; ++j )
            {
               float rr = 0.0f;
#pragma unroll
               for ( int kk = 0; kk < 5; ++kk )
               {
                  const int q = j + kk;
                  const float nr = xe::add( rr, xe::mul( x[q], kw[kk] ) );
                  if ( q < 2 )
                     rr = seg != 0 ? nr : rr;
                  else if ( q > 6 )
                     rr = seg != 4 ? nr : rr;
                  else
                     rr = nr;
               }
               Ds[5 * task + j] = xe::relu( xe::add( bias, rr ) );
            }
         }
      }
      __syncthreads();
      // pointwise conv of the depthwise branch + projection of the block input (conv.c:761-814), both "variant E" over 129 channels:
      // sixteen accumulators a[m] = sum_b x[16 b + m] w[16 b + m] (each a chain over b), combined as
      //   ((a0+a1)+(a2+a3)) + ((a8+a9)+(a10+a11))  +  ((a4+a5)+(a6+a7)) + ((a12+a13)+(a14+a15)),   0 + that, + tap 128, + bias.
      // The accumulators are visited in the tree's leaf order, so that four partial sums per output are all that is ever live.
      if ( tok && g < ng )
      {
         const float *xc = Xs + g * ( VB_BINS * VB_FRAMES ) + t, *dc = Ds + g * ( VB_BINS * VB_FRAMES ) + t;
         const float *wq = Wf + 8 * fq;
         const float mu = NORM ? MU[g] : 0.0f;
         float st[8][5]; // [output: 0..3 pointwise, 4..7 projection][tree stack: four folded entries + the leaf just pushed]
         constexpr int leaf[16] = { 0, 1, 2, 3, 8, 9, 10, 11, 4, 5, 6, 7, 12, 13, 14, 15 };
#pragma unroll
         for ( int li = 0; li < 16; ++li )
         {
            const int mm = leaf[li];
            float acc[8];
#pragma unroll
            for ( int b = 0; b < 8; ++b )
            {
               const int c = 16 * b + mm;
               const float dv = dc[c * VB_FRAMES], xv = NORM ? xe::sub( xc[c * VB_FRAMES], mu ) : xc[c * VB_FRAMES];
               const float4 w01 = ld4( wq + c * 32 ), w23 = ld4( wq + c * 32 + 4 );
               float p[8];
               xe::mul2f( dv, xv, w01.x, w01.y, p[0], p[4] );
               xe::mul2f( dv, xv, w01.z, w01.w, p[1], p[5] );
               xe::mul2f( dv, xv, w23.x, w23.y, p[2], p[6] );
               xe::mul2f( dv, xv, w23.z, w23.w, p[3], p[7] );
#pragma unroll
               for ( int o = 0; o < 8; ++o ) acc[o] = b == 0 ? p[o] : xe::add( acc[o], p[o] );
            }
            // push the leaf, then fold while the count of leaves so far is even (a binary counter over the tree)
            int depth = __builtin_popcount( li ); // stack entries before this leaf
#pragma unroll
            for ( int o = 0; o < 8; ++o ) st[o][depth] = acc[o];
            int n = li + 1;
            while ( ( n & 1 ) == 0 )
            {
#pragma unroll
               for ( int o = 0; o < 8; ++o ) st[o][depth - 1] = xe::add( st[o][depth - 1], st[o][depth] );
               --depth;
               n >>= 1;
            }
         }
         const float dv = dc[128 * VB_FRAMES], xv = NORM ? xe::sub( xc[128 * VB_FRAMES], mu ) : xc[128 * VB_FRAMES];
         const float4 w01 = ld4( wq + 128 * 32 ), w23 = ld4( wq + 128 * 32 + 4 );
         float tp[8];
         xe::mul2f( dv, xv, w01.x, w01.y, tp[0], tp[4] );
         xe::mul2f( dv, xv, w01.z, w01.w, tp[1], tp[5] );
         xe::mul2f( dv, xv, w23.x, w23.y, tp[2], tp[6] );
         xe::mul2f( dv, xv, w23.z, w23.w, tp[3], tp[7] );
         float *dst = y1 + (size_t)( c0 + g ) * ( 16 * VB_FRAMES ) + t;
#pragma unroll
         for ( int i = 0; i < 4; ++i )
         {
            const float yp = xe::add( xe::add( xe::add( 0.0f, st[i][0] ), tp[i] ), pb[4 * fq + i] );
            const float yj = xe::add( xe::add( xe::add( 0.0f, st[4 + i][0] ), tp[4 + i] ), pb[16 + 4 * fq + i] );
            dst[( 4 * fq + i ) * VB_FRAMES] = xe::relu( xe::add( yp, yj ) );
         }
      }
   }
}

// ------------------------------------------------------------------------------------------------------------------------------
// one transformer_layer (transformer.c:237-295), thread = token
// ------------------------------------------------------------------------------------------------------------------------------
#define XE_THREADS 512
template <int L>
struct XeCfg
{
   static constexpr xe::LayerOff O = xe::layer_off( L );
   static constexpr int GB = XE_THREADS / O.T;                       // chunks per batch
   static constexpr int W_FLOATS = O.total;
   static constexpr int SMEM_BYTES = W_FLOATS * 4;
   static constexpr int ROWS = 5 * O.C;                              // scratch rows per CTA: R0 [C], R1 [C], QKV [3C]
   static constexpr size_t SCRATCH_FLOATS = (size_t)ROWS * XE_THREADS; // per CTA
};

// in: L == 0: conv_block output [chunk][16][25] (exact_front_kernel); else the previous layer's output [chunk][cin][T]
// out: [chunk][C][Tout]; L == 3: [chunk][7][64]
template <int L>
__global__ void __launch_bounds__( XE_THREADS, 1 )
exact_layer_kernel( const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ wl, float *scratch, int nchunks, int gb /* chunks per batch, <= GB */ )
{
   using Cfg = XeCfg<L>;
   constexpr xe::LayerOff O = Cfg::O;
   constexpr int CIN = O.cin, C = O.C, T = O.T, D = C / 2, TOUT = O.Tout;
   constexpr int UNR = C <= 16 ? 4 : 2; // output features in flight per thread: the narrow layer has the registers for four
   extern __shared__ __align__( 16 ) float wsm[];
   const int tid = threadIdx.x;
   {
      const float4 *src = reinterpret_cast<const float4 *>( wl );
      float4 *dst = reinterpret_cast<float4 *>( wsm );
      for ( int i = tid; i < O.total / 4; i += XE_THREADS ) dst[i] = __ldg( src + i );
   }
   float *R0 = scratch + (size_t)blockIdx.x * Cfg::SCRATCH_FLOATS + tid; // my column; row r at R0[r * XE_THREADS]
   float *R1 = R0 + (size_t)C * XE_THREADS;
   float *QKV = R1 + (size_t)C * XE_THREADS;
   const int g = tid / T, t = tid - g * T;
   const bool tok = g < gb;
   const float *QKVg = QKV - tid + g * T; // column of token 0 of my chunk
   __syncthreads();

   for ( int c0 = blockIdx.x * gb; c0 < nchunks; c0 += gridDim.x * gb )
   {
      const bool live = tok && c0 + g < nchunks;
      const int ci = c0 + g;
      if ( live )
      {
         float u[C];
         if constexpr ( L == 0 )
         {
#pragma unroll
            for ( int f = 0; f < C; ++f ) u[f] = in[( (size_t)ci * C + f ) * T + t];
#pragma unroll
            for ( int f = 0; f < C; ++f ) R0[f * XE_THREADS] = u[f];
         }
         else
         {
            // conv_block (conv.c:761-814): depthwise k = 5 + ReLU, pointwise, projection or identity, sum, ReLU
            float x[CIN], d[CIN];
            const float *xin = in + (size_t)ci * CIN * T;
#pragma unroll
            for ( int c = 0; c < CIN; ++c )
            {
               float xv[5];
#pragma unroll
               for ( int k = 0; k < 5; ++k )
               {
                  const int tt = t + k - 2;
                  xv[k] = ( tt >= 0 && tt < T ) ? xin[c * T + tt] : 0.0f;
               }
               x[c] = xv[2];
               const float *kw = wsm + O.dw_w + c * 5;
               float r = 0.0f;
#pragma unroll
               for ( int k = 0; k < 5; ++k )
               {
                  const int tt = t + k - 2;
                  if ( tt >= 0 && tt < T ) r = xe::add( r, xe::mul( xv[k], kw[k] ) );
               }
               d[c] = xe::relu( xe::add( wsm[O.dw_b + c], r ) );
            }
#pragma unroll( UNR )
            for ( int f = 0; f < C; ++f )
            {
               float y = xe::conv1_e_r<CIN>( d, wsm + O.pw_w + f * CIN, wsm[O.pw_b + f] );
               if ( O.proj )
                  y = xe::add( y, xe::conv1_e_r<CIN>( x, wsm + O.proj_w + f * CIN, wsm[O.proj_b + f] ) );
               else
                  y = xe::add( y, xin[f * T + t] );
               R0[f * XE_THREADS] = xe::relu( y );
            }
#pragma unroll
            for ( int f = 0; f < C; ++f ) u[f] = R0[f * XE_THREADS];
         }
         // QKV = u qkv_w^T + b (tensor.h:675-723)
#pragma unroll( UNR )
         for ( int o = 0; o < 3 * C; ++o ) QKV[o * XE_THREADS] = xe::add( xe::dot_simd_r<C>( u, wsm + O.qkv_w + o * C ), wsm[O.qkv_b + o] );
      }
      __syncthreads(); // the chunk's q and v rows are complete
      if ( live )
      {
         // dual_head_attention (transformer.c:13-153): A_h[tk][tq] = (k_h[tk] . q_h[tq]) / sqrt(d), softmax over tq, O_h = A_h V_h
         const float scale = xe::quot( 1.0f, xe::root( (float)D ) );
#pragma unroll 1
         for ( int h = 0; h < 2; ++h )
         {
            float kh[D];
#pragma unroll
            for ( int i = 0; i < D; ++i ) kh[i] = QKV[( C + h * D + i ) * XE_THREADS];
            float A[T];
#pragma unroll
            for ( int tq = 0; tq < T; ++tq )
            {
               // dotproduct_simd( k row, q row, d ): d = 8 is all scalar tail, d = 16 one block, d = 32 two
               float q[D];
#pragma unroll
               for ( int i = 0; i < D; ++i ) q[i] = QKVg[( h * D + i ) * XE_THREADS + tq];
               float r[8], res = 0.0f;
               if ( D >= 16 )
               {
#pragma unroll
                  for ( int b = 0; b < D / 16; ++b )
                  {
                     float p[16];
#pragma unroll
                     for ( int j = 0; j < 16; j += 2 ) xe::mul2f( kh[16 * b + j], kh[16 * b + j + 1], q[16 * b + j], q[16 * b + j + 1], p[j], p[j + 1] );
                     const float s[8] = { xe::add( p[0], p[1] ), xe::add( p[2], p[3] ), xe::add( p[8], p[9] ), xe::add( p[10], p[11] ),
                                          xe::add( p[4], p[5] ), xe::add( p[6], p[7] ), xe::add( p[12], p[13] ), xe::add( p[14], p[15] ) };
#pragma unroll
                     for ( int j = 0; j < 8; ++j ) r[j] = b == 0 ? s[j] : xe::add( r[j], s[j] );
                  }
#pragma unroll
                  for ( int j = 0; j < 8; ++j ) res = xe::add( res, r[j] );
               }
#pragma unroll
               for ( int i = ( D / 16 ) * 16; i < D; ++i ) res = xe::add( res, xe::mul( kh[i], q[i] ) );
               A[tq] = xe::mul( res, scale );
            }
            // softmax (tensor.h:751-784)
            float mx = A[0];
#pragma unroll
            for ( int i = 0; i < T; ++i )
               if ( A[i] > mx ) mx = A[i];
            float sum = 0.0f;
#pragma unroll
            for ( int i = 0; i < T; ++i )
            {
               A[i] = lme::expf_ref( xe::sub( A[i], mx ) );
               sum = xe::add( sum, A[i] );
            }
            const float inv = xe::quot( 1.0f, sum );
#pragma unroll
            for ( int i = 0; i < T; ++i ) A[i] = xe::mul( A[i], inv );
            // O_h[j] = dotproduct_simd( A row, V column j, T ): T = 25 one block + 9 tail taps, T = 13 / 7 all tail
#pragma unroll( UNR )
            for ( int j = 0; j < D; ++j )
            {
               const float *vcol = QKVg + (size_t)( 2 * C + h * D + j ) * XE_THREADS;
               float res = 0.0f;
               if ( T >= 16 )
               {
                  float p[16];
#pragma unroll
                  for ( int i = 0; i < 16; i += 2 ) xe::mul2f( A[i], A[i + 1], vcol[i], vcol[i + 1], p[i], p[i + 1] );
                  const float s[8] = { xe::add( p[0], p[1] ), xe::add( p[2], p[3] ), xe::add( p[8], p[9] ), xe::add( p[10], p[11] ),
                                       xe::add( p[4], p[5] ), xe::add( p[6], p[7] ), xe::add( p[12], p[13] ), xe::add( p[14], p[15] ) };
#pragma unroll
                  for ( int i = 0; i < 8; ++i ) res = xe::add( res, s[i] );
               }
#pragma unroll
               for ( int i = ( T / 16 ) * 16; i < T; ++i ) res = xe::add( res, xe::mul( A[i], vcol[i] ) );
               R1[( h * D + j ) * XE_THREADS] = res;
            }
         }
      }
      __syncthreads(); // every token of the batch is done with the q and v rows (the next batch overwrites them)
      if ( live )
      {
         float v[C];
         // out-projection + residual (transformer.c:178-190)
#pragma unroll
         for ( int i = 0; i < C; ++i ) v[i] = R1[i * XE_THREADS];
#pragma unroll( UNR )
         for ( int o = 0; o < C; ++o )
         {
            const float att = xe::add( xe::dot_simd_r<C>( v, wsm + O.ao_w + o * C ), wsm[O.ao_b + o] );
            R0[o * XE_THREADS] = xe::add( R0[o * XE_THREADS], att );
         }
#pragma unroll
         for ( int i = 0; i < C; ++i ) v[i] = R0[i * XE_THREADS];
         xe::layer_norm_r<C>( v, wsm + O.n1_w, wsm + O.n1_b );
#pragma unroll
         for ( int i = 0; i < C; ++i ) R1[i * XE_THREADS] = v[i];
         // feed-forward + residual
#pragma unroll( UNR )
         for ( int o = 0; o < C; ++o ) R0[o * XE_THREADS] = xe::relu( xe::add( xe::dot_simd_r<C>( v, wsm + O.l1_w + o * C ), wsm[O.l1_b + o] ) );
#pragma unroll
         for ( int i = 0; i < C; ++i ) v[i] = R0[i * XE_THREADS];
#pragma unroll( UNR )
         for ( int o = 0; o < C; ++o )
         {
            const float f2 = xe::add( xe::dot_simd_r<C>( v, wsm + O.l2_w + o * C ), wsm[O.l2_b + o] );
            R1[o * XE_THREADS] = xe::add( R1[o * XE_THREADS], f2 );
         }
#pragma unroll
         for ( int i = 0; i < C; ++i ) v[i] = R1[i * XE_THREADS];
         xe::layer_norm_r<C>( v, wsm + O.n2_w, wsm + O.n2_b );
         // conv 1x1 with the layer's stride on the [C][T] view, batch norm (eval), ReLU
         if ( t % O.stride == 0 )
         {
            const int to = t / O.stride;
#pragma unroll( UNR )
            for ( int f = 0; f < C; ++f )
            {
               const float z = O.stride == 1 ? xe::conv1_e_r<C>( v, wsm + O.cv_w + f * C, wsm[O.cv_b + f] )
                                             : xe::conv1_generic_r<C>( v, wsm + O.cv_w + f * C, wsm[O.cv_b + f] );
               const float sd = xe::root( xe::add( wsm[O.bn_var + f], 1e-5f ) );
               const float nv = xe::quot( xe::sub( z, wsm[O.bn_mean + f] ), sd );
               const float r = xe::relu( xe::add( xe::mul( nv, wsm[O.bn_w + f] ), wsm[O.bn_b + f] ) );
               if ( L == 3 )
                  out[( (size_t)ci * TOUT + to ) * C + f] = r;
               else
                  out[( (size_t)ci * C + f ) * TOUT + to] = r;
            }
         }
      }
   }
}

// Tried and measured (r02): the same layer with TWO tokens per thread (256 threads x 255 registers, every weight quad feeding both
// tokens: half the LDS traffic per multiply-add). Bit-identical, not faster (12.1 ms against 11.5 ms for the four layers of 131 072
// chunks): two warps per scheduler with two tokens each hide the weight-load latency no better than four warps with one, and the
// shared-memory pipe was never the bound (34 % busy). The one-token kernel stays.
