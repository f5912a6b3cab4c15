// vadc_b200/csrc/stft_fft8_kernel.cuh -- STFT + magnitude + log1p, hybrid rule of stft_hybrid_kernel.cuh, with the
// FFT done mostly in registers: 8 lanes per frame, 16 complex points per lane.
//
// Replaces my_stft (stft.c:15-229) + the log1p and per-frame means of adaptive_audio_normalization_inplace
// (misc.c:40-82). Same contract and same accuracy rule as stft_hybrid_kernel: every frame is
// transformed by a fp32 FFT; every bin whose magnitude is below k_rel * ||windowed frame||_2 is re-evaluated with
// the reference's own rounding sequence (stft.c:108-184) and is bit-identical to the reference.
//
// Why a second FFT kernel: the warp-per-frame kernel spends 5 of its 7 butterfly stages in warp shuffles
// (48 SHFL + select/twiddle work per frame; ncu: issue 65 %, LSU pipe 39 %). Here the 256-point real transform of
// a frame is a 128-point complex FFT z[n] = y[2n] + i y[2n+1] factored 128 = 16 x 8 (n = 8 n1 + n2, k = k1 + 16 k2):
//   step 1  lane n2 of the frame's 8 lanes: 16-point FFT over n1 entirely in registers (two radix-4 passes with
//           compile-time twiddles), then the twiddle W128^(n2 k1) from a shared-memory table;
//   step 2  one exchange through shared memory (16 STS.64 + 8 LDS.128 per lane, conflict-free by padding);
//   step 3  lane i owns the rows k1 = i and 16 - i (lane 0: rows 0 and 8): two 8-point FFTs in registers. The
//           real-input post-processing Y[k] = E[k] + W256^k O[k] pairs Z[k] with Z[128 - k], and 128 - (i + 16 j) =
//           (16 - i) + 16 (7 - j): both members of every pair are already in the same lane, no further exchange.
// A warp transforms 4 frames at a time; the 7 compute warps of a CTA own one chunk (25 of 28 frame slots live).
//
// The compute warps never synchronise with each other: an eighth warp does all the I/O (s16/f32 chunk -> padded fp32
// tile with its reflect images; finished log spectrogram tile -> global memory with 16-byte stores; the normalization
// scalar) and hands tiles over through mbarriers (input tile full/empty: ring of three; output tile full/empty: two; the output
// tile leaves as one cp.async.bulk shared -> global copy), so a warp that has to re-evaluate many bins exactly delays nobody
// until it is more than a chunk behind.
#pragma once
#include "common.cuh"
#include "stft_fft_common.cuh"
#include "tc_common.cuh"

#define F8_CWARPS 7                      // compute warps
#define F8_THREADS ( ( F8_CWARPS + 1 ) * 32 )
#define F8_SLOTS ( F8_CWARPS * 4 )
#define F8_EX_SLOTS ( VB_FRAMES + 1 )    // exchange slots: one per live frame + one shared by the three dead slots (their results are discarded)
#define F8_XS_DEPTH 3                   // input tiles in flight (ring)
#define F8_XS_FLOATS ( 1792 + 16 * 28 ) // padded chunk, 16 pad words after every 64 samples: frames of one warp start 80 words apart
#define F8_EX_ROW 20                    // exchange row stride in words (8 complex + 4 pad): LDS.128 phases hit distinct banks
#define F8_EX_SLOT ( 8 * F8_EX_ROW + 16 ) // 8 rows per pass (rows 0..7, then rows 8..15); slots of a half-warp 16 banks apart
#define F8_OS_FLOATS 3228               // 3225 + up to 3 floats of alignment offset (see the copy-out)
#define F8_SMEM_FLOATS ( F8_XS_DEPTH * F8_XS_FLOATS + F8_EX_SLOTS * F8_EX_SLOT + 2 * F8_OS_FLOATS + 2 * 32 + 256 + 256 + 128 + 32 )
#define F8_SMEM_BYTES ( F8_SMEM_FLOATS * 4 )

__device__ __forceinline__ int f8_xaddr( int p ) { return p + ( ( p >> 6 ) << 4 ); }

// reflect-padding images of sample m (0..1535) in the padded tile (tensor.h:942-953)
__device__ __forceinline__ void f8_put_images( float *xs, int m, float v )
{
   if ( m >= 1 && m <= 128 ) xs[f8_xaddr( 128 - m )] = v;
   if ( m >= 1407 && m <= 1534 ) xs[f8_xaddr( 3198 - m )] = v;
}

// one warp moves a chunk: 12 x (4 samples per lane)
template <bool F32>
struct F8Raw
{
   int4 v4[F32 ? 12 : 1];
   int2 v2[F32 ? 1 : 12];
};

template <bool F32>
__device__ __forceinline__ void f8_load_raw( F8Raw<F32> &raw, const void *chunk, int lane )
{
#pragma unroll
   for ( int k = 0; k < 12; ++k )
   {
      if ( F32 )
         raw.v4[k] = __ldg( (const int4 *)chunk + lane + 32 * k );
      else
         raw.v2[k] = __ldg( (const int2 *)chunk + lane + 32 * k );
   }
}

template <bool F32>
__device__ __forceinline__ void f8_store_x( float *xs, const F8Raw<F32> &raw, int lane )
{
#pragma unroll
   for ( int k = 0; k < 12; ++k )
   {
      float4 f;
      if ( F32 )
         f = *reinterpret_cast<const float4 *>( &raw.v4[k] );
      else
      {
         const short *h = reinterpret_cast<const short *>( &raw.v2[k] );
         f = make_float4( (float)h[0] * ( 1.0f / 32768.0f ), (float)h[1] * ( 1.0f / 32768.0f ), (float)h[2] * ( 1.0f / 32768.0f ),
                          (float)h[3] * ( 1.0f / 32768.0f ) ); // vadc.c:884,898
      }
      const int m = 4 * ( lane + 32 * k );
      st4( xs + f8_xaddr( 128 + m ), f ); // 4 consecutive samples never straddle a 64-sample block
      if ( k <= 1 || k >= 10 )
      {
         f8_put_images( xs, m, f.x );
         f8_put_images( xs, m + 1, f.y );
         f8_put_images( xs, m + 2, f.z );
         f8_put_images( xs, m + 3, f.w );
      }
   }
}

// The reference's 256-tap tree (stft.c:108-184) for the two basis rows of bin f (Re: row f, Im: row 129 + f) at frame t on the
// padded tile, evaluated by one warp (lane = l*4 + g owns the 8-tap leaf {64 g + l + 8 v}; the g- and l-combines are xor
// butterflies, the same tree because fp32 addition is commutative; see hyb_exact_row), then the magnitude (stft.c:194-213).
// Same value in every lane. Out of line on purpose: 17 inlined copies (one per magnitude register) made the kernel 126 KB
// of code and the instruction-cache misses showed up as the second largest stall reason.
__device__ __noinline__ float f8_exact_mag( const float *__restrict__ xs, const float *__restrict__ basis, int f, int t, int lane )
{
   const int l = lane >> 2, g = lane & 3;
   const float *xp = xs + 80 * ( t + g ) + l;
   const float *br = basis + (size_t)f * 256 + 64 * g + l;
   const float *bi = br + 129 * 256;
   float wr[8], wi[8], x[8];
#pragma unroll
   for ( int v = 0; v < 8; ++v )
   {
      wr[v] = __ldg( br + 8 * v );
      wi[v] = __ldg( bi + 8 * v );
   }
#pragma unroll
   for ( int v = 0; v < 8; ++v ) x[v] = xp[8 * v];
   float pr[8], pi[8];
#pragma unroll
   for ( int v = 0; v < 8; ++v )
   {
      pr[v] = __fmul_rn( x[v], wr[v] );
      pi[v] = __fmul_rn( x[v], wi[v] );
   }
   float re = __fadd_rn( __fadd_rn( __fadd_rn( pr[0], pr[1] ), __fadd_rn( pr[2], pr[3] ) ), __fadd_rn( __fadd_rn( pr[4], pr[5] ), __fadd_rn( pr[6], pr[7] ) ) );
   float im = __fadd_rn( __fadd_rn( __fadd_rn( pi[0], pi[1] ), __fadd_rn( pi[2], pi[3] ) ), __fadd_rn( __fadd_rn( pi[4], pi[5] ), __fadd_rn( pi[6], pi[7] ) ) );
#pragma unroll
   for ( int off = 1; off < 32; off <<= 1 )
   {
      const float o_re = __shfl_xor_sync( 0xffffffffu, re, off ), o_im = __shfl_xor_sync( 0xffffffffu, im, off );
      re = __fadd_rn( re, o_re );
      im = __fadd_rn( im, o_im );
   }
   return sqrtf( __fadd_rn( __fmul_rn( re, re ), __fmul_rn( im, im ) ) );
}

// in-place 4-point forward DFT (W4 = -i), natural order
__device__ __forceinline__ void f8_fft4( cpx &a0, cpx &a1, cpx &a2, cpx &a3 )
{
   const cpx t0{ a0.re + a2.re, a0.im + a2.im }, t1{ a0.re - a2.re, a0.im - a2.im };
   const cpx t2{ a1.re + a3.re, a1.im + a3.im }, t3{ a1.re - a3.re, a1.im - a3.im };
   a0 = cpx{ t0.re + t2.re, t0.im + t2.im };
   a2 = cpx{ t0.re - t2.re, t0.im - t2.im };
   a1 = cpx{ t1.re + t3.im, t1.im - t3.re };
   a3 = cpx{ t1.re - t3.im, t1.im + t3.re };
}

// multiply by the compile-time constant W16^M = exp(-2 pi i M / 16)
template <int M>
__device__ __forceinline__ cpx f8_mul_w16( cpx a )
{
   constexpr float C1 = 0.92387953251128673848f, S1 = 0.38268343236508978178f, H = 0.70710678118654752440f;
   if ( M == 0 ) return a;
   if ( M == 1 ) return cpx{ fmaf( a.re, C1, a.im * S1 ), fmaf( a.im, C1, -a.re * S1 ) };
   if ( M == 2 ) return cpx{ ( a.re + a.im ) * H, ( a.im - a.re ) * H };
   if ( M == 3 ) return cpx{ fmaf( a.re, S1, a.im * C1 ), fmaf( a.im, S1, -a.re * C1 ) };
   if ( M == 4 ) return cpx{ a.im, -a.re };
   if ( M == 6 ) return cpx{ ( a.im - a.re ) * H, -( a.re + a.im ) * H };
   if ( M == 9 ) return cpx{ -fmaf( a.re, C1, a.im * S1 ), -fmaf( a.im, C1, -a.re * S1 ) }; // W16^9 = -W16^1
   return a;
}

// in-place 16-point forward DFT; on return z[4 ka + kb] holds X[ka + 4 kb]
__device__ __forceinline__ void f8_fft16( cpx ( &z )[16] )
{
#pragma unroll
   for ( int b = 0; b < 4; ++b ) f8_fft4( z[b], z[4 + b], z[8 + b], z[12 + b] ); // z[4 ka + b] = Y[b][ka]
   z[5] = f8_mul_w16<1>( z[5] );
   z[6] = f8_mul_w16<2>( z[6] );
   z[7] = f8_mul_w16<3>( z[7] );
   z[9] = f8_mul_w16<2>( z[9] );
   z[10] = f8_mul_w16<4>( z[10] );
   z[11] = f8_mul_w16<6>( z[11] );
   z[13] = f8_mul_w16<3>( z[13] );
   z[14] = f8_mul_w16<6>( z[14] );
   z[15] = f8_mul_w16<9>( z[15] );
#pragma unroll
   for ( int ka = 0; ka < 4; ++ka ) f8_fft4( z[4 * ka], z[4 * ka + 1], z[4 * ka + 2], z[4 * ka + 3] );
}

// in-place 8-point forward DFT, natural order in and out
__device__ __forceinline__ void f8_fft8( cpx ( &b )[8] )
{
   constexpr float H = 0.70710678118654752440f;
   f8_fft4( b[0], b[2], b[4], b[6] ); // Y[0][ka] in b[0], b[2], b[4], b[6]
   f8_fft4( b[1], b[3], b[5], b[7] ); // Y[1][ka] in b[1], b[3], b[5], b[7]
   const cpx y1 = cpx{ ( b[3].re + b[3].im ) * H, ( b[3].im - b[3].re ) * H };  // * W8^1
   const cpx y2 = cpx{ b[5].im, -b[5].re };                                      // * W8^2
   const cpx y3 = cpx{ ( b[7].im - b[7].re ) * H, -( b[7].re + b[7].im ) * H }; // * W8^3
   const cpx e0 = b[0], e1 = b[2], e2 = b[4], e3 = b[6], y0 = b[1];
   b[0] = cpx{ e0.re + y0.re, e0.im + y0.im };
   b[4] = cpx{ e0.re - y0.re, e0.im - y0.im };
   b[1] = cpx{ e1.re + y1.re, e1.im + y1.im };
   b[5] = cpx{ e1.re - y1.re, e1.im - y1.im };
   b[2] = cpx{ e2.re + y2.re, e2.im + y2.im };
   b[6] = cpx{ e2.re - y2.re, e2.im - y2.im };
   b[3] = cpx{ e3.re + y3.re, e3.im + y3.im };
   b[7] = cpx{ e3.re - y3.re, e3.im - y3.im };
}

// bin of pair slot j for lane i: lanes 1..7 pair Z[i + 16 j] with Z[128 - i - 16 j]; lane 0 owns the self-paired rows 0 and 8:
// slots 0..3 = (8 + 16 j, 120 - 16 j), slots 4..6 = (16 (j - 3), 128 - 16 (j - 3)), slot 7 = (0, 128) i.e. DC and Nyquist
__device__ __forceinline__ int f8_bin_a( int i, int j ) { return i ? i + 16 * j : ( j < 4 ? 8 + 16 * j : ( j < 7 ? 16 * ( j - 3 ) : 0 ) ); }
// bin of magnitude register r of lane i: r < 8: ma[r] (bin_a), r < 16: mb[r - 8] (128 - bin_a), r = 16: m64
__device__ __forceinline__ int f8_bin_of( int i, int r )
{
   const int a = f8_bin_a( i, r & 7 );
   return r < 8 ? a : ( r < 16 ? 128 - a : 64 );
}

// basis: the reference's forward_basis_buffer [258][256] (row 0 is the periodic Hann window).
// k_rel: fix-up threshold relative to ||windowed frame||_2. out_mode 0: log1p(m*2^20); 1: m.
// RAW: emit magnitudes instead of log1p(m * 2^20) (parity tap). A template parameter: as a run-time flag every one of the 17 values
// of a pass paid a select.
template <bool F32, bool RAW>
__global__ void __launch_bounds__( F8_THREADS, 3 )
stft_fft8_kernel( const void *__restrict__ in, long long stream_stride, int nw, int nchunks, const float *__restrict__ basis,
                  float *__restrict__ spec, float *__restrict__ mu_out, float k_rel, unsigned long long *__restrict__ flagged )
{
   extern __shared__ __align__( 16 ) float smem[];
   float *Xs = smem;                                      // [F8_XS_DEPTH][F8_XS_FLOATS]: ring of input tiles
   float *Ex = Xs + F8_XS_DEPTH * F8_XS_FLOATS;           // [slot][8 rows][20]
   float *Os = Ex + F8_EX_SLOTS * F8_EX_SLOT;             // [2][129][25] (+ alignment offset)
   float *Ms = Os + 2 * F8_OS_FLOATS;                     // [2][25] per-frame mean of the log spectrum
   float *Win = Ms + 2 * 32;                              // 0.5 * Hann[256]
   float2 *Tw1 = reinterpret_cast<float2 *>( Win + 256 ); // [k1 16][i 8]: W128^(i k1)
   float2 *Twp = Tw1 + 128;                               // [j 8][i 8]: (cos, sin)(2 pi bin_a(i, j) / 256)
   uint64_t *bars = reinterpret_cast<uint64_t *>( Twp + 64 ); // x_full[3], x_empty[3], o_full[2], o_empty[2]
   uint64_t *x_full = bars, *x_empty = bars + F8_XS_DEPTH, *o_full = bars + 2 * F8_XS_DEPTH, *o_empty = bars + 2 * F8_XS_DEPTH + 2;

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   constexpr unsigned FULL = 0xffffffffu;

   // ---- tables, barriers ----------------------------------------------------------------------
   for ( int q = tid; q < 256; q += F8_THREADS ) Win[q] = 0.5f * __ldg( basis + q ); // 1/2 of the real-FFT post-processing folded in (exact)
   if ( tid < 128 )
   {
      float s, c;
      sincospif( -(float)( ( tid >> 3 ) * ( tid & 7 ) ) / 64.0f, &s, &c );
      Tw1[tid] = make_float2( c, s );
   }
   if ( tid < 64 )
   {
      float s, c;
      sincospif( (float)f8_bin_a( tid & 7, tid >> 3 ) / 128.0f, &s, &c );
      Twp[tid] = make_float2( c, s );
   }
   if ( tid == 0 )
   {
      for ( int b = 0; b < F8_XS_DEPTH; ++b )
      {
         tc::mbar_init( &x_full[b], 1 );
         tc::mbar_init( &x_empty[b], F8_CWARPS );
      }
      for ( int b = 0; b < 2; ++b )
      {
         tc::mbar_init( &o_full[b], F8_CWARPS );
         tc::mbar_init( &o_empty[b], 1 );
      }
      tc::mbar_fence_init();
   }
   __syncthreads();

   if ( warp == F8_CWARPS )
   {
      // =========================== I/O warp ===========================================================
      F8Raw<F32> raw;
      int ci = blockIdx.x;
      if ( ci >= nchunks ) return;
      // the ring starts with tiles 0 and 1 in place and tile 2 in registers
      f8_load_raw<F32>( raw, hyb_chunk_ptr<F32>( in, stream_stride, nw, ci ), lane );
      f8_store_x<F32>( Xs, raw, lane );
      __syncwarp();
      if ( lane == 0 ) tc::mbar_arrive( &x_full[0] );
      if ( ci + (int)gridDim.x < nchunks )
      {
         f8_load_raw<F32>( raw, hyb_chunk_ptr<F32>( in, stream_stride, nw, ci + gridDim.x ), lane );
         f8_store_x<F32>( Xs + F8_XS_FLOATS, raw, lane );
         __syncwarp();
         if ( lane == 0 ) tc::mbar_arrive( &x_full[1] );
      }
      if ( ci + 2 * (int)gridDim.x < nchunks ) f8_load_raw<F32>( raw, hyb_chunk_ptr<F32>( in, stream_stride, nw, ci + 2 * gridDim.x ), lane );
      // Per iteration: first the next input tile (the compute warps need it the moment they finish the current chunk), then
      // the output tile of the PREVIOUS chunk: both become available at the same instant (all compute warps done with chunk
      // it-1), and the copy-out has a whole chunk of slack while a late input tile stalls 7 warps.
      auto copy_out = [&]( int co, int ito ) {
         const int b = ito & 1;
         tc::mbar_wait( &o_full[b], ( ito >> 1 ) & 1 );
         const float *os = Os + b * F8_OS_FLOATS;
         if ( mu_out )
         {
            // the scalar of adaptive_audio_normalization_inplace (misc.c:48-82): reflect-pad 3 + 7-tap smoothing of the
            // per-frame means, mean over the 25 frames (tree order; the frame means are warp reductions already)
            const float gk[7] = { 0.03663284704089164733887f, 0.11128076165914535522461f, 0.21674531698226928710938f, 0.27068215608596801757812f,
                                  0.21674531698226928710938f, 0.11128076165914535522461f, 0.03663284704089164733887f };
            const int tt = lane < VB_FRAMES ? lane : VB_FRAMES - 1;
            float v = 0.0f;
#pragma unroll
            for ( int k = 0; k < 7; ++k )
            {
               int idx = tt + k - 3;
               if ( idx < 0 ) idx = -idx;
               if ( idx >= VB_FRAMES ) idx = 2 * ( VB_FRAMES - 1 ) - idx;
               v = fmaf( Ms[b * 32 + idx], gk[k], v );
            }
            if ( lane >= VB_FRAMES ) v = 0.0f;
#pragma unroll
            for ( int off = 16; off > 0; off >>= 1 ) v += __shfl_xor_sync( FULL, v, off );
            if ( lane == 0 ) mu_out[co] = v / (float)VB_FRAMES;
         }
         // The tile sits at offset (co & 3) so that shared and global float indices agree modulo 4; its 16-byte-aligned body
         // (12.9 KB) leaves as ONE bulk asynchronous copy (cp.async.bulk shared -> global, the TMA engine), the <= 3 floats in
         // front of it and <= 3 behind it as scalar stores. With a loop of 26 x (LDS.128, STG.128) per lane this warp needed ~3 us
         // per chunk and the compute warps waited 14 % of their time for the output tile to be free again.
         const int off = co & 3;
         float *gq = spec + ( (size_t)co * HYB_OUT_FLOATS - off );
         const int s_begin = off ? 4 : 0, s_end = ( off + HYB_OUT_FLOATS ) & ~3;
         tc::fence_async_smem(); // the tile was written with generic-proxy stores
         if ( lane == 0 )
         {
            asm volatile( "cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"( gq + s_begin ), "r"( tc::smem_u32( os + s_begin ) ),
                          "r"( (uint32_t)( ( s_end - s_begin ) * 4 ) )
                          : "memory" );
            asm volatile( "cp.async.bulk.commit_group;" ::: "memory" );
         }
         else if ( lane < 8 )
         {
            const int sidx = lane < 4 ? off + ( lane - 1 ) : s_end + ( lane - 4 );
            const bool ok = lane < 4 ? sidx < s_begin : sidx < off + HYB_OUT_FLOATS;
            if ( ok ) gq[sidx] = os[sidx];
         }
         if ( lane == 0 ) asm volatile( "cp.async.bulk.wait_group.read 0;" ::: "memory" ); // the source may be overwritten from here on
         __syncwarp();
         if ( lane == 0 ) tc::mbar_arrive( &o_empty[b] );
      };
      int it = 0, cprev = -1;
      for ( ; ci < nchunks; ci += gridDim.x, ++it )
      {
         const int cn = ci + 2 * gridDim.x;
         if ( cn < nchunks )
         {
            // tile it+2 takes the place of tile it-1 as soon as the compute warps are done with that one: a warp can now be two
            // chunks ahead of the slowest one before it has to wait (with two tiles the warps a chunk ahead paid the latency of
            // this store on every chunk: 18 % of their time)
            const int j = it + 2, xb = j % F8_XS_DEPTH;
            tc::mbar_wait( &x_empty[xb], ( ( j / F8_XS_DEPTH ) & 1 ) ^ 1 );
            f8_store_x<F32>( Xs + xb * F8_XS_FLOATS, raw, lane );
            __syncwarp();
            if ( lane == 0 ) tc::mbar_arrive( &x_full[xb] );
            if ( cn + (int)gridDim.x < nchunks ) f8_load_raw<F32>( raw, hyb_chunk_ptr<F32>( in, stream_stride, nw, cn + gridDim.x ), lane );
         }
         if ( cprev >= 0 ) copy_out( cprev, it - 1 );
         cprev = ci;
      }
      copy_out( cprev, it - 1 );
      return;
   }

   // =========================== compute warps =========================================================
   const int i = lane & 7, g = lane >> 3;
   const int slot = warp * 4 + g;
   const bool live = slot < VB_FRAMES;
   const int t = live ? slot : VB_FRAMES - 1;
   unsigned nflag = 0;

   float *ex = Ex + ( live ? slot : VB_FRAMES ) * F8_EX_SLOT;
   const int ra = i, rb = i ? 8 - i : 0; // rows owned in step 3: ra of the first pass (k1 = i), rb of the second (k1 = 16 - i; lane 0: 8)
   const float tau_scale = 2.0f * k_rel;  // the window carries the factor 1/2
   const bool l0 = ( i == 0 );

   int it = 0;
   for ( int ci = blockIdx.x; ci < nchunks; ci += gridDim.x, ++it )
   {
      const int b = it & 1, xb = it % F8_XS_DEPTH;
      const float *xs = Xs + xb * F8_XS_FLOATS;
      float *os = Os + b * F8_OS_FLOATS + ( ci & 3 );
      tc::mbar_wait( &x_full[xb], ( it / F8_XS_DEPTH ) & 1 );

      // ---- step 1: windowed samples z[n1] = (y[16 n1 + 2 i], y[16 n1 + 2 i + 1]) / 2, 16-point FFT over n1 -----------
      cpx z[16];
      float e2 = 0.0f;
      {
         const float *xf = xs + 80 * t + 2 * i;
         const float *wf = Win + 2 * i;
#pragma unroll
         for ( int n1 = 0; n1 < 16; ++n1 )
         {
            const float2 v = *reinterpret_cast<const float2 *>( xf + 16 * n1 + 16 * ( n1 >> 2 ) );
            const float2 w = *reinterpret_cast<const float2 *>( wf + 16 * n1 );
            z[n1].re = v.x * w.x;
            z[n1].im = v.y * w.y;
            e2 = fmaf( z[n1].re, z[n1].re, e2 );
            e2 = fmaf( z[n1].im, z[n1].im, e2 );
         }
      }
      e2 += __shfl_xor_sync( FULL, e2, 1 );
      e2 += __shfl_xor_sync( FULL, e2, 2 );
      e2 += __shfl_xor_sync( FULL, e2, 4 );
      const float tau = tau_scale * sqrtf( e2 );

      f8_fft16( z );
      // twiddle W128^(i k1) and exchange in two passes (rows k1 = 0..7, then 8..15): row k1 of the slot, column i
      cpx za[8], zb[8];
#pragma unroll
      for ( int pass = 0; pass < 2; ++pass )
      {
         if ( pass ) __syncwarp(); // everybody has read the first pass
#pragma unroll
         for ( int kk = 0; kk < 8; ++kk )
         {
            const int k1 = 8 * pass + kk, ka = k1 & 3, kb = k1 >> 2;
            cpx a = z[4 * ka + kb];
            if ( k1 )
            {
               const float2 w = Tw1[k1 * 8 + i];
               a = cmul( a, cpx{ w.x, w.y } );
            }
            *reinterpret_cast<float2 *>( ex + kk * F8_EX_ROW + 2 * i ) = make_float2( a.re, a.im );
         }
         __syncwarp();
         const float4 *pr = reinterpret_cast<const float4 *>( ex + ( pass ? rb : ra ) * F8_EX_ROW );
#pragma unroll
         for ( int c = 0; c < 4; ++c )
         {
            const float4 v = pr[c];
            if ( pass )
            {
               zb[2 * c] = cpx{ v.x, v.y };
               zb[2 * c + 1] = cpx{ v.z, v.w };
            }
            else
            {
               za[2 * c] = cpx{ v.x, v.y };
               za[2 * c + 1] = cpx{ v.z, v.w };
            }
         }
      }
      __syncwarp(); // the slot is rewritten by this warp's next chunk

      // ---- step 3: 8-point FFTs over n2 -> Z[ra + 16 k2], Z[rb + 16 k2] -----------------------------------------------
      f8_fft8( za );
      f8_fft8( zb );

      // ---- real-input post-processing: pair slot j = (U[j], V[7 - j]) ----------------------------------------------
      // lanes 1..7: U = za, V = zb. Lane 0 (rows 0 and 8 pair with themselves): U = zb[0..3], za[1..3], za[0];
      // V[7 - j] = zb[7 - j] (j < 4), za[8 - (j - 3)] (j = 4..6), za[0] (j = 7)
      float ma[8], mb[8];
#pragma unroll
      for ( int j = 0; j < 8; ++j )
      {
         const cpx u0 = j < 4 ? zb[j] : ( j < 7 ? za[j - 3] : za[0] );
         const cpx v0 = j < 4 ? zb[7 - j] : ( j < 7 ? za[11 - j] : za[0] );
         const cpx a = l0 ? u0 : za[j];
         const cpx p = l0 ? v0 : zb[7 - j];
         const float2 w = Twp[j * 8 + i]; // (cos, sin)(2 pi k / 256)
         const float er = a.re + p.re, ei = a.im - p.im, orr = a.re - p.re, oi = a.im + p.im;
         const float t1 = fmaf( w.y, orr, -w.x * oi ), t2 = fmaf( w.y, oi, w.x * orr );
         const float yr = er - t1, yi = ei - t2, yr2 = er + t1, yi2 = ei + t2;
         ma[j] = hyb_sqrt_fast( fmaf( yr, yr, yi * yi ) );
         mb[j] = hyb_sqrt_fast( fmaf( yr2, yr2, yi2 * yi2 ) );
      }
      float m64 = 2.0f * hyb_sqrt_fast( fmaf( za[4].re, za[4].re, za[4].im * za[4].im ) ); // lane 0 only: |Y[64]| = |Z[64]|

      // ---- which bins are small: bit r of fl <-> magnitude register r (ma[0..7], mb[0..7], m64). One vote for the whole pass
      //      (17 vote/branch sequences, one per register, were 14 % of the pass's instructions).
      unsigned fl = 0;
#pragma unroll
      for ( int r = 0; r < 8; ++r )
      {
         fl |= ( ma[r] < tau ? 1u : 0u ) << r;
         fl |= ( mb[r] < tau ? 1u : 0u ) << ( 8 + r );
      }
      fl |= ( l0 && m64 < tau ? 1u : 0u ) << 16;
      if ( !live ) fl = 0;

      // ---- log1p(m * 2^20) (misc.c:40-46) into the chunk's output tile; per-frame mean (misc.c:48-62) --------------
      tc::mbar_wait( &o_empty[b], ( ( it >> 1 ) & 1 ) ^ 1 );
      float fsum = 0.0f;
#pragma unroll
      for ( int j = 0; j < 8; ++j )
      {
         const int fa = f8_bin_a( i, j );
         const float la = RAW ? ma[j] : hyb_log1p_scaled( ma[j] );
         const float lb = RAW ? mb[j] : hyb_log1p_scaled( mb[j] );
         if ( live )
         {
            os[fa * VB_FRAMES + t] = la;
            os[( 128 - fa ) * VB_FRAMES + t] = lb;
         }
         fsum += la + lb;
      }
      if ( l0 )
      {
         const float lv = RAW ? m64 : hyb_log1p_scaled( m64 );
         if ( live ) os[64 * VB_FRAMES + t] = lv;
         fsum += lv;
      }

      // ---- exact re-evaluation of the small bins (whole warp per bin); the owning lane replaces the value in the tile and
      //      corrects its frame sum (the magnitude registers cannot be indexed by a run-time r) ----------------------------------
      for ( unsigned anym = __ballot_sync( FULL, fl != 0 ); anym; anym &= anym - 1 )
      {
         const int src = __ffs( anym ) - 1;
         const int st = warp * 4 + ( src >> 3 );
         for ( unsigned bits = __shfl_sync( FULL, fl, src ); bits; bits &= bits - 1 )
         {
            const int f = f8_bin_of( src & 7, __ffs( bits ) - 1 );
            const float exv = f8_exact_mag( xs, basis, f, st, lane );
            if ( lane == src )
            {
               float *po = os + f * VB_FRAMES + st;
               const float nv = RAW ? exv : hyb_log1p_scaled( exv );
               fsum += nv - *po;
               *po = nv;
            }
            ++nflag;
         }
      }

      fsum += __shfl_xor_sync( FULL, fsum, 1 );
      fsum += __shfl_xor_sync( FULL, fsum, 2 );
      fsum += __shfl_xor_sync( FULL, fsum, 4 );
      if ( l0 && live ) Ms[b * 32 + t] = fsum / (float)VB_BINS;

      __syncwarp();
      if ( lane == 0 )
      {
         tc::mbar_arrive( &x_empty[xb] );
         tc::mbar_arrive( &o_full[b] );
      }
   }
   if ( flagged && lane == 0 && nflag ) atomicAdd( flagged, (unsigned long long)nflag );
}
