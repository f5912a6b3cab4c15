/* oracle/silero_oracle.c -- TEST INFRASTRUCTURE ONLY. See silero_oracle.h.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * Summation orders (including the AVX2 lane/tree orders of the reference's fast paths) are kept, so
 * that with -ffp-contract=off and the same libm the results are bit-identical to the reference's
 * own C backend built with the pinned flags. Written from the algorithm, not copied: plain loops,
 * no arena, no TestTensor.
 */
#include "silero_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* .testtensor container: int32 version(=1), int32 count; count x {int32 len; char name[len]};    */
/* count x {int32 ndim; int32 dims[ndim]; int32 size; int32 nbytes; float data[size]}             */
/* (tensor.h:201-253, utils.py:7-53)                                                              */
/* ------------------------------------------------------------------------------------------- */
typedef struct so_tensor
{
   int ndim;
   int dims[8];
   int size;
   float *data;
} so_tensor;

struct so_model
{
   int count;
   so_tensor *t;
   unsigned char *blob;
};

static int rd_i32( const unsigned char *p, size_t n, size_t *off, int *v )
{
   if ( *off + 4 > n ) return -1;
   memcpy( v, p + *off, 4 );
   *off += 4;
   return 0;
}

so_model *so_model_load( const void *bytes, size_t nbytes )
{
   const unsigned char *p = bytes;
   size_t off = 0;
   int version = 0, count = 0;
   if ( rd_i32( p, nbytes, &off, &version ) || rd_i32( p, nbytes, &off, &count ) ) return 0;
   if ( version != 1 || count <= 0 || count > 4096 ) return 0;
   so_model *m = calloc( 1, sizeof( *m ) );
   m->count = count;
   m->t = calloc( (size_t)count, sizeof( so_tensor ) );
   m->blob = malloc( nbytes );
   memcpy( m->blob, bytes, nbytes );
   for ( int i = 0; i < count; ++i )
   {
      int len = 0;
      if ( rd_i32( p, nbytes, &off, &len ) || len < 0 || off + (size_t)len > nbytes ) goto fail;
      off += (size_t)len;
   }
   for ( int i = 0; i < count; ++i )
   {
      so_tensor *t = m->t + i;
      int nb = 0;
      if ( rd_i32( p, nbytes, &off, &t->ndim ) || t->ndim < 0 || t->ndim > 8 ) goto fail;
      for ( int d = 0; d < t->ndim; ++d )
         if ( rd_i32( p, nbytes, &off, &t->dims[d] ) ) goto fail;
      if ( rd_i32( p, nbytes, &off, &t->size ) || rd_i32( p, nbytes, &off, &nb ) ) goto fail;
      if ( nb != t->size * 4 || off + (size_t)nb > nbytes ) goto fail;
      t->data = (float *)(m->blob + off); /* offsets are 4-byte aligned only when names are; copy below */
      off += (size_t)nb;
   }
   if ( off != nbytes ) goto fail;
   /* names have arbitrary lengths, so tensor payloads may be misaligned: give each its own buffer */
   for ( int i = 0; i < count; ++i )
   {
      float *d = malloc( sizeof( float ) * (size_t)( m->t[i].size > 0 ? m->t[i].size : 1 ) );
      memcpy( d, m->t[i].data, sizeof( float ) * (size_t)m->t[i].size );
      m->t[i].data = d;
   }
   free( m->blob );
   m->blob = 0;
   return m;
fail:
   free( m->blob );
   free( m->t );
   free( m );
   return 0;
}

so_model *so_model_load_file( const char *path )
{
   FILE *f = fopen( path, "rb" );
   if ( !f ) return 0;
   fseek( f, 0, SEEK_END );
   long n = ftell( f );
   fseek( f, 0, SEEK_SET );
   unsigned char *b = malloc( (size_t)n );
   size_t got = fread( b, 1, (size_t)n, f );
   fclose( f );
   so_model *m = got == (size_t)n ? so_model_load( b, (size_t)n ) : 0;
   free( b );
   return m;
}

void so_model_free( so_model *m )
{
   if ( !m ) return;
   for ( int i = 0; i < m->count; ++i ) free( m->t[i].data );
   free( m->t );
   free( m );
}

int so_model_tensor_count( const so_model *m ) { return m->count; }

const float *so_model_tensor( const so_model *m, int index, int *ndim, int *dims )
{
   if ( index < 0 || index >= m->count ) return 0;
   if ( ndim ) *ndim = m->t[index].ndim;
   if ( dims )
      for ( int d = 0; d < m->t[index].ndim; ++d ) dims[d] = m->t[index].dims[d];
   return m->t[index].data;
}

/* ------------------------------------------------------------------------------------------- */
/* inner products, in the reference's orders                                                    */
/* ------------------------------------------------------------------------------------------- */

/* maths.h:250-263 dotproduct_slow: left-to-right */
static float dot_seq( const float *a, const float *b, int n )
{
   float r = 0.0f;
   for ( int i = 0; i < n; ++i )
   {
      float v = a[i] * b[i];
      r += v;
   }
   return r;
}

/* maths.h:123-158 dotproduct_simd: 16 taps per step; eight running lanes fed with pair sums in
   _mm256_hadd_ps order; lanes summed left-to-right; then a scalar tail */
static float dot_simd( const float *a, const float *b, int n )
{
   float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
   int blocks = ( n / 16 ) * 16;
   for ( int i = 0; i < blocks; i += 16 )
   {
      float p[16];
      for ( int j = 0; j < 16; ++j ) p[j] = a[i + j] * b[i + j];
      /* hadd(ab, cd) with ab = p[0..7], cd = p[8..15] */
      float h0 = p[0] + p[1], h1 = p[2] + p[3], h2 = p[8] + p[9], h3 = p[10] + p[11];
      float h4 = p[4] + p[5], h5 = p[6] + p[7], h6 = p[12] + p[13], h7 = p[14] + p[15];
      r[0] += h0; r[1] += h1; r[2] += h2; r[3] += h3;
      r[4] += h4; r[5] += h5; r[6] += h6; r[7] += h7;
   }
   float result = 0.0f;
   result = result + r[0] + r[1] + r[2] + r[3] + r[4] + r[5] + r[6] + r[7];
   for ( int i = blocks; i < n; ++i )
   {
      float v = a[i] * b[i];
      result += v;
   }
   return result;
}

/* ------------------------------------------------------------------------------------------- */
/* STFT                                                                                         */
/* ------------------------------------------------------------------------------------------- */

/* tensor.h:912-958: reflect without repeating the edge sample */
void so_reflect_pad( const float *x, int n, int pad_l, int pad_r, float *out )
{
   for ( int j = 0; j < pad_l; ++j ) out[j] = x[pad_l - j];
   memcpy( out + pad_l, x, sizeof( float ) * (size_t)n );
   for ( int j = 0; j < pad_r; ++j ) out[pad_l + n + j] = x[n - 2 - j];
}

/* stft.c:82-190: one 256-tap correlation in the AVX2 tree order:
   four 64-tap groups; per group, 8 vector products of 8 lanes combined ((0+1)+(2+3))+((4+5)+(6+7));
   groups combined (g0+g1)+(g2+g3); lanes combined ((0+1)+(2+3))+((4+5)+(6+7)) */
static float stft_dot256( const float *x, const float *k )
{
   float g[4][8];
   for ( int grp = 0; grp < 4; ++grp )
   {
      const float *xa = x + 64 * grp;
      const float *ka = k + 64 * grp;
      for ( int l = 0; l < 8; ++l )
      {
         float p0 = xa[l] * ka[l], p1 = xa[8 + l] * ka[8 + l];
         float p2 = xa[16 + l] * ka[16 + l], p3 = xa[24 + l] * ka[24 + l];
         float p4 = xa[32 + l] * ka[32 + l], p5 = xa[40 + l] * ka[40 + l];
         float p6 = xa[48 + l] * ka[48 + l], p7 = xa[56 + l] * ka[56 + l];
         float s01 = p0 + p1, s23 = p2 + p3, s45 = p4 + p5, s67 = p6 + p7;
         float s0123 = s01 + s23, s4567 = s45 + s67;
         g[grp][l] = s0123 + s4567;
      }
   }
   float lane[8];
   for ( int l = 0; l < 8; ++l )
   {
      float r01 = g[0][l] + g[1][l];
      float r23 = g[2][l] + g[3][l];
      lane[l] = r01 + r23;
   }
   float s01 = lane[0] + lane[1], s23 = lane[2] + lane[3], s45 = lane[4] + lane[5], s67 = lane[6] + lane[7];
   float s0123 = s01 + s23, s4567 = s45 + s67;
   return s0123 + s4567;
}

/* stft.c:15-229 (my_stft with hop 64, pad 128): [B,1536] -> magnitude [B,129,25] */
void so_stft( const so_model *m, const float *x, int batch, float *out )
{
   const float *basis = m->t[0].data; /* [258,1,256] */
   float xp[1792];
   float conv[258 * 25];
   for ( int b = 0; b < batch; ++b )
   {
      so_reflect_pad( x + (size_t)b * 1536, 1536, 128, 128, xp );
      for ( int f = 0; f < 258; ++f )
         for ( int t = 0; t < 25; ++t ) conv[f * 25 + t] = stft_dot256( xp + 64 * t, basis + 256 * f );
      /* stft.c:194-213 */
      float *o = out + (size_t)b * 129 * 25;
      for ( int i = 0; i < 129 * 25; ++i )
      {
         float re = conv[i];
         float im = conv[129 * 25 + i];
         o[i] = sqrtf( re * re + im * im );
      }
   }
}

/* ------------------------------------------------------------------------------------------- */
/* adaptive audio normalization (misc.c:1-124)                                                  */
/* ------------------------------------------------------------------------------------------- */
void so_adaptive_norm( float *x, int batch, int channels, int frames )
{
   static const float filter[7] = { 0.03663284704089164733887f, 0.11128076165914535522461f,
                                    0.21674531698226928710938f, 0.27068215608596801757812f,
                                    0.21674531698226928710938f, 0.11128076165914535522461f,
                                    0.03663284704089164733887f };
   const float million = (float)( 1024 * 1024 );
   float *mean = malloc( sizeof( float ) * (size_t)frames );
   float *padded = malloc( sizeof( float ) * (size_t)( frames + 6 ) );
   for ( int b = 0; b < batch; ++b )
   {
      float *xb = x + (size_t)b * channels * frames;
      for ( int i = 0; i < channels * frames; ++i ) xb[i] = log1pf( xb[i] * million );
      for ( int t = 0; t < frames; ++t )
      {
         float s = 0.0f;
         for ( int c = 0; c < channels; ++c ) s += xb[c * frames + t];
         mean[t] = s / channels;
      }
      so_reflect_pad( mean, frames, 3, 3, padded );
      float mean_sum = 0.0f;
      for ( int t = 0; t < frames; ++t )
      {
         /* generic conv_tensor path (conv.c:597-709): out starts at 0, += dotproduct_slow */
         float v = 0.0f;
         v += dot_seq( padded + t, filter, 7 );
         mean_sum += v;
      }
      float mm = mean_sum / frames;
      for ( int i = 0; i < channels * frames; ++i ) xb[i] = xb[i] - mm;
   }
   free( mean );
   free( padded );
}

/* ------------------------------------------------------------------------------------------- */
/* convolutions (conv.c)                                                                        */
/* ------------------------------------------------------------------------------------------- */

/* conv.c:17-53 + 60-113: depthwise k=5, zero pad 2. `dotproduct` is dotproduct_simd whose n<=5 calls
   are all scalar tail: 0 + p0 + p1 ... left to right; then bias + value */
void so_dw_conv( const float *in, int channels, int T, const float *w, const float *b, float *out )
{
   for ( int c = 0; c < channels; ++c )
   {
      const float *a = in + (size_t)c * T;
      const float *k = w + (size_t)c * 5;
      float *o = out + (size_t)c * T;
      float bias = b[c];
      o[0] = bias + dot_simd( a, k + 2, 3 );
      o[1] = bias + dot_simd( a, k + 1, 4 );
      for ( int i = 0; i < T - 4; ++i ) o[2 + i] = bias + dot_simd( a + i, k, 5 );
      const float *tail = a + T - 5;
      o[T - 2] = bias + dot_simd( tail + 1, k, 4 );
      o[T - 1] = bias + dot_simd( tail + 2, k, 3 );
   }
}

/* conv.c:170-189 + 532-589 ("variant E"): per output filter, products laid out [t][cin]; two 8-lane
   running sums over 16-channel blocks, hadd tree, scalar tail, then bias. out accumulates (+=). */
void so_pw_conv( const float *in, int cin, int T, const float *w, const float *b, int cout, float *out )
{
   float *temp = malloc( sizeof( float ) * (size_t)cin * (size_t)T );
   for ( int f = 0; f < cout; ++f )
   {
      float bias = b ? b[f] : 0.0f;
      for ( int c = 0; c < cin; ++c )
      {
         float kv = w[(size_t)f * cin + c];
         for ( int i = 0; i < T; ++i ) temp[(size_t)i * cin + c] = in[(size_t)c * T + i] * kv;
      }
      float *o = out + (size_t)f * T;
      for ( int i = 0; i < T; ++i )
      {
         const float *row = temp + (size_t)i * cin;
         float r1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
         int j = 0;
         for ( ; j < cin - 15; j += 16 )
            for ( int l = 0; l < 8; ++l )
            {
               r1[l] += row[j + l];
               r2[l] += row[j + 8 + l];
            }
         /* hadd(r1,r2) */
         float h0 = r1[0] + r1[1], h1 = r1[2] + r1[3], h2 = r2[0] + r2[1], h3 = r2[2] + r2[3];
         float h4 = r1[4] + r1[5], h5 = r1[6] + r1[7], h6 = r2[4] + r2[5], h7 = r2[6] + r2[7];
         /* hadd(h,h) twice: lane0 = (h0+h1)+(h2+h3), lane4 = (h4+h5)+(h6+h7) */
         float a0 = h0 + h1, a1 = h2 + h3, a4 = h4 + h5, a5 = h6 + h7;
         float q0 = a0 + a1, q4 = a4 + a5;
         o[i] += q0 + q4;
         for ( ; j < cin; ++j ) o[i] += row[j];
         o[i] += bias;
      }
   }
   free( temp );
}

/* conv.c:597-709 generic path with kernel_size 1: channel-outer accumulation, bias last */
void so_conv1x1_strided( const float *in, int cin, int T, const float *w, const float *b, int cout, int stride, float *out )
{
   int Tout = 1 + ( T - 1 ) / stride;
   for ( int i = 0; i < cout * Tout; ++i ) out[i] = 0.0f;
   for ( int c = 0; c < cin; ++c )
      for ( int f = 0; f < cout; ++f )
      {
         float kv = w[(size_t)f * cin + c];
         for ( int i = 0; i < Tout; ++i )
         {
            float d = 0.0f;
            d += in[(size_t)c * T + i * stride] * kv;
            out[(size_t)f * Tout + i] += d;
         }
      }
   if ( b )
      for ( int f = 0; f < cout; ++f )
         for ( int i = 0; i < Tout; ++i ) out[(size_t)f * Tout + i] += b[f];
}

/* conv_tensor dispatch (conv.c:170 vs 597): k=1 & hop=1 -> variant E, else generic */
static void conv1x1( const float *in, int cin, int T, const float *w, const float *b, int cout, int stride, float *out )
{
   if ( stride == 1 )
   {
      for ( int i = 0; i < cout * T; ++i ) out[i] = 0.0f;
      so_pw_conv( in, cin, T, w, b, cout, out );
   }
   else
      so_conv1x1_strided( in, cin, T, w, b, cout, stride, out );
}

/* conv.c:761-814 */
void so_conv_block( const float *in, int cin, int T, int has_proj,
                    const float *dw_w, const float *dw_b, const float *pw_w, const float *pw_b,
                    const float *proj_w, const float *proj_b, int cout, float *out )
{
   float *dw = malloc( sizeof( float ) * (size_t)cin * T );
   so_dw_conv( in, cin, T, dw_w, dw_b, dw );
   for ( int i = 0; i < cin * T; ++i )
      if ( dw[i] < 0.0f ) dw[i] = 0.0f;
   conv1x1( dw, cin, T, pw_w, pw_b, cout, 1, out );
   if ( has_proj )
   {
      float *pr = malloc( sizeof( float ) * (size_t)cout * T );
      conv1x1( in, cin, T, proj_w, proj_b, cout, 1, pr );
      for ( int i = 0; i < cout * T; ++i ) out[i] += pr[i];
      free( pr );
   }
   else
      for ( int i = 0; i < cout * T; ++i ) out[i] += in[i];
   for ( int i = 0; i < cout * T; ++i )
      if ( out[i] < 0.0f ) out[i] = 0.0f;
   free( dw );
}

/* ------------------------------------------------------------------------------------------- */
/* linear / softmax / norms                                                                     */
/* ------------------------------------------------------------------------------------------- */

/* tensor.h:675-723 -> mymatmul (maths.h:286-300) -> dotproduct_simd; bias added afterwards */
void so_linear( const float *in, int rows, int k, const float *w, const float *b, int n, float *out )
{
   for ( int r = 0; r < rows; ++r )
      for ( int o = 0; o < n; ++o ) out[(size_t)r * n + o] = dot_simd( in + (size_t)r * k, w + (size_t)o * k, k );
   if ( b )
      for ( int r = 0; r < rows; ++r )
         for ( int o = 0; o < n; ++o ) out[(size_t)r * n + o] += b[o];
}

/* tensor.h:751-784 */
void so_softmax_rows( float *x, int rows, int cols )
{
   float *e = malloc( sizeof( float ) * (size_t)cols );
   for ( int r = 0; r < rows; ++r )
   {
      float *row = x + (size_t)r * cols;
      float mx = row[0];
      for ( int i = 0; i < cols; ++i )
         if ( row[i] > mx ) mx = row[i];
      float sum = 0.0f;
      for ( int i = 0; i < cols; ++i )
      {
         e[i] = expf( row[i] - mx );
         sum += e[i];
      }
      float inv = 1.0f / sum;
      for ( int i = 0; i < cols; ++i ) row[i] = e[i] * inv;
   }
   free( e );
}

/* misc.c:143-210 */
void so_layer_norm( const float *in, int rows, int features, const float *w, const float *b, float *out )
{
   const float eps = 1e-5f;
   float inv_features = 1.0f / features;
   for ( int r = 0; r < rows; ++r )
   {
      const float *x = in + (size_t)r * features;
      float sum = 0.0f;
      for ( int i = 0; i < features; ++i ) sum += x[i];
      float mean = sum * inv_features;
      float vs = 0.0f;
      for ( int i = 0; i < features; ++i )
      {
         float d = x[i] - mean;
         vs += d * d;
      }
      float var = vs * inv_features;
      float sd = sqrtf( var + eps );
      float rstd = 1.0f / sd;
      float mean_over_sd = mean * rstd;
      for ( int i = 0; i < features; ++i ) out[(size_t)r * features + i] = ( x[i] * rstd - mean_over_sd ) * w[i] + b[i];
   }
}

/* misc.c:221-258 */
void so_batch_norm( const float *in, int batch, int channels, int T, const float *mean, const float *var,
                    const float *w, const float *b, float *out )
{
   const float eps = 1e-5f;
   for ( int n = 0; n < batch; ++n )
      for ( int c = 0; c < channels; ++c )
      {
         float sd = sqrtf( var[c] + eps );
         for ( int t = 0; t < T; ++t )
         {
            size_t idx = ( (size_t)n * channels + c ) * T + t;
            float nv = ( in[idx] - mean[c] ) / sd;
            out[idx] = nv * w[c] + b[c];
         }
      }
}

/* ------------------------------------------------------------------------------------------- */
/* attention / transformer                                                                      */
/* ------------------------------------------------------------------------------------------- */

/* transformer.c:13-153 for one batch item. in/out are [T,C]. Two heads of C/2.
   A_h = softmax_rows( (K_h Q_h^T) * 1/sqrt(d) ) -- rows are K positions; O_h = A_h V_h. */
void so_attention( const float *in, int T, int C, const float *qkv_w, const float *qkv_b,
                   const float *proj_w, const float *proj_b, float *out )
{
   int d = C / 2;
   float *qkv = malloc( sizeof( float ) * (size_t)T * 3 * C );
   so_linear( in, T, C, qkv_w, qkv_b, 3 * C, qkv );
   float *q = malloc( sizeof( float ) * (size_t)T * d );
   float *k = malloc( sizeof( float ) * (size_t)T * d );
   float *vT = malloc( sizeof( float ) * (size_t)T * d ); /* [d,T] like the reference's v1/v2 */
   float *a = malloc( sizeof( float ) * (size_t)T * T );
   float *o = malloc( sizeof( float ) * (size_t)T * d );
   float *cat = malloc( sizeof( float ) * (size_t)T * C );
   const float scale = 1.0f / sqrtf( (float)d );
   for ( int h = 0; h < 2; ++h )
   {
      for ( int t = 0; t < T; ++t )
         for ( int j = 0; j < d; ++j )
         {
            q[t * d + j] = qkv[(size_t)t * 3 * C + h * d + j];
            k[t * d + j] = qkv[(size_t)t * 3 * C + C + h * d + j];
            vT[j * T + t] = qkv[(size_t)t * 3 * C + 2 * C + h * d + j];
         }
      so_linear( k, T, d, q, 0, T, a ); /* a[tk][tq] = k[tk] . q[tq] */
      for ( int i = 0; i < T * T; ++i ) a[i] *= scale;
      so_softmax_rows( a, T, T );
      so_linear( a, T, T, vT, 0, d, o ); /* o[tk][j] = sum_tq a[tk][tq] v[tq][j] */
      for ( int t = 0; t < T; ++t )
         for ( int j = 0; j < d; ++j ) cat[(size_t)t * C + h * d + j] = o[t * d + j];
   }
   so_linear( cat, T, C, proj_w, proj_b, C, out );
   free( qkv ); free( q ); free( k ); free( vT ); free( a ); free( o ); free( cat );
}

/* transformer.c:160-234. w12 (fill_transformer_weights order, tensor.h:131-142):
   qkv_w, qkv_b, attn_proj_w, attn_proj_b, norm1_w, norm1_b, lin1_w, lin1_b, lin2_w, lin2_b, norm2_w, norm2_b */
void so_transformer_block( const float *in, int C, int T, const float *const *w, float *out )
{
   size_t n = (size_t)C * T;
   float *u = malloc( sizeof( float ) * n );
   float *att = malloc( sizeof( float ) * n );
   float *n1 = malloc( sizeof( float ) * n );
   float *l1 = malloc( sizeof( float ) * n );
   float *l2 = malloc( sizeof( float ) * n );
   float *n2 = malloc( sizeof( float ) * n );
   for ( int c = 0; c < C; ++c )
      for ( int t = 0; t < T; ++t ) u[(size_t)t * C + c] = in[(size_t)c * T + t];
   so_attention( u, T, C, w[0], w[1], w[2], w[3], att );
   for ( size_t i = 0; i < n; ++i ) u[i] += att[i];
   so_layer_norm( u, T, C, w[4], w[5], n1 );
   so_linear( n1, T, C, w[6], w[7], C, l1 );
   for ( size_t i = 0; i < n; ++i )
      if ( l1[i] < 0.0f ) l1[i] = 0.0f;
   so_linear( l1, T, C, w[8], w[9], C, l2 );
   for ( size_t i = 0; i < n; ++i ) n1[i] += l2[i];
   so_layer_norm( n1, T, C, w[10], w[11], n2 );
   for ( int c = 0; c < C; ++c )
      for ( int t = 0; t < T; ++t ) out[(size_t)c * T + t] = n2[(size_t)t * C + c];
   free( u ); free( att ); free( n1 ); free( l1 ); free( l2 ); free( n2 );
}

/* transformer.c:237-295 */
void so_transformer_layer_w( const float *in, int cin, int T, int cout, int stride, int has_proj,
                             const float *const *w, float *out )
{
   int i = 0;
   const float *dw_w = w[i++], *dw_b = w[i++], *pw_w = w[i++], *pw_b = w[i++];
   const float *proj_w = 0, *proj_b = 0;
   if ( has_proj )
   {
      proj_w = w[i++];
      proj_b = w[i++];
   }
   const float *const *tb = w + i;
   i += 12;
   const float *conv_w = w[i++], *conv_b = w[i++];
   const float *bn_w = w[i++], *bn_b = w[i++], *bn_mean = w[i++], *bn_var = w[i++];

   int Tout = 1 + ( T - 1 ) / stride;
   float *cb = malloc( sizeof( float ) * (size_t)cout * T );
   float *tbo = malloc( sizeof( float ) * (size_t)cout * T );
   float *cv = malloc( sizeof( float ) * (size_t)cout * Tout );
   so_conv_block( in, cin, T, has_proj, dw_w, dw_b, pw_w, pw_b, proj_w, proj_b, cout, cb );
   so_transformer_block( cb, cout, T, tb, tbo );
   conv1x1( tbo, cout, T, conv_w, conv_b, cout, stride, cv );
   so_batch_norm( cv, 1, cout, Tout, bn_mean, bn_var, bn_w, bn_b, out );
   for ( int j = 0; j < cout * Tout; ++j )
      if ( out[j] < 0.0f ) out[j] = 0.0f;
   free( cb ); free( tbo ); free( cv );
}

/* positional binding of the 94 encoder tensors (tensor.h:154-191) */
static const struct { int first, cin, cout, T, stride, has_proj; } LAYERS[4] = {
   { 1, 129, 16, 25, 2, 1 }, { 25, 16, 32, 13, 2, 1 }, { 49, 32, 32, 7, 1, 0 }, { 71, 32, 64, 7, 1, 1 } };

void so_transformer_layer( const so_model *m, int layer, const float *in, int batch, float *out )
{
   const float *w[24];
   int nw = LAYERS[layer].has_proj ? 24 : 22;
   for ( int i = 0; i < nw; ++i ) w[i] = m->t[LAYERS[layer].first + i].data;
   int cin = LAYERS[layer].cin, cout = LAYERS[layer].cout, T = LAYERS[layer].T, s = LAYERS[layer].stride;
   int Tout = 1 + ( T - 1 ) / s;
   for ( int b = 0; b < batch; ++b )
      so_transformer_layer_w( in + (size_t)b * cin * T, cin, T, cout, s, LAYERS[layer].has_proj, w,
                              out + (size_t)b * cout * Tout );
}

/* silero_v3.c:4-64 */
void so_encoder( const so_model *m, const float *in, int batch, float *out )
{
   float *l1 = malloc( sizeof( float ) * (size_t)batch * 16 * 13 );
   float *l2 = malloc( sizeof( float ) * (size_t)batch * 32 * 7 );
   float *l3 = malloc( sizeof( float ) * (size_t)batch * 32 * 7 );
   so_transformer_layer( m, 0, in, batch, l1 );
   so_transformer_layer( m, 1, l1, batch, l2 );
   so_transformer_layer( m, 2, l2, batch, l3 );
   so_transformer_layer( m, 3, l3, batch, out );
   free( l1 ); free( l2 ); free( l3 );
}

/* ------------------------------------------------------------------------------------------- */
/* LSTM (lstm.c:31-218) and decoder (silero_v3.c:231-303)                                       */
/* ------------------------------------------------------------------------------------------- */
static float sigmoidf_( float v ) { return 1.0f / ( 1.0f + expf( -v ) ); }

void so_lstm_seq( const float *x, int steps, int hidden, const float *h0, const float *c0,
                  const float *w, const float *b, int layers, float *out )
{
   int H = hidden;
   float *h = malloc( sizeof( float ) * (size_t)layers * H );
   float *c = malloc( sizeof( float ) * (size_t)layers * H );
   float *xh = malloc( sizeof( float ) * 2 * (size_t)H );
   float *z = malloc( sizeof( float ) * 4 * (size_t)H );
   memcpy( h, h0, sizeof( float ) * (size_t)layers * H );
   memcpy( c, c0, sizeof( float ) * (size_t)layers * H );
   for ( int s = 0; s < steps; ++s )
   {
      const float *input = x + (size_t)s * H;
      for ( int l = 0; l < layers; ++l )
      {
         const float *W = w + (size_t)l * ( 2 * H ) * ( 4 * H );
         const float *B = b + (size_t)l * 4 * H;
         memcpy( xh, input, sizeof( float ) * (size_t)H );
         memcpy( xh + H, h + (size_t)l * H, sizeof( float ) * (size_t)H );
         for ( int r = 0; r < 4 * H; ++r ) z[r] = dot_simd( xh, W + (size_t)r * 2 * H, 2 * H );
         for ( int r = 0; r < 4 * H; ++r ) z[r] += B[r];
         for ( int j = 0; j < H; ++j )
         {
            float ig = sigmoidf_( z[j] ), fg = sigmoidf_( z[H + j] ), gg = tanhf( z[2 * H + j] ), og = sigmoidf_( z[3 * H + j] );
            float cn = fg * c[(size_t)l * H + j] + ig * gg;
            c[(size_t)l * H + j] = cn;
            float hn = tanhf( cn );
            hn *= og;
            h[(size_t)l * H + j] = hn;
         }
         input = h + (size_t)l * H;
      }
      memcpy( out + (size_t)s * H, h + (size_t)( layers - 1 ) * H, sizeof( float ) * (size_t)H );
   }
   memcpy( out + (size_t)steps * H, h, sizeof( float ) * (size_t)layers * H );
   memcpy( out + (size_t)steps * H + (size_t)layers * H, c, sizeof( float ) * (size_t)layers * H );
   free( h ); free( c ); free( xh ); free( z );
}

/* silero_v3.c:231-303: relu -> 1x1 conv (maths.h:352-400 channel-outer muladd, then bias) -> mean -> sigmoid */
void so_decoder( const float *in, int batch, int channels, int T, const float *w, const float *b, int nout, float *out )
{
   float *acc = malloc( sizeof( float ) * (size_t)T );
   for ( int n = 0; n < batch; ++n )
      for ( int f = 0; f < nout; ++f )
      {
         for ( int t = 0; t < T; ++t ) acc[t] = 0.0f;
         for ( int c = 0; c < channels; ++c )
         {
            float kv = w[(size_t)f * channels + c];
            for ( int t = 0; t < T; ++t )
            {
               float v = in[( (size_t)n * channels + c ) * T + t];
               if ( v < 0.0f ) v = 0.0f;
               acc[t] += kv * v;
            }
         }
         for ( int t = 0; t < T; ++t ) acc[t] += b ? b[f] : 0.0f;
         float s = 0.0f;
         for ( int t = 0; t < T; ++t ) s += acc[t];
         float mean = s / (float)T;
         out[(size_t)n * nout + f] = 1.0f / ( 1.0f + expf( -mean ) );
      }
   free( acc );
}

/* ------------------------------------------------------------------------------------------- */
/* whole model (silero_v3.c:72-215)                                                             */
/* ------------------------------------------------------------------------------------------- */
void so_run_chunks_stages( const so_model *m, so_state *state, const float *samples, int batch,
                           float *stft_out, float *norm_out, float *l1o, float *l2o, float *l3o, float *l4o,
                           float *lstm_out, float *out )
{
   size_t B = (size_t)batch;
   float *spec = malloc( sizeof( float ) * B * 129 * 25 );
   float *l1 = malloc( sizeof( float ) * B * 16 * 13 );
   float *l2 = malloc( sizeof( float ) * B * 32 * 7 );
   float *l3 = malloc( sizeof( float ) * B * 32 * 7 );
   float *l4 = malloc( sizeof( float ) * B * 64 * 7 );
   float *l4t = malloc( sizeof( float ) * B * 7 * 64 );
   float *lo = malloc( sizeof( float ) * ( B * 7 * 64 + 4 * 64 ) );
   float *lot = malloc( sizeof( float ) * B * 64 * 7 );

   so_stft( m, samples, batch, spec );
   if ( stft_out ) memcpy( stft_out, spec, sizeof( float ) * B * 129 * 25 );
   so_adaptive_norm( spec, batch, 129, 25 );
   if ( norm_out ) memcpy( norm_out, spec, sizeof( float ) * B * 129 * 25 );
   so_transformer_layer( m, 0, spec, batch, l1 );
   so_transformer_layer( m, 1, l1, batch, l2 );
   so_transformer_layer( m, 2, l2, batch, l3 );
   so_transformer_layer( m, 3, l3, batch, l4 );
   if ( l1o ) memcpy( l1o, l1, sizeof( float ) * B * 16 * 13 );
   if ( l2o ) memcpy( l2o, l2, sizeof( float ) * B * 32 * 7 );
   if ( l3o ) memcpy( l3o, l3, sizeof( float ) * B * 32 * 7 );
   if ( l4o ) memcpy( l4o, l4, sizeof( float ) * B * 64 * 7 );

   for ( size_t n = 0; n < B; ++n )
      for ( int c = 0; c < 64; ++c )
         for ( int t = 0; t < 7; ++t ) l4t[( n * 7 + t ) * 64 + c] = l4[( n * 64 + c ) * 7 + t];
   /* the LSTM walks batch*7 steps carrying state across batch items (lstm.c:275-286, finding F6) */
   so_lstm_seq( l4t, batch * 7, 64, state->h, state->c, m->t[95].data, m->t[96].data, 2, lo );
   memcpy( state->h, lo + B * 7 * 64, sizeof( float ) * 128 );
   memcpy( state->c, lo + B * 7 * 64 + 128, sizeof( float ) * 128 );
   if ( lstm_out ) memcpy( lstm_out, lo, sizeof( float ) * B * 7 * 64 );
   for ( size_t n = 0; n < B; ++n )
      for ( int t = 0; t < 7; ++t )
         for ( int c = 0; c < 64; ++c ) lot[( n * 64 + c ) * 7 + t] = lo[( n * 7 + t ) * 64 + c];
   so_decoder( lot, batch, 64, 7, m->t[97].data, m->t[98].data, 2, out );

   free( spec ); free( l1 ); free( l2 ); free( l3 ); free( l4 ); free( l4t ); free( lo ); free( lot );
}

void so_run_chunks( const so_model *m, so_state *state, const float *samples, int batch, float *out )
{
   so_run_chunks_stages( m, state, samples, batch, 0, 0, 0, 0, 0, 0, 0, out );
}

void so_run_pcm( const so_model *m, so_state *state, const int16_t *pcm, long long nsamples, float *out )
{
   enum { BATCH = 96 }; /* vadc.c:797, 1116 */
   long long nchunks = nsamples / 1536;
   float *buf = malloc( sizeof( float ) * 1536 * BATCH );
   for ( long long c0 = 0; c0 < nchunks; c0 += BATCH )
   {
      int n = (int)( nchunks - c0 < BATCH ? nchunks - c0 : BATCH );
      for ( long long i = 0; i < (long long)n * 1536; ++i )
      {
         float v = pcm[c0 * 1536 + i];
         buf[i] = v / 32768.0f; /* vadc.c:884,898 */
      }
      so_run_chunks( m, state, buf, n, out + c0 * 2 );
   }
   free( buf );
}

/* ------------------------------------------------------------------------------------------- */
/* timestamps (vadc.c:165-299, 756-768, 846, 1005-1027)                                         */
/* ------------------------------------------------------------------------------------------- */
void so_segment_params_default( so_segment_params *p )
{
   p->min_silence_ms = 200.0f;
   p->min_speech_ms = 250.0f;
   p->threshold = 0.5f;
   p->neg_threshold_relative = 0.15f;
   p->speech_pad_ms = 30.0f;
   p->centiseconds = 0;
}

typedef struct seg_sink
{
   char *text;
   size_t cap, len;
   int *pairs;
   long long max_pairs, npairs;
   int centi;
   float pad_ms, spc;
} seg_sink;

/* vadc.c:223-260 */
static void seg_emit( seg_sink *s, int start, int end )
{
   if ( s->pairs && s->npairs < s->max_pairs )
   {
      s->pairs[2 * s->npairs] = start;
      s->pairs[2 * s->npairs + 1] = end;
   }
   s->npairs++;
   if ( !s->text ) return;
   const float pad_s = s->pad_ms / 1000.0f;
   float end_p = ( end * s->spc ) + pad_s;
   float start_p = ( start * s->spc ) - pad_s;
   if ( start_p < 0.0f ) start_p = 0.0f;
   char line[96];
   int n;
   if ( !s->centi )
      n = snprintf( line, sizeof( line ), "%.2f,%.2f\n", start_p, end_p );
   else
      n = snprintf( line, sizeof( line ), "%lld,%lld\n", (long long)( (double)start_p * 100.0 + 0.5 ),
                    (long long)( (double)end_p * 100.0 + 0.5 ) );
   for ( int i = 0; i < n && s->len + 1 < s->cap; ++i ) s->text[s->len++] = line[i];
   if ( s->cap ) s->text[s->len] = 0;
}

typedef struct seg_buf { int start, end, valid; } seg_buf;

/* vadc.c:262-299 */
static void seg_combine( seg_sink *s, seg_buf *buffered, int start, int end )
{
   const float pad_s = s->pad_ms / 1000.0f;
   float cur_start_p = ( start * s->spc ) - pad_s;
   if ( cur_start_p < 0.0f ) cur_start_p = 0.0f;
   if ( buffered->valid )
   {
      float buf_end_p = ( buffered->end * s->spc ) + pad_s;
      if ( buf_end_p >= cur_start_p )
         buffered->end = end;
      else
      {
         seg_emit( s, buffered->start, buffered->end );
         buffered->start = start;
         buffered->end = end;
      }
   }
   else
   {
      buffered->start = start;
      buffered->end = end;
      buffered->valid = 1;
   }
}

static void seg_run( const float *prob, long long nchunks, const so_segment_params *p, seg_sink *s )
{
   const int input_count = 1536;
   const float chunk_ms = input_count / (float)16000 * 1000.0f;            /* vadc.c:756 */
   int min_speech = (int)( p->min_speech_ms / chunk_ms + 0.5f );           /* vadc.c:758-762 */
   if ( min_speech < 1 ) min_speech = 1;
   int min_silence = (int)( p->min_silence_ms / chunk_ms + 0.5f );         /* vadc.c:764-768 */
   if ( min_silence < 1 ) min_silence = 1;
   const float thr = p->threshold;
   const float neg = p->threshold - p->neg_threshold_relative;             /* vadc.c:1244 */
   s->spc = (float)input_count / 16000;                                     /* vadc.c:846 */
   s->pad_ms = p->speech_pad_ms;
   s->centi = p->centiseconds;

   int temp_end = 0, cur_start = 0, triggered = 0;
   seg_buf buffered = { 0, 0, 0 };
   int g = 0;
   for ( long long i = 0; i < nchunks; ++i, ++g )
   {
      float pr = prob[i];
      /* vadc.c:165-221 */
      if ( pr >= thr && temp_end > 0 ) temp_end = 0;
      if ( !triggered )
      {
         if ( pr >= thr )
         {
            triggered = 1;
            cur_start = g;
         }
      }
      else if ( pr < neg )
      {
         if ( temp_end == 0 ) temp_end = g;
         if ( g - temp_end >= min_silence )
         {
            if ( temp_end - cur_start >= min_speech ) seg_combine( s, &buffered, cur_start, temp_end );
            cur_start = 0;
            temp_end = 0;
            triggered = 0;
         }
      }
   }
   /* vadc.c:1005-1027 */
   if ( triggered )
   {
      int audio_len = ( g - 1 ) * input_count;
      if ( audio_len - cur_start * input_count > min_speech * input_count )
         seg_combine( s, &buffered, cur_start, audio_len / input_count );
   }
   if ( buffered.valid ) seg_emit( s, buffered.start, buffered.end );
}

size_t so_segments_text( const float *prob, long long nchunks, const so_segment_params *p, char *text, size_t cap )
{
   seg_sink s;
   memset( &s, 0, sizeof( s ) );
   s.text = text;
   s.cap = cap;
   if ( cap ) text[0] = 0;
   seg_run( prob, nchunks, p, &s );
   return s.len;
}

long long so_segments_chunks( const float *prob, long long nchunks, const so_segment_params *p,
                              int *pairs, long long max_pairs )
{
   seg_sink s;
   memset( &s, 0, sizeof( s ) );
   s.pairs = pairs;
   s.max_pairs = max_pairs;
   seg_run( prob, nchunks, p, &s );
   return s.npairs;
}
