/* vadc_b200/csrc/testtensor.c -- see testtensor.h */
#include "testtensor.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct cursor
{
   const unsigned char *p;
   size_t n, off;
} cursor;

static int take_i32( cursor *c, int *v )
{
   if ( c->off + 4 > c->n ) return -1;
   int32_t x;
   memcpy( &x, c->p + c->off, 4 );
   c->off += 4;
   *v = (int)x;
   return 0;
}

static int fail( char *err, size_t cap, const char *msg )
{
   if ( err && cap ) snprintf( err, cap, "%s", msg );
   return -1;
}

int vb_testtensor_parse( const void *bytes, size_t nbytes, vb_tensor_file *out, char *err, size_t errcap )
{
   memset( out, 0, sizeof( *out ) );
   if ( !bytes ) return fail( err, errcap, "null blob" );
   cursor c = { (const unsigned char *)bytes, nbytes, 0 };
   int version = 0, count = 0;
   if ( take_i32( &c, &version ) || take_i32( &c, &count ) ) return fail( err, errcap, "truncated header" );
   if ( version != 1 ) return fail( err, errcap, "unsupported .testtensor version" );
   if ( count <= 0 || count > 65536 ) return fail( err, errcap, "bad tensor count" );

   vb_tensor *t = (vb_tensor *)calloc( (size_t)count, sizeof( vb_tensor ) );
   if ( !t ) return fail( err, errcap, "out of memory" );
   for ( int i = 0; i < count; ++i )
   {
      int len = 0;
      if ( take_i32( &c, &len ) || len < 0 || c.off + (size_t)len > c.n )
      {
         free( t );
         return fail( err, errcap, "truncated name table" );
      }
      size_t keep = (size_t)len < sizeof( t[i].name ) - 1 ? (size_t)len : sizeof( t[i].name ) - 1;
      memcpy( t[i].name, c.p + c.off, keep );
      c.off += (size_t)len;
   }
   /* first pass over the bodies: shapes and total payload */
   size_t body = c.off;
   size_t total = 0;
   for ( int i = 0; i < count; ++i )
   {
      int nb = 0;
      if ( take_i32( &c, &t[i].ndim ) || t[i].ndim < 0 || t[i].ndim > 8 ) goto bad;
      for ( int d = 0; d < t[i].ndim; ++d )
         if ( take_i32( &c, &t[i].dims[d] ) || t[i].dims[d] < 0 ) goto bad;
      if ( take_i32( &c, &t[i].size ) || take_i32( &c, &nb ) || t[i].size < 0 ) goto bad;
      /* all in 64 bits with a cap per step: a crafted header (dims 32768 x 32768, nbytes 0) must not wrap `size * 4` or the product */
      if ( t[i].size > 0x7fffffff / 4 || nb < 0 ) goto bad;
      long long prod = 1;
      for ( int d = 0; d < t[i].ndim; ++d )
      {
         prod *= t[i].dims[d];
         if ( prod > 0x7fffffffll ) goto bad;
      }
      if ( prod != (long long)t[i].size || (long long)nb != (long long)t[i].size * 4 || c.off + (size_t)nb > c.n ) goto bad;
      c.off += (size_t)nb;
      total += ( (size_t)t[i].size + 3 ) & ~(size_t)3;
   }
   if ( c.off != c.n ) goto bad;

   float *storage = 0;
   if ( posix_memalign( (void **)&storage, 64, ( total + 4 ) * sizeof( float ) ) )
   {
      free( t );
      return fail( err, errcap, "out of memory" );
   }
   memset( storage, 0, ( total + 4 ) * sizeof( float ) );
   /* second pass: copy payloads into aligned storage */
   c.off = body;
   size_t at = 0;
   for ( int i = 0; i < count; ++i )
   {
      c.off += 4 + 4 * (size_t)t[i].ndim + 8;
      t[i].data = storage + at;
      memcpy( t[i].data, c.p + c.off, (size_t)t[i].size * 4 );
      c.off += (size_t)t[i].size * 4;
      at += ( (size_t)t[i].size + 3 ) & ~(size_t)3;
   }
   out->count = count;
   out->tensors = t;
   out->storage = storage;
   return 0;
bad:
   free( t );
   return fail( err, errcap, "malformed tensor body" );
}

void vb_testtensor_free( vb_tensor_file *f )
{
   if ( !f ) return;
   free( f->tensors );
   free( f->storage );
   memset( f, 0, sizeof( *f ) );
}

static int shape_is( const vb_tensor *t, int ndim, int d0, int d1, int d2 )
{
   if ( t->ndim != ndim ) return 0;
   if ( ndim > 0 && t->dims[0] != d0 ) return 0;
   if ( ndim > 1 && t->dims[1] != d1 ) return 0;
   if ( ndim > 2 && t->dims[2] != d2 ) return 0;
   return 1;
}

int vb_silero_v31_check( const vb_tensor_file *f, char *err, size_t errcap )
{
   if ( f->count != 99 ) return fail( err, errcap, "expected 99 tensors (1 basis + 94 encoder + 2 lstm + 2 decoder)" );
   const vb_tensor *t = f->tensors;
   if ( !shape_is( t + 0, 3, 258, 1, 256 ) ) return fail( err, errcap, "tensor 0 must be the [258,1,256] STFT basis" );
   static const int first[4] = { 1, 25, 49, 71 };
   static const int cin[4] = { 129, 16, 32, 32 }, cc[4] = { 16, 32, 32, 64 }, proj[4] = { 1, 1, 0, 1 };
   for ( int l = 0; l < 4; ++l )
   {
      const vb_tensor *w = t + first[l];
      int C = cc[l], I = cin[l], i = 0, ok = 1;
      ok &= shape_is( w + i++, 3, I, 1, 5 );
      ok &= shape_is( w + i++, 1, I, 0, 0 );
      ok &= shape_is( w + i++, 3, C, I, 1 );
      ok &= shape_is( w + i++, 1, C, 0, 0 );
      if ( proj[l] )
      {
         ok &= shape_is( w + i++, 3, C, I, 1 );
         ok &= shape_is( w + i++, 1, C, 0, 0 );
      }
      ok &= shape_is( w + i++, 2, 3 * C, C, 0 );
      ok &= shape_is( w + i++, 1, 3 * C, 0, 0 );
      ok &= shape_is( w + i++, 2, C, C, 0 );
      ok &= shape_is( w + i++, 1, C, 0, 0 );
      ok &= shape_is( w + i++, 1, C, 0, 0 ); /* norm1 w,b */
      ok &= shape_is( w + i++, 1, C, 0, 0 );
      ok &= shape_is( w + i++, 2, C, C, 0 ); /* linear1 */
      ok &= shape_is( w + i++, 1, C, 0, 0 );
      ok &= shape_is( w + i++, 2, C, C, 0 ); /* linear2 */
      ok &= shape_is( w + i++, 1, C, 0, 0 );
      ok &= shape_is( w + i++, 1, C, 0, 0 ); /* norm2 w,b */
      ok &= shape_is( w + i++, 1, C, 0, 0 );
      ok &= shape_is( w + i++, 3, C, C, 1 ); /* conv */
      ok &= shape_is( w + i++, 1, C, 0, 0 );
      for ( int k = 0; k < 4; ++k ) ok &= shape_is( w + i++, 1, C, 0, 0 ); /* bn w,b,mean,var */
      if ( !ok ) return fail( err, errcap, "encoder layer tensor has an unexpected shape" );
   }
   if ( !shape_is( t + 95, 3, 2, 256, 128 ) ) return fail( err, errcap, "tensor 95 must be LSTM weights [2,256,128]" );
   if ( !shape_is( t + 96, 2, 2, 256, 0 ) ) return fail( err, errcap, "tensor 96 must be LSTM biases [2,256]" );
   if ( !shape_is( t + 97, 3, 2, 64, 1 ) ) return fail( err, errcap, "tensor 97 must be decoder weights [2,64,1]" );
   if ( !shape_is( t + 98, 1, 2, 0, 0 ) ) return fail( err, errcap, "tensor 98 must be decoder biases [2]" );
   return 0;
}
