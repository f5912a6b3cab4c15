// scripts/microbench/libm_latency.cu -- dependent-chain latency (cycles per call, one warp alone on an SM) of the bit-exact libm restatements
// the LSTM's serial step is made of. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vadc_b200/csrc -o libm_latency libm_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "libm_exact.cuh"

template <int WHAT>
__global__ void k( float *out, float x0, long long *cycles, int n )
{
   float x = x0 + threadIdx.x * 1e-3f;
   long long t0 = clock64();
   for ( int i = 0; i < n; ++i )
   {
      float y;
      if ( WHAT == 0 ) y = lme::tanhf_ref( x );
      if ( WHAT == 1 ) y = lme::sigmoid_ref( x );
      if ( WHAT == 2 ) y = lme::expf_ref( -x );
      if ( WHAT == 3 ) y = __fdiv_rn( 1.0f, __fadd_rn( x, 2.0f ) );
      if ( WHAT == 4 ) y = lme::expm1f_ref( x );
      if ( WHAT == 5 ) y = tanhf( x );
      x = __fadd_rn( __fmul_rn( y, 0.37f ), x0 ); // next argument depends on this result, stays in the typical gate range
   }
   long long t1 = clock64();
   out[threadIdx.x] = x;
   if ( threadIdx.x == 0 ) *cycles = t1 - t0;
}

template <int WHAT>
void run( const char *name, float x0 )
{
   float *out;
   long long *cyc, h;
   cudaMalloc( &out, 128 );
   cudaMalloc( &cyc, 8 );
   const int n = 4096;
   k<WHAT><<<1, 32>>>( out, x0, cyc, 64 );
   k<WHAT><<<1, 32>>>( out, x0, cyc, n );
   cudaMemcpy( &h, cyc, 8, cudaMemcpyDeviceToHost );
   printf( "%-28s x0 = %6.2f: %7.1f cycles per call (incl. 2 dependent flops)\n", name, x0, (double)h / n );
   cudaFree( out ); cudaFree( cyc );
}

int main()
{
   for ( float x0 : { 0.3f, 1.7f, 6.0f, -2.5f } )
   {
      run<0>( "tanhf_ref", x0 );
      run<1>( "sigmoid_ref", x0 );
      run<2>( "expf_ref", x0 );
      run<3>( "__fdiv_rn", x0 );
      run<4>( "expm1f_ref", x0 );
      run<5>( "CUDA tanhf (for scale)", x0 );
   }
   return 0;
}
