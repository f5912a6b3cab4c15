// H2D copy rates from pinned memory: 1-D vs the strided 2-D pattern of run_streams (4096 rows, pitch 125*3072 B)
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
int main()
{
   const size_t S = 4096, rowbytes = 125 * 3072, total = S * rowbytes;
   char *h, *d;
   cudaHostAlloc( &h, total, cudaHostAllocDefault );
   cudaMalloc( &d, total );
   cudaStream_t st;
   cudaStreamCreate( &st );
   for ( int rep = 0; rep < 3; ++rep )
   {
      auto t0 = std::chrono::steady_clock::now();
      cudaMemcpyAsync( d, h, total, cudaMemcpyHostToDevice, st );
      cudaStreamSynchronize( st );
      double s = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
      printf( "1-D %.1f GB/s\n", total / s / 1e9 );
   }
   const int widths[] = { 5, 10, 20, 25, 125 };
   for ( int wi = 0; wi < 5; ++wi )
   {
      const size_t w = widths[wi] * 3072;
      const int nwin = 125 / widths[wi];
      for ( int rep = 0; rep < 2; ++rep )
      {
         auto t0 = std::chrono::steady_clock::now();
         for ( int k = 0; k < nwin; ++k ) cudaMemcpy2DAsync( d + k * w * S, w, h + k * w, rowbytes, w, S, cudaMemcpyHostToDevice, st );
         cudaStreamSynchronize( st );
         double s = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
         printf( "2-D rows of %d chunks (%zu B) x %d windows: %.1f GB/s\n", widths[wi], w, nwin, nwin * w * S / s / 1e9 );
      }
   }
   return 0;
}
