// tests/hostcheck/faithful_host_check.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles vadc_b200/csrc/faithful_kernel.cuh for the host (g++ -O2 -ffp-contract=off, the flags the oracle is pinned
// with) and runs the very code of the CUDA path serially: one "thread" (tid 0 of 1), barriers as no-ops. What the kernels
// compute on the device differs from this only in the per-operation intrinsics (__fmul_rn etc., IEEE like the host's) and
// in expf, which libm_exact.cuh restates. tests/test_faithful_host.py compares the results with the oracle bit for bit.
#include "../../vadc_b200/csrc/faithful_kernel.cuh"

#include <stdlib.h>

extern "C" void faithful_host_encoder( const float *const *tensors /*[99]*/, const float *magnitude /*[B][129][25]*/, int batch, float *a4 /*[B][7][64]*/ )
{
   fq::Weights W;
   for ( int i = 0; i < 99; ++i ) W.t[i] = tensors[i];
   float *tt[fq::N_TRANSPOSED];
   for ( int k = 0; k < fq::N_TRANSPOSED; ++k )
   {
      // the transposed copies create_impl (engine.cu) uploads next to the originals
      int idx, n_out, n_in;
      fq::transposed_slot( k, &idx, &n_out, &n_in );
      tt[k] = (float *)malloc( sizeof( float ) * (size_t)n_out * n_in );
      for ( int r = 0; r < n_out; ++r )
         for ( int c = 0; c < n_in; ++c ) tt[k][(size_t)c * n_out + r] = tensors[idx][(size_t)r * n_in + c];
      W.tt[k] = tt[k];
   }
   float *sm = (float *)malloc( sizeof( float ) * fq::SM_FLOATS );
   float *logspec = (float *)malloc( sizeof( float ) * 129 * 25 );
   for ( int b = 0; b < batch; ++b )
   {
      // misc.c:40-46, the part of the normalization the STFT kernel does on the device
      for ( int i = 0; i < 129 * 25; ++i ) logspec[i] = log1pf( magnitude[(size_t)b * 3225 + i] * 1048576.0f );
      fq::encoder_chunk( W, logspec, a4 + (size_t)b * 448, sm, 0, 1 );
   }
   free( sm );
   free( logspec );
   for ( int k = 0; k < fq::N_TRANSPOSED; ++k ) free( tt[k] );
}

extern "C" void faithful_host_decoder( const float *hs /*[B][7][64]*/, int batch, const float *w /*[2][64]*/, const float *b /*[2]*/, float *out /*[B][2]*/ )
{
   for ( int n = 0; n < batch; ++n )
      for ( int head = 0; head < 2; ++head ) out[n * 2 + head] = fq::decoder_head( hs + (size_t)n * 448, w + head * 64, b[head] );
}

// The decoder LSTM as faithful_lstm_kernel<LAYER> walks it: one layer over all steps, then the next; W packed as
// [layer][k/4][row][4] (pack_lstm, engine.cu), gate pre-activations through fq::gate_dot. The cell update below restates the
// kernel's (lstm.c:64-88) with the host's libm. x: [steps][64]; state: h[2][64], c[2][64] (updated); out: [steps][64].
extern "C" void faithful_host_lstm( const float *x, int steps, float *h, float *c, const float *wpack, const float *bias, float *out )
{
   float *seq = (float *)malloc( sizeof( float ) * (size_t)steps * 64 );
   const float *in = x;
   for ( int layer = 0; layer < 2; ++layer )
   {
      const float *Ws = wpack + (size_t)layer * 32 * 256 * 4;
      alignas( 16 ) float xh[128];
      float *hl = h + layer * 64, *cl = c + layer * 64;
      float *dst = layer == 0 ? seq : out;
      for ( int s = 0; s < steps; ++s )
      {
         for ( int j = 0; j < 64; ++j )
         {
            xh[j] = in[(size_t)s * 64 + j];
            xh[64 + j] = hl[j];
         }
         for ( int j = 0; j < 64; ++j )
         {
            float z[4];
            for ( int g = 0; g < 4; ++g ) z[g] = fq::gate_dot( xh, Ws + ( g * 64 + j ) * 4, 1024 ) + bias[layer * 256 + g * 64 + j];
            const float ig = 1.0f / ( 1.0f + expf( -z[0] ) ), fg = 1.0f / ( 1.0f + expf( -z[1] ) ), gg = tanhf( z[2] ), og = 1.0f / ( 1.0f + expf( -z[3] ) );
            const float cn = fg * cl[j] + ig * gg;
            cl[j] = cn;
            dst[(size_t)s * 64 + j] = tanhf( cn ) * og;
         }
         for ( int j = 0; j < 64; ++j ) hl[j] = dst[(size_t)s * 64 + j];
      }
      in = seq;
   }
   free( seq );
}
