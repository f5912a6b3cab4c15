/* vadc_b200/csrc/testtensor.h -- reader for the reference's .testtensor container (host C).
 *
 * Format (tensor.h:201-253, writer utils.py:7-53), little endian:
 *   int32 version (=1), int32 count
 *   count x { int32 name_len; char name[name_len] }
 *   count x { int32 ndim; int32 dims[ndim]; int32 size; int32 nbytes; float32 data[size] }
 * Binding is positional (tensor.h:114-191); names are informational.
 */
#ifndef VB_TESTTENSOR_H
#define VB_TESTTENSOR_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vb_tensor
{
   int ndim;
   int dims[8];
   int size;
   float *data; /* owned by the file object, 16-byte aligned */
   char name[96];
} vb_tensor;

typedef struct vb_tensor_file
{
   int count;
   vb_tensor *tensors;
   float *storage;
} vb_tensor_file;

/* returns 0 on success; on failure `err` (if given) receives a short reason */
int vb_testtensor_parse( const void *bytes, size_t nbytes, vb_tensor_file *out, char *err, size_t errcap );
void vb_testtensor_free( vb_tensor_file *f );

/* checks the 99-tensor Silero v3.1 16 kHz layout (silero.h:31-33, tensor.h:154-191) and shapes */
int vb_silero_v31_check( const vb_tensor_file *f, char *err, size_t errcap );

#ifdef __cplusplus
}
#endif
#endif
