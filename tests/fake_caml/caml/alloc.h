#ifndef FAKE_CAML_ALLOC_H
#define FAKE_CAML_ALLOC_H
#include "mlvalues.h"
value caml_copy_double(double);
value caml_alloc_tuple(mlsize_t);
value caml_alloc_float_array(mlsize_t);
#endif
