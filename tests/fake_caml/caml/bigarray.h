#ifndef FAKE_CAML_BIGARRAY_H
#define FAKE_CAML_BIGARRAY_H
#include "mlvalues.h"
#define Caml_ba_data_val(v) ((void *)Field(v, 1))
#endif
