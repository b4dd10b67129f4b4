/* Minimal stand-ins for the OCaml runtime headers: ONLY so that gcc can parse and type-check mmo_b200/ocaml/gpu_stubs.c
 * against include/mmo_b200.h in an image without an OCaml toolchain (tests/test_abi.py).  Nothing here is linked or run. */
#ifndef FAKE_CAML_MLVALUES_H
#define FAKE_CAML_MLVALUES_H
#include <stddef.h>
#include <stdint.h>
typedef intptr_t value;
typedef size_t mlsize_t;
#define Val_unit ((value)1)
#define Val_long(x) ((value)(((intptr_t)(x) << 1) + 1))
#define Long_val(v) ((intptr_t)(v) >> 1)
#define Int_val(v) ((int)Long_val(v))
#define Bool_val(v) Int_val(v)
#define Is_block(v) (((v) & 1) == 0)
#define Field(v, i) (((value *)(v))[i])
#define Wosize_val(v) ((mlsize_t)(((value *)(v))[-1] >> 10))
#define Double_wosize 1
#define Double_val(v) (*(double *)(v))
#define Double_flat_field(v, i) (((double *)(v))[i])
#define Store_double_flat_field(v, i, d) (((double *)(v))[i] = (d))
#define CAMLprim
#endif
