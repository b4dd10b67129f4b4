#ifndef FAKE_CAML_CUSTOM_H
#define FAKE_CAML_CUSTOM_H
#include "mlvalues.h"
struct custom_operations {
  const char *identifier; void (*finalize)(value); int (*compare)(value, value); intptr_t (*hash)(value);
  void (*serialize)(value, uintptr_t *, uintptr_t *); uintptr_t (*deserialize)(void *); int (*compare_ext)(value, value);
  const void *fixed_length;
};
#define custom_compare_default NULL
#define custom_hash_default NULL
#define custom_serialize_default NULL
#define custom_deserialize_default NULL
#define custom_compare_ext_default NULL
#define custom_fixed_length_default NULL
value caml_alloc_custom(struct custom_operations *, uintptr_t, mlsize_t, mlsize_t);
#define Data_custom_val(v) ((void *)&Field(v, 1))
#endif
