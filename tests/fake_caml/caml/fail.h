#ifndef FAKE_CAML_FAIL_H
#define FAKE_CAML_FAIL_H
void caml_failwith(const char *) __attribute__((noreturn));
void caml_raise_out_of_memory(void) __attribute__((noreturn));
#endif
