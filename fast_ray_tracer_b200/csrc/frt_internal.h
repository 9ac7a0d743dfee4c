/* frt_internal.h -- shared between the C and CUDA translation units of libfrt_b200.so (not installed). */
#ifndef FRT_INTERNAL_H
#define FRT_INTERNAL_H

#ifdef __cplusplus
extern "C" {
#endif

/* records the message for frt_last_error() (thread-local) and returns `code` */
int frt_set_error(int code, const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
