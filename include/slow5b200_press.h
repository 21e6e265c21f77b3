/* slow5b200_press.h -- the stateful half of the reference's codec interface (slow5lib/include/slow5/slow5_press.h:61-125),
 * provided by libslow5b200.so.  Same struct layouts (field order and sizes) and call shapes as the reference, names with
 * the s5b_ prefix; slow5b200_file.h's S5B_SLOW5_COMPAT maps the slow5_* spellings onto them and libslow5b200_compat.so
 * exports the slow5_* symbols themselves.
 *
 *   s5b_press_init / s5b_press_free            slow5_press_init / slow5_press_free       slow5_press.c:174-208
 *   __s5b_press_init / __s5b_press_free        __slow5_press_init / __slow5_press_free   slow5_press.c:218-328
 *   s5b_ptr_compress / s5b_ptr_depress         slow5_ptr_compress / slow5_ptr_depress    slow5_press.c:383-434, :499-571
 *   s5b_compress_footer_next                   slow5_compress_footer_next                slow5_press.c:784-811
 *
 * One behavioural difference, by design: the reference's zlib press object can continue ONE deflate stream over several
 * slow5_ptr_compress calls until slow5_compress_footer_next asks for Z_FINISH (used for nothing but whole records by
 * slow5lib itself: slow5_rec_to_mem calls footer_next before every record, slow5.c:4046-4050).  Here every call produces a
 * complete stream, with or without footer_next.  The codec work runs on the GPU through the calling thread's context
 * (s5b_ptr_compress_solo): correct for single buffers, meant for throughput only through the batch entry points.
 */
#ifndef SLOW5B200_PRESS_H
#define SLOW5B200_PRESS_H
#include <stddef.h>
#include "slow5b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* enum slow5_press_method values are S5B_COMPRESS_* (slow5b200.h) */
typedef struct {
    int record_method;
    int signal_method;
} s5b_press_method_t;                 /* slow5_press_method_t, slow5_press.h:68-71 */

struct __s5b_press {                  /* struct __slow5_press, slow5_press.h:86-89 */
    int method;
    void *stream;                     /* the reference keeps its z_streams here; here: per-object flags */
};
typedef struct s5b_press {            /* slow5_press_t, slow5_press.h:91-94 */
    struct __s5b_press *record_press;
    struct __s5b_press *signal_press;
} s5b_press_t;

s5b_press_t *s5b_press_init(s5b_press_method_t method);      /* NULL + s5b_last_error() on an unknown method */
struct __s5b_press *__s5b_press_init(int method);
void s5b_press_free(s5b_press_t *comp);
void __s5b_press_free(struct __s5b_press *comp);
/* malloc()'d result, *n its size; NULL and *n = 0 on failure (comp == NULL behaves like method NONE for compress and is an
 * argument error for depress, like the reference) */
void *s5b_ptr_compress(struct __s5b_press *comp, const void *ptr, size_t count, size_t *n);
void *s5b_ptr_depress(struct __s5b_press *comp, const void *ptr, size_t count, size_t *n);
void s5b_compress_footer_next(struct __s5b_press *comp);

#ifdef __cplusplus
}
#endif
#endif
