/* <slow5/slow5.h> for code written against slow5lib's low-level API (slow5lib/include/slow5/slow5.h): maps the slow5_* names
 * this library provides onto its s5b_* entry points, so that e.g. slow5lib/examples/adv/sequential_read_pthreads.c builds
 * unchanged with  -I include/compat -L slow5tools_b200 -lslow5b200 . */
#ifndef S5B_COMPAT_SLOW5_H
#define S5B_COMPAT_SLOW5_H
#ifndef S5B_SLOW5_COMPAT
#define S5B_SLOW5_COMPAT
#endif
#include "../../slow5b200_file.h"
/* error codes, slow5lib/include/slow5/slow5_error.h + slow5_defs.h:137-154 */
#define SLOW5_ERR_EOF      S5B_ERR_EOF
#define SLOW5_ERR_ARG      S5B_ERR_ARG
#define SLOW5_ERR_IO       S5B_ERR_IO
#define SLOW5_ERR_RECPARSE S5B_ERR_RECPARSE
#define SLOW5_ERR_MEM      S5B_ERR_MEM
#define SLOW5_ERR_PRESS    S5B_ERR_PRESS
#define SLOW5_ERR_NOIDX    S5B_ERR_NOIDX
#define SLOW5_ERR_NOTFOUND S5B_ERR_NOTFOUND
#define SLOW5_ERR_NOAUX    S5B_ERR_NOAUX
#define SLOW5_ERR_NOFLD    S5B_ERR_NOFLD
#define SLOW5_ERR_TYPE     S5B_ERR_TYPE
/* enum slow5_aux_type, slow5.h:106-133 (what slow5_get_aux_types hands out) */
enum slow5_aux_type {
    SLOW5_INT8_T = 0, SLOW5_INT16_T, SLOW5_INT32_T, SLOW5_INT64_T, SLOW5_UINT8_T, SLOW5_UINT16_T, SLOW5_UINT32_T, SLOW5_UINT64_T,
    SLOW5_FLOAT, SLOW5_DOUBLE, SLOW5_CHAR, SLOW5_ENUM,
    SLOW5_INT8_T_ARRAY, SLOW5_INT16_T_ARRAY, SLOW5_INT32_T_ARRAY, SLOW5_INT64_T_ARRAY, SLOW5_UINT8_T_ARRAY, SLOW5_UINT16_T_ARRAY,
    SLOW5_UINT32_T_ARRAY, SLOW5_UINT64_T_ARRAY, SLOW5_FLOAT_ARRAY, SLOW5_DOUBLE_ARRAY, SLOW5_STRING, SLOW5_ENUM_ARRAY
};
#define SLOW5_IS_PTR(type) ((type) >= SLOW5_INT8_T_ARRAY)
/* what a primitive auxiliary field holds when its value is missing (slow5.h:139-150) */
#include <math.h>
#define SLOW5_INT8_T_NULL   INT8_MAX
#define SLOW5_INT16_T_NULL  INT16_MAX
#define SLOW5_INT32_T_NULL  INT32_MAX
#define SLOW5_INT64_T_NULL  INT64_MAX
#define SLOW5_UINT8_T_NULL  UINT8_MAX
#define SLOW5_UINT16_T_NULL UINT16_MAX
#define SLOW5_UINT32_T_NULL UINT32_MAX
#define SLOW5_UINT64_T_NULL UINT64_MAX
#define SLOW5_FLOAT_NULL    nanf("")
#define SLOW5_DOUBLE_NULL   nan("")
#define SLOW5_CHAR_NULL     0
#define SLOW5_ENUM_NULL     SLOW5_UINT8_T_NULL
/* enum slow5_press_method, slow5_press.h:61-67 */
#define SLOW5_COMPRESS_NONE   S5B_COMPRESS_NONE
#define SLOW5_COMPRESS_ZLIB   S5B_COMPRESS_ZLIB
#define SLOW5_COMPRESS_SVB_ZD S5B_COMPRESS_SVB_ZD
#define SLOW5_COMPRESS_ZSTD   S5B_COMPRESS_ZSTD
#define SLOW5_COMPRESS_EX_ZD  S5B_COMPRESS_EX_ZD
#endif
