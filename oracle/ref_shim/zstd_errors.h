/* Declaration shim, see zstd.h in this directory. Values from zstd v1.5.5's public ABI. */
#ifndef STENOS_B200_ZSTD_ERRORS_SHIM_H
#define STENOS_B200_ZSTD_ERRORS_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { ZSTD_error_no_error = 0, ZSTD_error_dstSize_tooSmall = 70 } ZSTD_ErrorCode;
ZSTD_ErrorCode ZSTD_getErrorCode(size_t functionResult);
#ifdef __cplusplus
}
#endif
#endif
