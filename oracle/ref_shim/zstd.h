/* Declaration shim for the 9 libzstd (v1.5.5) entry points the reference calls
 * (reference: stenos/internal/zstd_wrapper.h:55-88, stenos/internal/stenos.cpp:696-732).
 * The image has libzstd.so.1.5.5 at run time but no development headers.
 * TEST INFRASTRUCTURE ONLY -- used to build the untouched reference into oracle/_ref/. */
#ifndef STENOS_B200_ZSTD_SHIM_H
#define STENOS_B200_ZSTD_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_CDict_s ZSTD_CDict;
ZSTD_CCtx* ZSTD_createCCtx(void);
size_t ZSTD_freeCCtx(ZSTD_CCtx* cctx);
size_t ZSTD_compressCCtx(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize, int compressionLevel);
size_t ZSTD_compress_usingCDict(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize, const ZSTD_CDict* cdict);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
unsigned ZSTD_isError(size_t code);
int ZSTD_maxCLevel(void);
#ifdef __cplusplus
}
#endif
#endif
