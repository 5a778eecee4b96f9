/*
 * cudpp.h -- the CUDPP entry points of libb200lc.so for the lossless-compression path.
 *
 * Source-compatible subset of cudpp-inpar/include/cudpp.h (reference lines in brackets): the
 * result codes [32-49], options [54-77], datatypes [84-97], operators [104-111], algorithm ids
 * [118-137] and bucket mappers keep the reference's enumerator ORDER (= values), and
 * CUDPPConfiguration [171-178] keeps its layout and is passed by value, so a caller compiled
 * against the reference header links against libb200lc.so unchanged.
 *
 * Implemented algorithms: CUDPP_COMPRESS, CUDPP_BWT, CUDPP_MTF, CUDPP_SA with datatype
 * CUDPP_UCHAR, and CUDPP_SORT_RADIX (unsigned char / unsigned int keys) for the reference test's
 * own decoder.  cudppPlan with any other algorithm returns CUDPP_ERROR_ILLEGAL_CONFIGURATION
 * (those primitives are outside the hot path).
 *
 * All data pointers are DEVICE pointers owned by the caller; plans own their scratch
 * (reference: cudpp_plan.cpp:712-762).  Work is issued on the default stream.  The BWT stage
 * synchronises internally (prefix doubling reads one counter per round), like the reference's
 * cudppCompress, which blocks on a D2H copy (compress_app.cu:106).
 */
#ifndef B200LC_CUDPP_H_
#define B200LC_CUDPP_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum CUDPPResult
{
    CUDPP_SUCCESS = 0,
    CUDPP_ERROR_INVALID_HANDLE,
    CUDPP_ERROR_ILLEGAL_CONFIGURATION,
    CUDPP_ERROR_INVALID_PLAN,
    CUDPP_ERROR_INSUFFICIENT_RESOURCES,
    CUDPP_ERROR_UNKNOWN = 9999
};

enum CUDPPOption
{
    CUDPP_OPTION_FORWARD = 0x1,
    CUDPP_OPTION_BACKWARD = 0x2,
    CUDPP_OPTION_EXCLUSIVE = 0x4,
    CUDPP_OPTION_INCLUSIVE = 0x8,
    CUDPP_OPTION_CTA_LOCAL = 0x10,
    CUDPP_OPTION_KEYS_ONLY = 0x20,
    CUDPP_OPTION_KEY_VALUE_PAIRS = 0x40
};

enum CUDPPDatatype
{
    CUDPP_CHAR, CUDPP_UCHAR, CUDPP_SHORT, CUDPP_USHORT, CUDPP_INT, CUDPP_UINT, CUDPP_FLOAT,
    CUDPP_DOUBLE, CUDPP_LONGLONG, CUDPP_ULONGLONG, CUDPP_DATATYPE_INVALID
};

enum CUDPPOperator { CUDPP_ADD, CUDPP_MULTIPLY, CUDPP_MIN, CUDPP_MAX, CUDPP_OPERATOR_INVALID };

enum CUDPPAlgorithm
{
    CUDPP_SCAN, CUDPP_SEGMENTED_SCAN, CUDPP_COMPACT, CUDPP_REDUCE, CUDPP_SORT_RADIX,
    CUDPP_SORT_MERGE, CUDPP_SORT_STRING, CUDPP_SPMVMULT, CUDPP_RAND_MD5, CUDPP_TRIDIAGONAL,
    CUDPP_COMPRESS, CUDPP_LISTRANK, CUDPP_BWT, CUDPP_MTF, CUDPP_SA, CUDPP_MULTISPLIT,
    CUDPP_ALGORITHM_INVALID
};

enum CUDPPBucketMapper
{
    CUDPP_LSB_BUCKET_MAPPER, CUDPP_MSB_BUCKET_MAPPER, CUDPP_DEFAULT_BUCKET_MAPPER,
    CUDPP_CUSTOM_BUCKET_MAPPER
};

struct CUDPPConfiguration
{
    enum CUDPPAlgorithm algorithm;
    enum CUDPPOperator op;
    enum CUDPPDatatype datatype;
    unsigned int options;
    enum CUDPPBucketMapper bucket_mapper;
};

#define CUDPP_INVALID_HANDLE 0xC0DABAD1
typedef size_t CUDPPHandle;

#ifdef __cplusplus
typedef CUDPPResult CUDPPResult_t;
typedef CUDPPConfiguration CUDPPConfiguration_t;
#else
typedef enum CUDPPResult CUDPPResult_t;
typedef struct CUDPPConfiguration CUDPPConfiguration_t;
#endif

/* cudpp.h:200-217 / cudpp_manager.cpp:39-45 / cudpp_plan.cpp:81-199 */
CUDPPResult_t cudppCreate(CUDPPHandle *theCudpp);
CUDPPResult_t cudppDestroy(CUDPPHandle theCudpp);
CUDPPResult_t cudppPlan(const CUDPPHandle cudppHandle, CUDPPHandle *planHandle,
                        CUDPPConfiguration_t config, size_t n, size_t rows, size_t rowPitch);
CUDPPResult_t cudppDestroyPlan(CUDPPHandle plan);

/* cudpp.cpp:764-806.  numElements bytes at d_uncompressed (the reference documents exactly
 * 1,048,576; any multiple of 4096 up to the plan's n works here).  Outputs: *d_bwtIndex,
 * d_hist[256] (histogram of the MTF output), d_encodeOffset[numElements/4096] (word offset of
 * every block's [nWords] cell), *d_compressedSize (words), d_compressed (capacity
 * (1536 + 1) words per 4096 input bytes, cudpp_globals.h:65-66,81-84).  d_histSize is unused,
 * as in the reference. */
CUDPPResult_t cudppCompress(CUDPPHandle planHandle, unsigned char *d_uncompressed, int *d_bwtIndex,
                            unsigned int *d_histSize, unsigned int *d_hist,
                            unsigned int *d_encodeOffset, unsigned int *d_compressedSize,
                            unsigned int *d_compressed, size_t numElements);
/* cudpp.cpp:826-862 */
CUDPPResult_t cudppBurrowsWheelerTransform(CUDPPHandle planHandle, unsigned char *d_in,
                                           unsigned char *d_out, int *d_index, size_t numElements);
/* cudpp.cpp:881-915 */
CUDPPResult_t cudppMoveToFrontTransform(CUDPPHandle planHandle, unsigned char *d_in,
                                        unsigned char *d_out, size_t numElements);
/* cudpp.cpp (cudppSuffixArray) / sa_app.cu:365-391: d_keys_sa[0..numElements) = suffix start
 * positions in sorted order. */
CUDPPResult_t cudppSuffixArray(CUDPPHandle planHandle, unsigned char *d_str,
                               unsigned int *d_keys_sa, size_t numElements);
/* cudpp.h:256-259: stable ascending radix sort in place, plan algorithm CUDPP_SORT_RADIX with
 * datatype CUDPP_UCHAR or CUDPP_UINT and CUDPP_OPTION_KEYS_ONLY or CUDPP_OPTION_KEY_VALUE_PAIRS
 * (values: unsigned int).  This is the sort the reference's own compress test decodes its BWT
 * with (apps/cudpp_testrig/test_compress.cpp:318-344), so the unmodified testrig links against
 * libb200lc.so alone. */
CUDPPResult_t cudppRadixSort(const CUDPPHandle planHandle, void *d_keys, void *d_values, size_t numElements);

#ifdef __cplusplus
}
#endif
#endif /* B200LC_CUDPP_H_ */
