// Prints every enumerator / layout fact of cudpp.h that a caller compiled against the reference header
// bakes in.  Built twice: against /root/reference/cudpp-inpar/include (tools/make_abi_golden.sh ->
// tests/golden/cudpp_abi.txt) and against include/ (tests/test_abi.py); the outputs must be equal.
#include <stdio.h>
#include <stddef.h>
#include "cudpp.h"
#define P(x) printf("    \"%s\": %ld,\n", #x, (long)(x))
int main(void){
P(CUDPP_SUCCESS);P(CUDPP_ERROR_INVALID_HANDLE);P(CUDPP_ERROR_ILLEGAL_CONFIGURATION);P(CUDPP_ERROR_INVALID_PLAN);P(CUDPP_ERROR_INSUFFICIENT_RESOURCES);P(CUDPP_ERROR_UNKNOWN);
P(CUDPP_OPTION_FORWARD);P(CUDPP_OPTION_BACKWARD);P(CUDPP_OPTION_EXCLUSIVE);P(CUDPP_OPTION_INCLUSIVE);P(CUDPP_OPTION_CTA_LOCAL);P(CUDPP_OPTION_KEYS_ONLY);P(CUDPP_OPTION_KEY_VALUE_PAIRS);
P(CUDPP_CHAR);P(CUDPP_UCHAR);P(CUDPP_SHORT);P(CUDPP_USHORT);P(CUDPP_INT);P(CUDPP_UINT);P(CUDPP_FLOAT);P(CUDPP_DOUBLE);P(CUDPP_LONGLONG);P(CUDPP_ULONGLONG);P(CUDPP_DATATYPE_INVALID);
P(CUDPP_ADD);P(CUDPP_MULTIPLY);P(CUDPP_MIN);P(CUDPP_MAX);P(CUDPP_OPERATOR_INVALID);
P(CUDPP_SCAN);P(CUDPP_SEGMENTED_SCAN);P(CUDPP_COMPACT);P(CUDPP_REDUCE);P(CUDPP_SORT_RADIX);P(CUDPP_SPMVMULT);P(CUDPP_RAND_MD5);P(CUDPP_TRIDIAGONAL);P(CUDPP_COMPRESS);P(CUDPP_LISTRANK);P(CUDPP_BWT);P(CUDPP_MTF);P(CUDPP_SA);P(CUDPP_ALGORITHM_INVALID);
P(sizeof(CUDPPConfiguration));P(offsetof(CUDPPConfiguration,algorithm));P(offsetof(CUDPPConfiguration,op));P(offsetof(CUDPPConfiguration,datatype));P(offsetof(CUDPPConfiguration,options));P(sizeof(CUDPPHandle));P(CUDPP_INVALID_HANDLE);
return 0;}
