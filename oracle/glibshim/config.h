/* config.h -- hand-written stand-in for the reference's autoconf output.
 * TEST INFRASTRUCTURE ONLY (oracle build).  Mirrors what configure.ac would
 * define on this x86-64 host (configure.ac:404-480). */
#ifndef ORACLE_CONFIG_H
#define ORACLE_CONFIG_H
#define HAVE_MMX_INTRINSICS 1
#define HAVE_SSE41_INTRINSICS 1
#define HAVE_AVX2_INTRINSICS 1
#define HAVE_POPCNT64_INTRINSICS 1
#define HAVE_POPCNT32_INTRINSICS 1
#define HAVE_GCC_X86_FEATURE_BUILTINS 1
#define CHAFA_VERSION "1.19.0"
#define _CHAFA_EXTERN __attribute__((visibility("default"))) extern
#endif
