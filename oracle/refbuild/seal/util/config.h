// Hand-written build configuration for compiling the vendored SEAL-4.1-bs sources directly with
// g++ (no cmake).  This file stands in for the header the reference's build system would
// generate from native/src/seal/util/config.h.in; the switches mirror the fork's CMake defaults
// (thirdparty/SEAL-4.1-bs/CMakeLists.txt:61-370) with every network-fetched dependency OFF
// (MSGSL, ZLIB, ZSTD, HEXL).  It is test infrastructure for oracle/_ref only.
#pragma once

#define SEAL_VERSION "4.1.2"
#define SEAL_VERSION_MAJOR 4
#define SEAL_VERSION_MINOR 1
#define SEAL_VERSION_PATCH 2

// C++17 features (SEAL_USE_CXX17 default ON)
#define SEAL_USE_STD_BYTE
#define SEAL_USE_ALIGNED_ALLOC
#define SEAL_USE_SHARED_MUTEX
#define SEAL_USE_IF_CONSTEXPR
#define SEAL_USE_MAYBE_UNUSED
#define SEAL_USE_NODISCARD
#define SEAL_USE_STD_FOR_EACH_N

// Security: the fork ships SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT OFF (CMakeLists.txt:246-248)
#define SEAL_DEFAULT_PRNG Blake2xb

// Intrinsics
#define SEAL_USE_INTRIN
#define SEAL_USE___BUILTIN_CLZLL
#define SEAL_USE___INT128
#define SEAL_USE__ADDCARRY_U64
#define SEAL_USE__SUBBORROW_U64

// Zero memory functions
#define SEAL_USE_EXPLICIT_BZERO
