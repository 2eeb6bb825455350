#ifndef AWS_COMPRESSION_EXPORTS_H
#define AWS_COMPRESSION_EXPORTS_H
/*
 * Symbol visibility for libaws-c-compression (B200 build).
 * Same macro name and switches as the reference's include/aws/compression/exports.h:7-25 so
 * downstream build flags (AWS_COMPRESSION_USE_IMPORT_EXPORT / AWS_COMPRESSION_EXPORTS) keep working.
 */
#if defined(_WIN32) || defined(AWS_CRT_USE_WINDOWS_DLL_SEMANTICS)
#    if defined(AWS_COMPRESSION_USE_IMPORT_EXPORT) && defined(AWS_COMPRESSION_EXPORTS)
#        define AWS_COMPRESSION_API __declspec(dllexport)
#    elif defined(AWS_COMPRESSION_USE_IMPORT_EXPORT)
#        define AWS_COMPRESSION_API __declspec(dllimport)
#    else
#        define AWS_COMPRESSION_API
#    endif
#elif defined(AWS_COMPRESSION_USE_IMPORT_EXPORT) && defined(AWS_COMPRESSION_EXPORTS)
#    define AWS_COMPRESSION_API __attribute__((visibility("default")))
#else
#    define AWS_COMPRESSION_API
#endif

#endif /* AWS_COMPRESSION_EXPORTS_H */
