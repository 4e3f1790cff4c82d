/* Hand-written stand-in for the autoconf-generated config.h of Jellyfish 2.2.5
 * (reference: src/modifiedJellyfish/config.h.in). Test infrastructure only:
 * lets oracle/Makefile compile the reference's own sources in place with g++,
 * without running the reference's configure/cmake. Values = what configure
 * detects on x86-64 Linux with g++ 13. */
#ifndef RG_ORACLE_JF_CONFIG_H
#define RG_ORACLE_JF_CONFIG_H
#define HAVE_DLFCN_H 1
#define HAVE_EXECINFO_H 1
#define HAVE_EXT_STDIO_FILEBUF_H 1
#define HAVE_INT128 1
#define HAVE_INTTYPES_H 1
#define HAVE_MEMORY_H 1
#define HAVE_NUMERIC_LIMITS128 1
#define HAVE_SI_INT 1
#define HAVE_SSE 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRINGS_H 1
#define HAVE_STRING_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define STDC_HEADERS 1
#define PACKAGE "jellyfish"
#define PACKAGE_BUGREPORT "gmarcais@umd.edu"
#define PACKAGE_NAME "jellyfish"
#define PACKAGE_STRING "jellyfish 2.2.5"
#define PACKAGE_TARNAME "jellyfish"
#define PACKAGE_URL ""
#define PACKAGE_VERSION "2.2.5"
#define VERSION "2.2.5"
#endif
