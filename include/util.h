/* util.h -- the diagnostic convention of libhorizonator's host code.
 *
 * The reference reports every failure as one line on stderr, "<file>(<line>) at <function>(): <text>",
 * before returning false (its util.h:4); callers and log scrapers may rely on that shape, so the
 * B200-native library keeps it.  Nothing here aborts: the reference's assert(0) on GL errors has no
 * counterpart.
 */
#pragma once

#include <stdarg.h>
#include <stdio.h>

#if defined(__GNUC__)
__attribute__((format(printf, 4, 5), unused))
#endif
static void hz_msg_(const char* file, int line, const char* func, const char* fmt, ...)
{
    va_list ap;
    fprintf(stderr, "%s(%d) at %s(): ", file, line, func);
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}

#define MSG(...) hz_msg_(__FILE__, __LINE__, __func__, __VA_ARGS__)
