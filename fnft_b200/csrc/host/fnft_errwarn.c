/*
 * fnft_b200 host library -- error and warning channel.
 * Same observable behaviour as the reference: a per-thread printf-like pointer
 * (src/fnft_errwarn.c:42-60) through which messages of the form
 *   "FNFT Error: <msg>\n in <func>(<line>)-<version>\n"
 * are emitted (src/private/fnft__errwarn.c:28-45); NULL silences them.
 */
#include "fnft_internal.h"
#include <stdarg.h>
#include <stdio.h>

static FNFT_INT default_printf(const char *format, ...)
{
    va_list args;
    va_start(args, format);
    const FNFT_INT n = vfprintf(stderr, format, args);
    va_end(args);
    return n;
}

static __thread fnft_printf_ptr_t tl_printf = default_printf;

void fnft_errwarn_setprintf(fnft_printf_ptr_t printf_ptr) { tl_printf = printf_ptr; }

fnft_printf_ptr_t fnft_errwarn_getprintf(void) { return tl_printf; }

FNFT_INT fnftb__errmsg(const FNFT_INT ec, const char *func, const FNFT_INT line, const char *msg)
{
    if (tl_printf != NULL)
        tl_printf("FNFT Error: %s\n in %s(%i)-%d.%d.%d%s\n", msg, func, line, FNFT_VERSION_MAJOR,
                  FNFT_VERSION_MINOR, FNFT_VERSION_PATCH, FNFT_VERSION_SUFFIX);
    return ec;
}

void fnftb__warn(const char *func, const FNFT_INT line, const char *msg)
{
    if (tl_printf != NULL)
        tl_printf("FNFT Warning: %s\n in %s(%i)-%d.%d.%d%s\n", msg, func, line, FNFT_VERSION_MAJOR,
                  FNFT_VERSION_MINOR, FNFT_VERSION_PATCH, FNFT_VERSION_SUFFIX);
}

/* The names the reference's private error macros expand to (include/private/fnft__errwarn.h:36,125,
 * src/private/fnft__errwarn.c:28-45): test programs compiled against the reference's headers call them. */
FNFT_INT fnft__errmsg_aux(const FNFT_INT ec, const char *func, const FNFT_INT line, const char *msg)
{
    return fnftb__errmsg(ec, func, line, msg);
}

void fnft__warn_aux(const char *func, const FNFT_INT line, const char *msg) { fnftb__warn(func, line, msg); }

FNFT_INT fnftb__device_error(const char *func, const FNFT_INT line)
{
    char buf[640];
    snprintf(buf, sizeof(buf), "GPU path failed: %s", fnftb_last_error());
    return fnftb__errmsg(FNFT_EC_OTHER, func, line, buf);
}

/* include/fnft_version.h:47, src/fnft_version.c:27-45 */
FNFT_INT fnft_version(FNFT_UINT *major, FNFT_UINT *minor, FNFT_UINT *patch,
                      char suffix[FNFT_VERSION_SUFFIX_MAXLEN + 1])
{
    if (major == NULL)
        return E_INVALID_ARGUMENT(major);
    if (minor == NULL)
        return E_INVALID_ARGUMENT(minor);
    if (patch == NULL)
        return E_INVALID_ARGUMENT(patch);
    if (suffix == NULL)
        return E_INVALID_ARGUMENT(suffix);
    *major = FNFT_VERSION_MAJOR;
    *minor = FNFT_VERSION_MINOR;
    *patch = FNFT_VERSION_PATCH;
    strncpy(suffix, FNFT_VERSION_SUFFIX, FNFT_VERSION_SUFFIX_MAXLEN);
    suffix[FNFT_VERSION_SUFFIX_MAXLEN] = '\0';
    return FNFT_SUCCESS;
}
