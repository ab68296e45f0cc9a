/* glib.h -- minimal from-scratch stand-in for the parts of GLib that the
 * reference's canvas path uses.  TEST INFRASTRUCTURE ONLY (oracle build).
 *
 * The reference (hpjansson/chafa 1.19.0) requires GLib >= 2.58, which is not
 * installed in this image.  This header + glibshim.c provide just enough of
 * the API (about 95 identifiers) for the reference's own C files to compile
 * unmodified from /root/reference into oracle/_ref/libchafa_ref.so.
 *
 * Three behaviours of real GLib leak into the reference's *results* and are
 * pinned here on purpose (SURVEY.md 7.3 a,b):
 *   - GHashTable iteration order for g_direct_hash keys (decides the symbol
 *     table order and therefore every tie-break): implemented as GLib >= 2.58
 *     documents it -- open addressing, (hash * 11) % prime, triangular
 *     probing, iteration by slot index, growth at 15/16 occupancy to
 *     nnodes * 1.333;
 *   - Unicode property tables behind g_unichar_* (decide symbol membership):
 *     Unicode 15.0 + the Unicode 16 legacy-computing supplement;
 *   - nothing else: threads are pthreads, memory is malloc.
 */
#ifndef GLIBSHIM_GLIB_H
#define GLIBSHIM_GLIB_H

#include <stddef.h>
#include <stdint.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <limits.h>
#include <float.h>
#include <assert.h>

#ifdef __cplusplus
# define G_BEGIN_DECLS extern "C" {
# define G_END_DECLS }
#else
# define G_BEGIN_DECLS
# define G_END_DECLS
#endif

G_BEGIN_DECLS

typedef char gchar;
typedef short gshort;
typedef long glong;
typedef int gint;
typedef gint gboolean;
typedef unsigned char guchar;
typedef unsigned short gushort;
typedef unsigned long gulong;
typedef unsigned int guint;
typedef float gfloat;
typedef double gdouble;
typedef int8_t gint8;
typedef uint8_t guint8;
typedef int16_t gint16;
typedef uint16_t guint16;
typedef int32_t gint32;
typedef uint32_t guint32;
typedef int64_t gint64;
typedef uint64_t guint64;
typedef size_t gsize;
typedef ptrdiff_t gssize;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef guint32 gunichar;
typedef guint32 GQuark;
typedef intptr_t gintptr;
typedef uintptr_t guintptr;

#ifndef TRUE
# define TRUE 1
#endif
#ifndef FALSE
# define FALSE 0
#endif
#ifndef NULL
# define NULL ((void *) 0)
#endif

#undef MIN
#undef MAX
#undef ABS
#undef CLAMP
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#define ABS(a) (((a) < 0) ? -(a) : (a))
#define CLAMP(x, low, high) (((x) > (high)) ? (high) : (((x) < (low)) ? (low) : (x)))

#define G_N_ELEMENTS(arr) (sizeof (arr) / sizeof ((arr) [0]))
#define G_STRINGIFY(x) G_STRINGIFY_ARG (x)
#define G_STRINGIFY_ARG(x) #x

#define G_MININT8 ((gint8) -128)
#define G_MAXINT8 ((gint8) 127)
#define G_MAXUINT8 ((guint8) 255)
#define G_MININT16 ((gint16) -32768)
#define G_MAXINT16 ((gint16) 32767)
#define G_MAXUINT16 ((guint16) 65535)
#define G_MININT32 ((gint32) INT32_MIN)
#define G_MAXINT32 ((gint32) INT32_MAX)
#define G_MAXUINT32 ((guint32) UINT32_MAX)
#define G_MININT64 ((gint64) INT64_MIN)
#define G_MAXINT64 ((gint64) INT64_MAX)
#define G_MAXUINT64 ((guint64) UINT64_MAX)
#define G_MININT INT_MIN
#define G_MAXINT INT_MAX
#define G_MAXUINT UINT_MAX
#define G_MAXSIZE SIZE_MAX
#define G_MAXFLOAT FLT_MAX
#define G_MINFLOAT FLT_MIN
#define G_MAXDOUBLE DBL_MAX

#define G_GINT64_CONSTANT(v) (v##L)
#define G_GUINT64_CONSTANT(v) (v##UL)
#define G_GSIZE_FORMAT "lu"
#define G_GINT64_FORMAT "ld"
#define G_GUINT64_FORMAT "lu"

#define GPOINTER_TO_INT(p) ((gint) (glong) (p))
#define GPOINTER_TO_UINT(p) ((guint) (gulong) (p))
#define GINT_TO_POINTER(i) ((gpointer) (glong) (i))
#define GUINT_TO_POINTER(u) ((gpointer) (gulong) (u))
#define GSIZE_TO_POINTER(s) ((gpointer) (gsize) (s))
#define GPOINTER_TO_SIZE(p) ((gsize) (p))

#define G_GNUC_PURE __attribute__ ((__pure__))
#define G_GNUC_CONST __attribute__ ((__const__))
#define G_GNUC_UNUSED __attribute__ ((__unused__))
#define G_GNUC_MALLOC __attribute__ ((__malloc__))
#define G_GNUC_NORETURN __attribute__ ((__noreturn__))
#define G_GNUC_WARN_UNUSED_RESULT __attribute__ ((warn_unused_result))
#define G_GNUC_PRINTF(f, a) __attribute__ ((__format__ (__printf__, f, a)))
#define G_GNUC_INTERNAL __attribute__ ((visibility ("hidden")))
#define G_GNUC_NULL_TERMINATED __attribute__ ((__sentinel__))
#define G_DEPRECATED __attribute__ ((__deprecated__))
#define G_DEPRECATED_FOR(f) __attribute__ ((__deprecated__ ("Use '" #f "' instead")))
#define G_UNAVAILABLE(maj, min) __attribute__ ((deprecated ("Not available before " #maj "." #min)))
#define G_LIKELY(e) (__builtin_expect (!!(e), 1))
#define G_UNLIKELY(e) (__builtin_expect (!!(e), 0))
#define G_STMT_START do
#define G_STMT_END while (0)
#define G_STATIC_ASSERT(e) _Static_assert (e, "static assertion failed")
#define G_STRFUNC ((const char *) (__func__))
#define G_ENCODE_VERSION(major, minor) ((major) << 16 | (minor) << 8)
#define G_INLINE_FUNC static inline
#define G_GNUC_BEGIN_IGNORE_DEPRECATIONS _Pragma ("GCC diagnostic push") _Pragma ("GCC diagnostic ignored \"-Wdeprecated-declarations\"")
#define G_GNUC_END_IGNORE_DEPRECATIONS _Pragma ("GCC diagnostic pop")
#define G_DIR_SEPARATOR '/'
#define G_DIR_SEPARATOR_S "/"
#define G_OS_UNIX 1
#define G_BYTE_ORDER 1234
#define G_LITTLE_ENDIAN 1234
#define G_BIG_ENDIAN 4321
#define GLIB_CHECK_VERSION(a, b, c) 1
#define GUINT32_TO_BE(v) (__builtin_bswap32 ((guint32) (v)))
#define GUINT32_FROM_BE(v) (__builtin_bswap32 ((guint32) (v)))
#define GUINT16_TO_BE(v) (__builtin_bswap16 ((guint16) (v)))
#define GUINT16_FROM_BE(v) (__builtin_bswap16 ((guint16) (v)))
#define GUINT32_TO_LE(v) ((guint32) (v))
#define GUINT32_FROM_LE(v) ((guint32) (v))
#define GUINT16_TO_LE(v) ((guint16) (v))
#define GUINT16_FROM_LE(v) ((guint16) (v))
#define GUINT64_TO_BE(v) (__builtin_bswap64 ((guint64) (v)))

typedef void (*GFunc) (gpointer data, gpointer user_data);
typedef void (*GDestroyNotify) (gpointer data);
typedef guint (*GHashFunc) (gconstpointer key);
typedef gboolean (*GEqualFunc) (gconstpointer a, gconstpointer b);
typedef gint (*GCompareFunc) (gconstpointer a, gconstpointer b);
typedef gpointer (*GThreadFunc) (gpointer data);

/* ---- messages / assertions ---- */

void glibshim_log (const char *level, const char *fmt, ...) G_GNUC_PRINTF (2, 3);
void g_printerr (const gchar *fmt, ...) G_GNUC_PRINTF (1, 2);
void g_print (const gchar *fmt, ...) G_GNUC_PRINTF (1, 2);

#define g_warning(...) glibshim_log ("WARNING", __VA_ARGS__)
#define g_critical(...) glibshim_log ("CRITICAL", __VA_ARGS__)
#define g_message(...) glibshim_log ("MESSAGE", __VA_ARGS__)
#define g_debug(...) do { } while (0)
#define g_error(...) do { glibshim_log ("ERROR", __VA_ARGS__); abort (); } while (0)

#define g_assert(e) do { if (G_UNLIKELY (!(e))) { \
    fprintf (stderr, "%s:%d: assertion failed: %s\n", __FILE__, __LINE__, #e); abort (); } } while (0)
#define g_assert_not_reached() do { \
    fprintf (stderr, "%s:%d: should not be reached\n", __FILE__, __LINE__); abort (); } while (0)
#define g_assert_true(e) g_assert (e)
#define g_assert_false(e) g_assert (!(e))
#define g_assert_nonnull(e) g_assert ((e) != NULL)
#define g_assert_null(e) g_assert ((e) == NULL)

#define g_return_if_fail(e) do { if (G_UNLIKELY (!(e))) { \
    glibshim_log ("CRITICAL", "%s: assertion '%s' failed", __func__, #e); return; } } while (0)
#define g_return_val_if_fail(e, v) do { if (G_UNLIKELY (!(e))) { \
    glibshim_log ("CRITICAL", "%s: assertion '%s' failed", __func__, #e); return (v); } } while (0)
#define g_return_if_reached() do { return; } while (0)
#define g_return_val_if_reached(v) do { return (v); } while (0)

/* ---- memory ---- */

gpointer g_malloc (gsize n);
gpointer g_malloc0 (gsize n);
gpointer g_realloc (gpointer p, gsize n);
gpointer g_try_malloc (gsize n);
gpointer g_try_malloc0 (gsize n);
gpointer g_try_realloc (gpointer p, gsize n);
void g_free (gpointer p);
gpointer g_memdup (gconstpointer p, guint n);
gpointer g_memdup2 (gconstpointer p, gsize n);

#define g_new(type, n) ((type *) g_malloc (sizeof (type) * (gsize) (n)))
#define g_new0(type, n) ((type *) g_malloc0 (sizeof (type) * (gsize) (n)))
#define g_renew(type, p, n) ((type *) g_realloc ((p), sizeof (type) * (gsize) (n)))
#define g_try_new(type, n) ((type *) g_try_malloc (sizeof (type) * (gsize) (n)))
#define g_try_new0(type, n) ((type *) g_try_malloc0 (sizeof (type) * (gsize) (n)))
#define g_alloca(n) __builtin_alloca (n)
#define g_newa(type, n) ((type *) g_alloca (sizeof (type) * (gsize) (n)))
#define g_clear_pointer(pp, destroy) do { if (*(pp)) { (destroy) (*(pp)); *(pp) = NULL; } } while (0)

/* ---- strings ---- */

gchar *g_strdup (const gchar *s);
gchar *g_strndup (const gchar *s, gsize n);
gchar *g_strdup_printf (const gchar *fmt, ...) G_GNUC_PRINTF (1, 2);
gchar *g_strjoin (const gchar *sep, ...) G_GNUC_NULL_TERMINATED;
gchar *g_strconcat (const gchar *first, ...) G_GNUC_NULL_TERMINATED;
gchar **g_strsplit (const gchar *s, const gchar *delim, gint max_tokens);
void g_strfreev (gchar **v);
guint g_strv_length (gchar **v);
gint g_snprintf (gchar *buf, gulong n, const gchar *fmt, ...) G_GNUC_PRINTF (3, 4);
gint g_ascii_strcasecmp (const gchar *a, const gchar *b);
gint g_ascii_strncasecmp (const gchar *a, const gchar *b, gsize n);
gchar g_ascii_tolower (gchar c);
gchar g_ascii_toupper (gchar c);
gboolean g_ascii_isdigit_f (gchar c);
gboolean g_str_has_prefix (const gchar *s, const gchar *prefix);
gboolean g_str_has_suffix (const gchar *s, const gchar *suffix);
gint g_strcmp0 (const gchar *a, const gchar *b);
guint64 g_ascii_strtoull (const gchar *nptr, gchar **endptr, guint base);
gint64 g_ascii_strtoll (const gchar *nptr, gchar **endptr, guint base);
#define g_ascii_isdigit(c) ((c) >= '0' && (c) <= '9')
#define g_ascii_isspace(c) ((c) == ' ' || ((c) >= '\t' && (c) <= '\r'))
#define g_ascii_isalpha(c) ((((c) | 0x20) >= 'a') && (((c) | 0x20) <= 'z'))
#define g_ascii_isalnum(c) (g_ascii_isalpha (c) || g_ascii_isdigit (c))
#define g_ascii_isxdigit(c) (g_ascii_isdigit (c) || ((((c) | 0x20) >= 'a') && (((c) | 0x20) <= 'f')))

typedef struct
{
    gchar *str;
    gsize len;
    gsize allocated_len;
}
GString;

GString *g_string_new (const gchar *init);
GString *g_string_new_len (const gchar *init, gssize len);
GString *g_string_sized_new (gsize dfl_size);
gchar *g_string_free (GString *string, gboolean free_segment);
gchar *g_string_free_and_steal (GString *string);
GString *g_string_append (GString *string, const gchar *val);
GString *g_string_append_len (GString *string, const gchar *val, gssize len);
GString *g_string_append_c (GString *string, gchar c);
GString *g_string_append_unichar (GString *string, gunichar wc);
GString *g_string_append_printf (GString *string, const gchar *fmt, ...) G_GNUC_PRINTF (2, 3);
void g_string_printf (GString *string, const gchar *fmt, ...) G_GNUC_PRINTF (2, 3);
GString *g_string_truncate (GString *string, gsize len);
GString *g_string_set_size (GString *string, gsize len);
GString *g_string_assign (GString *string, const gchar *rval);
GString *g_string_prepend (GString *string, const gchar *val);
GString *g_string_erase (GString *string, gssize pos, gssize len);

/* ---- unicode ---- */

typedef enum
{
    G_UNICODE_SCRIPT_INVALID_CODE = -1,
    G_UNICODE_SCRIPT_COMMON = 0,
    G_UNICODE_SCRIPT_INHERITED,
    G_UNICODE_SCRIPT_ARABIC,
    G_UNICODE_SCRIPT_ARMENIAN,
    G_UNICODE_SCRIPT_BENGALI,
    G_UNICODE_SCRIPT_BOPOMOFO,
    G_UNICODE_SCRIPT_CHEROKEE,
    G_UNICODE_SCRIPT_COPTIC,
    G_UNICODE_SCRIPT_CYRILLIC,
    G_UNICODE_SCRIPT_DESERET,
    G_UNICODE_SCRIPT_DEVANAGARI,
    G_UNICODE_SCRIPT_ETHIOPIC,
    G_UNICODE_SCRIPT_GEORGIAN,
    G_UNICODE_SCRIPT_GOTHIC,
    G_UNICODE_SCRIPT_GREEK,
    G_UNICODE_SCRIPT_GUJARATI,
    G_UNICODE_SCRIPT_GURMUKHI,
    G_UNICODE_SCRIPT_HAN,
    G_UNICODE_SCRIPT_HANGUL,
    G_UNICODE_SCRIPT_HEBREW,
    G_UNICODE_SCRIPT_HIRAGANA,
    G_UNICODE_SCRIPT_KANNADA,
    G_UNICODE_SCRIPT_KATAKANA,
    G_UNICODE_SCRIPT_KHMER,
    G_UNICODE_SCRIPT_LAO,
    G_UNICODE_SCRIPT_LATIN,
    G_UNICODE_SCRIPT_MALAYALAM,
    G_UNICODE_SCRIPT_MONGOLIAN,
    G_UNICODE_SCRIPT_MYANMAR,
    G_UNICODE_SCRIPT_OGHAM,
    G_UNICODE_SCRIPT_OLD_ITALIC,
    G_UNICODE_SCRIPT_ORIYA,
    G_UNICODE_SCRIPT_RUNIC,
    G_UNICODE_SCRIPT_SINHALA,
    G_UNICODE_SCRIPT_SYRIAC,
    G_UNICODE_SCRIPT_TAMIL,
    G_UNICODE_SCRIPT_TELUGU,
    G_UNICODE_SCRIPT_THAANA,
    G_UNICODE_SCRIPT_THAI,
    G_UNICODE_SCRIPT_TIBETAN,
    G_UNICODE_SCRIPT_UNKNOWN
}
GUnicodeScript;

gboolean g_unichar_isprint (gunichar c);
gboolean g_unichar_isalpha (gunichar c);
gboolean g_unichar_isdigit (gunichar c);
gboolean g_unichar_ismark (gunichar c);
gboolean g_unichar_iszerowidth (gunichar c);
gboolean g_unichar_iswide (gunichar c);
gboolean g_unichar_iswide_cjk (gunichar c);
GUnicodeScript g_unichar_get_script (gunichar c);
gint g_unichar_to_utf8 (gunichar c, gchar *outbuf);
gunichar g_utf8_get_char (const gchar *p);
gunichar g_utf8_get_char_validated (const gchar *p, gssize max_len);
extern const gchar *const g_utf8_skip;
#define g_utf8_next_char(p) ((p) + g_utf8_skip [*(const guchar *) (p)])
glong g_utf8_strlen (const gchar *p, gssize max);

/* ---- errors ---- */

typedef struct
{
    GQuark domain;
    gint code;
    gchar *message;
}
GError;

GQuark g_quark_from_static_string (const gchar *s);
void g_set_error (GError **err, GQuark domain, gint code, const gchar *fmt, ...) G_GNUC_PRINTF (4, 5);
void g_set_error_literal (GError **err, GQuark domain, gint code, const gchar *message);
void g_error_free (GError *err);
void g_clear_error (GError **err);
void g_propagate_error (GError **dest, GError *src);

#define G_DEFINE_QUARK(QN, q_n) \
    GQuark q_n##_quark (void) { static GQuark q; if (!q) q = g_quark_from_static_string (#QN); return q; }

GQuark g_option_error_quark (void);
#define G_OPTION_ERROR (g_option_error_quark ())
typedef enum
{
    G_OPTION_ERROR_UNKNOWN_OPTION,
    G_OPTION_ERROR_BAD_VALUE,
    G_OPTION_ERROR_FAILED
}
GOptionError;

/* ---- atomics ---- */

#define g_atomic_int_get(p) (__atomic_load_n ((p), __ATOMIC_SEQ_CST))
#define g_atomic_int_set(p, v) (__atomic_store_n ((p), (v), __ATOMIC_SEQ_CST))
#define g_atomic_int_inc(p) ((void) __atomic_fetch_add ((p), 1, __ATOMIC_SEQ_CST))
#define g_atomic_int_dec_and_test(p) (__atomic_fetch_sub ((p), 1, __ATOMIC_SEQ_CST) == 1)
#define g_atomic_int_add(p, v) (__atomic_fetch_add ((p), (v), __ATOMIC_SEQ_CST))
#define g_atomic_int_compare_and_exchange(p, o, n) \
    (__extension__ ({ gint _o = (o); __atomic_compare_exchange_n ((p), &_o, (n), FALSE, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); }))
#define g_atomic_pointer_get(p) (__atomic_load_n ((p), __ATOMIC_SEQ_CST))
#define g_atomic_pointer_set(p, v) (__atomic_store_n ((p), (v), __ATOMIC_SEQ_CST))

/* ---- once ---- */

typedef enum
{
    G_ONCE_STATUS_NOTCALLED,
    G_ONCE_STATUS_PROGRESS,
    G_ONCE_STATUS_READY
}
GOnceStatus;

typedef struct
{
    volatile GOnceStatus status;
    volatile gpointer retval;
}
GOnce;

#define G_ONCE_INIT { G_ONCE_STATUS_NOTCALLED, NULL }
gpointer g_once_impl (GOnce *once, GThreadFunc func, gpointer arg);
#define g_once(once, func, arg) \
    (((once)->status == G_ONCE_STATUS_READY) ? (once)->retval : g_once_impl ((once), (func), (arg)))

/* ---- mutex (tiny) ---- */

typedef struct { void *impl [8]; } GMutex;
typedef struct { void *impl [8]; } GCond;
void g_mutex_init (GMutex *m);
void g_mutex_clear (GMutex *m);
void g_mutex_lock (GMutex *m);
void g_mutex_unlock (GMutex *m);
void g_cond_init (GCond *c);
void g_cond_clear (GCond *c);
void g_cond_wait (GCond *c, GMutex *m);
void g_cond_signal (GCond *c);
void g_cond_broadcast (GCond *c);

/* ---- thread pool ---- */

typedef struct _GThreadPool GThreadPool;
GThreadPool *g_thread_pool_new (GFunc func, gpointer user_data, gint max_threads,
                                gboolean exclusive, GError **error);
gboolean g_thread_pool_push (GThreadPool *pool, gpointer data, GError **error);
void g_thread_pool_free (GThreadPool *pool, gboolean immediate, gboolean wait_);
guint g_get_num_processors (void);

/* ---- environment ---- */

gchar **g_get_environ (void);
const gchar *g_environ_getenv (gchar **envp, const gchar *variable);
const gchar *g_getenv (const gchar *variable);

/* ---- hash table ---- */

typedef struct _GHashTable GHashTable;
typedef struct
{
    gpointer dummy1;
    gpointer dummy2;
    gpointer dummy3;
    int dummy4;
    gboolean dummy5;
    gpointer dummy6;
}
GHashTableIter;

guint g_direct_hash (gconstpointer v);
gboolean g_direct_equal (gconstpointer a, gconstpointer b);
guint g_str_hash (gconstpointer v);
gboolean g_str_equal (gconstpointer a, gconstpointer b);
GHashTable *g_hash_table_new (GHashFunc hash_func, GEqualFunc key_equal_func);
GHashTable *g_hash_table_new_full (GHashFunc hash_func, GEqualFunc key_equal_func,
                                   GDestroyNotify key_destroy_func,
                                   GDestroyNotify value_destroy_func);
void g_hash_table_destroy (GHashTable *ht);
void g_hash_table_unref (GHashTable *ht);
gboolean g_hash_table_insert (GHashTable *ht, gpointer key, gpointer value);
gboolean g_hash_table_replace (GHashTable *ht, gpointer key, gpointer value);
gboolean g_hash_table_remove (GHashTable *ht, gconstpointer key);
void g_hash_table_remove_all (GHashTable *ht);
gpointer g_hash_table_lookup (GHashTable *ht, gconstpointer key);
gboolean g_hash_table_contains (GHashTable *ht, gconstpointer key);
guint g_hash_table_size (GHashTable *ht);
void g_hash_table_iter_init (GHashTableIter *iter, GHashTable *ht);
gboolean g_hash_table_iter_next (GHashTableIter *iter, gpointer *key, gpointer *value);

/* ---- array ---- */

typedef struct
{
    gchar *data;
    guint len;
}
GArray;

GArray *g_array_new (gboolean zero_terminated, gboolean clear_, guint element_size);
GArray *g_array_sized_new (gboolean zero_terminated, gboolean clear_, guint element_size, guint reserved);
gchar *g_array_free (GArray *array, gboolean free_segment);
GArray *g_array_append_vals (GArray *array, gconstpointer data, guint len);
GArray *g_array_set_size (GArray *array, guint length);
GArray *g_array_remove_index (GArray *array, guint index_);
void g_array_unref (GArray *array);
#define g_array_append_val(a, v) g_array_append_vals (a, &(v), 1)
#define g_array_index(a, t, i) (((t *) (void *) (a)->data) [(i)])

/* ---- poll fd (only referenced in headers that are out of scope) ---- */

typedef struct
{
    gint fd;
    gushort events;
    gushort revents;
}
GPollFD;

gint64 g_get_monotonic_time (void);
gint64 g_get_real_time (void);
void g_usleep (gulong microseconds);

G_END_DECLS

#endif /* GLIBSHIM_GLIB_H */
