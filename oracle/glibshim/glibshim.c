/* glibshim.c -- implementation of the GLib stand-in declared in glib.h.
 * TEST INFRASTRUCTURE ONLY: linked into oracle/_ref/libchafa_ref.so next to the
 * reference's own, unmodified C files.  Written from scratch from GLib's
 * documented behaviour; see glib.h for what is pinned and why. */

#include "glib.h"
#include <pthread.h>
#include <unistd.h>
#include <time.h>
#include <ctype.h>

#include "unicode_ranges.inc"

extern char **environ;

/* ------------------------------------------------------------------ */
/* messages                                                            */

void
glibshim_log (const char *level, const char *fmt, ...)
{
    va_list ap;

    /* Criticals from g_return_if_fail are part of normal test flow (the
     * reference's API has no other error channel); keep them quiet unless asked. */
    if (!getenv ("GLIBSHIM_VERBOSE") && strcmp (level, "ERROR"))
        return;

    fprintf (stderr, "** %s **: ", level);
    va_start (ap, fmt);
    vfprintf (stderr, fmt, ap);
    va_end (ap);
    fputc ('\n', stderr);
}

void
g_printerr (const gchar *fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    vfprintf (stderr, fmt, ap);
    va_end (ap);
}

void
g_print (const gchar *fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    vfprintf (stdout, fmt, ap);
    va_end (ap);
}

/* ------------------------------------------------------------------ */
/* memory                                                              */

gpointer
g_malloc (gsize n)
{
    gpointer p;
    if (!n)
        return NULL;
    p = malloc (n);
    if (!p)
    {
        fprintf (stderr, "glibshim: out of memory (%zu bytes)\n", n);
        abort ();
    }
    return p;
}

gpointer
g_malloc0 (gsize n)
{
    gpointer p;
    if (!n)
        return NULL;
    p = calloc (1, n);
    if (!p)
    {
        fprintf (stderr, "glibshim: out of memory (%zu bytes)\n", n);
        abort ();
    }
    return p;
}

gpointer
g_realloc (gpointer p, gsize n)
{
    gpointer q;
    if (!n)
    {
        free (p);
        return NULL;
    }
    q = realloc (p, n);
    if (!q)
    {
        fprintf (stderr, "glibshim: out of memory (%zu bytes)\n", n);
        abort ();
    }
    return q;
}

gpointer g_try_malloc (gsize n) { return n ? malloc (n) : NULL; }
gpointer g_try_malloc0 (gsize n) { return n ? calloc (1, n) : NULL; }
gpointer g_try_realloc (gpointer p, gsize n) { if (!n) { free (p); return NULL; } return realloc (p, n); }
void g_free (gpointer p) { free (p); }

gpointer
g_memdup2 (gconstpointer p, gsize n)
{
    gpointer q;
    if (!p || !n)
        return NULL;
    q = g_malloc (n);
    memcpy (q, p, n);
    return q;
}

gpointer g_memdup (gconstpointer p, guint n) { return g_memdup2 (p, n); }

/* ------------------------------------------------------------------ */
/* strings                                                             */

gchar *
g_strdup (const gchar *s)
{
    if (!s)
        return NULL;
    return g_memdup2 (s, strlen (s) + 1);
}

gchar *
g_strndup (const gchar *s, gsize n)
{
    gchar *r;
    gsize len;
    if (!s)
        return NULL;
    len = strnlen (s, n);
    r = g_malloc (n + 1);
    memcpy (r, s, len);
    memset (r + len, 0, n + 1 - len);
    return r;
}

gchar *
g_strdup_printf (const gchar *fmt, ...)
{
    va_list ap;
    char *out = NULL;
    va_start (ap, fmt);
    if (vasprintf (&out, fmt, ap) < 0)
        out = NULL;
    va_end (ap);
    return out;
}

gchar *
g_strjoin (const gchar *sep, ...)
{
    va_list ap;
    GString *gs = g_string_new ("");
    const gchar *s;
    gboolean first = TRUE;

    if (!sep)
        sep = "";
    va_start (ap, sep);
    while ((s = va_arg (ap, const gchar *)))
    {
        if (!first)
            g_string_append (gs, sep);
        g_string_append (gs, s);
        first = FALSE;
    }
    va_end (ap);
    return g_string_free (gs, FALSE);
}

gchar *
g_strconcat (const gchar *first, ...)
{
    va_list ap;
    GString *gs;
    const gchar *s;

    if (!first)
        return NULL;
    gs = g_string_new (first);
    va_start (ap, first);
    while ((s = va_arg (ap, const gchar *)))
        g_string_append (gs, s);
    va_end (ap);
    return g_string_free (gs, FALSE);
}

gchar **
g_strsplit (const gchar *s, const gchar *delim, gint max_tokens)
{
    gchar **v;
    gsize n = 0, cap = 8, dl = strlen (delim);
    const gchar *p;

    if (max_tokens < 1)
        max_tokens = G_MAXINT;
    v = g_new (gchar *, cap);

    if (*s)
    {
        while ((p = strstr (s, delim)) && (gint) n + 1 < max_tokens)
        {
            if (n + 2 > cap)
                v = g_renew (gchar *, v, cap *= 2);
            v [n++] = g_strndup (s, p - s);
            s = p + dl;
        }
        if (n + 2 > cap)
            v = g_renew (gchar *, v, cap *= 2);
        v [n++] = g_strdup (s);
    }
    v [n] = NULL;
    return v;
}

void
g_strfreev (gchar **v)
{
    gchar **p;
    if (!v)
        return;
    for (p = v; *p; p++)
        g_free (*p);
    g_free (v);
}

guint
g_strv_length (gchar **v)
{
    guint n = 0;
    while (v && v [n])
        n++;
    return n;
}

gint
g_snprintf (gchar *buf, gulong n, const gchar *fmt, ...)
{
    va_list ap;
    gint r;
    va_start (ap, fmt);
    r = vsnprintf (buf, n, fmt, ap);
    va_end (ap);
    return r;
}

gchar g_ascii_tolower (gchar c) { return (c >= 'A' && c <= 'Z') ? c - 'A' + 'a' : c; }
gchar g_ascii_toupper (gchar c) { return (c >= 'a' && c <= 'z') ? c - 'a' + 'A' : c; }

gint
g_ascii_strcasecmp (const gchar *a, const gchar *b)
{
    while (*a && *b)
    {
        gint ca = (guchar) g_ascii_tolower (*a), cb = (guchar) g_ascii_tolower (*b);
        if (ca != cb)
            return ca - cb;
        a++; b++;
    }
    return ((gint) (guchar) *a) - ((gint) (guchar) *b);
}

gint
g_ascii_strncasecmp (const gchar *a, const gchar *b, gsize n)
{
    while (n && *a && *b)
    {
        gint ca = (guchar) g_ascii_tolower (*a), cb = (guchar) g_ascii_tolower (*b);
        if (ca != cb)
            return ca - cb;
        a++; b++; n--;
    }
    if (!n)
        return 0;
    return ((gint) (guchar) *a) - ((gint) (guchar) *b);
}

gboolean
g_str_has_prefix (const gchar *s, const gchar *prefix)
{
    return strncmp (s, prefix, strlen (prefix)) == 0;
}

gboolean
g_str_has_suffix (const gchar *s, const gchar *suffix)
{
    gsize a = strlen (s), b = strlen (suffix);
    return a >= b && !strcmp (s + a - b, suffix);
}

gint
g_strcmp0 (const gchar *a, const gchar *b)
{
    if (!a)
        return -(a != b);
    if (!b)
        return a != b;
    return strcmp (a, b);
}

guint64 g_ascii_strtoull (const gchar *nptr, gchar **endptr, guint base) { return strtoull (nptr, endptr, base); }
gint64 g_ascii_strtoll (const gchar *nptr, gchar **endptr, guint base) { return strtoll (nptr, endptr, base); }

/* GString */

static void
gstring_expand (GString *s, gsize extra)
{
    gsize need = s->len + extra + 1;
    if (need > s->allocated_len)
    {
        gsize n = s->allocated_len ? s->allocated_len : 16;
        while (n < need)
            n *= 2;
        s->str = g_realloc (s->str, n);
        s->allocated_len = n;
    }
}

GString *
g_string_sized_new (gsize dfl_size)
{
    GString *s = g_new0 (GString, 1);
    gstring_expand (s, MAX (dfl_size, 2));
    s->str [0] = 0;
    return s;
}

GString *
g_string_new (const gchar *init)
{
    GString *s = g_string_sized_new (init ? strlen (init) + 2 : 2);
    if (init)
        g_string_append (s, init);
    return s;
}

GString *
g_string_new_len (const gchar *init, gssize len)
{
    GString *s;
    if (len < 0)
        return g_string_new (init);
    s = g_string_sized_new (len);
    if (init)
        g_string_append_len (s, init, len);
    return s;
}

gchar *
g_string_free (GString *string, gboolean free_segment)
{
    gchar *r;
    if (!string)
        return NULL;
    r = string->str;
    if (free_segment)
    {
        g_free (r);
        r = NULL;
    }
    g_free (string);
    return r;
}

gchar *g_string_free_and_steal (GString *string) { return g_string_free (string, FALSE); }

GString *
g_string_append_len (GString *string, const gchar *val, gssize len)
{
    if (len < 0)
        len = strlen (val);
    gstring_expand (string, len);
    memcpy (string->str + string->len, val, len);
    string->len += len;
    string->str [string->len] = 0;
    return string;
}

GString *g_string_append (GString *string, const gchar *val) { return g_string_append_len (string, val, -1); }

GString *
g_string_append_c (GString *string, gchar c)
{
    gstring_expand (string, 1);
    string->str [string->len++] = c;
    string->str [string->len] = 0;
    return string;
}

GString *
g_string_append_unichar (GString *string, gunichar wc)
{
    gchar buf [8];
    gint n = g_unichar_to_utf8 (wc, buf);
    return g_string_append_len (string, buf, n);
}

GString *
g_string_append_printf (GString *string, const gchar *fmt, ...)
{
    va_list ap;
    char *out = NULL;
    va_start (ap, fmt);
    if (vasprintf (&out, fmt, ap) >= 0)
    {
        g_string_append (string, out);
        free (out);
    }
    va_end (ap);
    return string;
}

void
g_string_printf (GString *string, const gchar *fmt, ...)
{
    va_list ap;
    char *out = NULL;
    g_string_truncate (string, 0);
    va_start (ap, fmt);
    if (vasprintf (&out, fmt, ap) >= 0)
    {
        g_string_append (string, out);
        free (out);
    }
    va_end (ap);
}

GString *
g_string_truncate (GString *string, gsize len)
{
    string->len = MIN (len, string->len);
    string->str [string->len] = 0;
    return string;
}

GString *
g_string_set_size (GString *string, gsize len)
{
    if (len >= string->allocated_len)
        gstring_expand (string, len - string->len);
    string->len = len;
    string->str [len] = 0;
    return string;
}

GString *
g_string_assign (GString *string, const gchar *rval)
{
    g_string_truncate (string, 0);
    return g_string_append (string, rval);
}

GString *
g_string_prepend (GString *string, const gchar *val)
{
    gsize l = strlen (val);
    gstring_expand (string, l);
    memmove (string->str + l, string->str, string->len + 1);
    memcpy (string->str, val, l);
    string->len += l;
    return string;
}

GString *
g_string_erase (GString *string, gssize pos, gssize len)
{
    if (len < 0)
        len = string->len - pos;
    if ((gsize) (pos + len) < string->len)
        memmove (string->str + pos, string->str + pos + len, string->len - (pos + len));
    string->len -= len;
    string->str [string->len] = 0;
    return string;
}

/* ------------------------------------------------------------------ */
/* unicode                                                             */

static gboolean
in_ranges (gunichar c, const unsigned int (*r) [2], int n)
{
    int lo = 0, hi = n - 1;
    while (lo <= hi)
    {
        int mid = (lo + hi) / 2;
        if (c < r [mid] [0])
            hi = mid - 1;
        else if (c > r [mid] [1])
            lo = mid + 1;
        else
            return TRUE;
    }
    return FALSE;
}

gboolean g_unichar_isprint (gunichar c) { return in_ranges (c, ucd_print_ranges, UCD_PRINT_N); }
gboolean g_unichar_isalpha (gunichar c) { return in_ranges (c, ucd_alpha_ranges, UCD_ALPHA_N); }
gboolean g_unichar_isdigit (gunichar c) { return in_ranges (c, ucd_digit_ranges, UCD_DIGIT_N); }
gboolean g_unichar_ismark (gunichar c) { return in_ranges (c, ucd_mark_ranges, UCD_MARK_N); }

gboolean
g_unichar_iszerowidth (gunichar c)
{
    if (c == 0x00ad)
        return FALSE;
    if (in_ranges (c, ucd_zwtype_ranges, UCD_ZWTYPE_N))
        return TRUE;
    if ((c >= 0x1160 && c < 0x1200) || c == 0x200b)
        return TRUE;
    return FALSE;
}

gboolean g_unichar_iswide (gunichar c) { return in_ranges (c, ucd_wide_ranges, UCD_WIDE_N); }

gboolean
g_unichar_iswide_cjk (gunichar c)
{
    if (g_unichar_iswide (c))
        return TRUE;
    if (c == 0)
        return FALSE;
    return in_ranges (c, ucd_ambiguous_ranges, UCD_AMBIGUOUS_N);
}

GUnicodeScript
g_unichar_get_script (gunichar c)
{
    if (in_ranges (c, ucd_script_arabic_ranges, UCD_SCRIPT_ARABIC_N))
        return G_UNICODE_SCRIPT_ARABIC;
    if (in_ranges (c, ucd_script_hebrew_ranges, UCD_SCRIPT_HEBREW_N))
        return G_UNICODE_SCRIPT_HEBREW;
    if (in_ranges (c, ucd_script_syriac_ranges, UCD_SCRIPT_SYRIAC_N))
        return G_UNICODE_SCRIPT_SYRIAC;
    if (in_ranges (c, ucd_script_thaana_ranges, UCD_SCRIPT_THAANA_N))
        return G_UNICODE_SCRIPT_THAANA;
    return G_UNICODE_SCRIPT_COMMON;
}

gint
g_unichar_to_utf8 (gunichar c, gchar *outbuf)
{
    guint len;
    int first, i;

    if (c < 0x80) { first = 0; len = 1; }
    else if (c < 0x800) { first = 0xc0; len = 2; }
    else if (c < 0x10000) { first = 0xe0; len = 3; }
    else if (c < 0x200000) { first = 0xf0; len = 4; }
    else if (c < 0x4000000) { first = 0xf8; len = 5; }
    else { first = 0xfc; len = 6; }

    if (outbuf)
    {
        for (i = len - 1; i > 0; --i)
        {
            outbuf [i] = (c & 0x3f) | 0x80;
            c >>= 6;
        }
        outbuf [0] = c | first;
    }
    return len;
}

static const gchar utf8_skip_data [256] = {
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
    2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,
    3,3,3,3,3,3,3,3,3,3,3,3,3,3,3,3,4,4,4,4,4,4,4,4,5,5,5,5,6,6,1,1
};
const gchar *const g_utf8_skip = utf8_skip_data;

gunichar
g_utf8_get_char (const gchar *p)
{
    const guchar *s = (const guchar *) p;
    guchar c = s [0];
    gunichar r;
    int len, i;

    if (c < 0x80) return c;
    else if ((c & 0xe0) == 0xc0) { len = 2; r = c & 0x1f; }
    else if ((c & 0xf0) == 0xe0) { len = 3; r = c & 0x0f; }
    else if ((c & 0xf8) == 0xf0) { len = 4; r = c & 0x07; }
    else if ((c & 0xfc) == 0xf8) { len = 5; r = c & 0x03; }
    else if ((c & 0xfe) == 0xfc) { len = 6; r = c & 0x01; }
    else return (gunichar) -1;

    for (i = 1; i < len; i++)
    {
        if ((s [i] & 0xc0) != 0x80)
            return (gunichar) -1;
        r = (r << 6) | (s [i] & 0x3f);
    }
    return r;
}

gunichar
g_utf8_get_char_validated (const gchar *p, gssize max_len)
{
    const guchar *s = (const guchar *) p;
    gint need;
    gssize i;
    gunichar r;

    if (max_len == 0)
        return (gunichar) -2;
    if (s [0] < 0x80)
        return s [0];
    need = utf8_skip_data [s [0]];
    if (need == 1)
        return (gunichar) -1;
    if (max_len >= 0 && need > max_len)
    {
        for (i = 1; i < max_len; i++)
            if ((s [i] & 0xc0) != 0x80)
                return (gunichar) -1;
        return (gunichar) -2;
    }
    for (i = 1; i < need; i++)
        if ((s [i] & 0xc0) != 0x80)
            return (gunichar) -1;
    r = g_utf8_get_char (p);
    if (r > 0x10ffff || (r >= 0xd800 && r < 0xe000))
        return (gunichar) -1;
    return r;
}

glong
g_utf8_strlen (const gchar *p, gssize max)
{
    glong n = 0;
    const gchar *start = p;
    while (*p && (max < 0 || p - start < max))
    {
        p = g_utf8_next_char (p);
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------ */
/* errors                                                              */

GQuark
g_quark_from_static_string (const gchar *s)
{
    /* Not a real interning table; only used to tell error domains apart. */
    guint h = 5381;
    while (*s)
        h = h * 33 + (guchar) *s++;
    return h ? h : 1;
}

GQuark g_option_error_quark (void) { return g_quark_from_static_string ("g-option-context-error-quark"); }

void
g_set_error (GError **err, GQuark domain, gint code, const gchar *fmt, ...)
{
    va_list ap;
    char *msg = NULL;

    if (!err)
        return;
    va_start (ap, fmt);
    if (vasprintf (&msg, fmt, ap) < 0)
        msg = NULL;
    va_end (ap);
    if (*err)
    {
        free (msg);
        return;
    }
    *err = g_new0 (GError, 1);
    (*err)->domain = domain;
    (*err)->code = code;
    (*err)->message = msg;
}

void
g_set_error_literal (GError **err, GQuark domain, gint code, const gchar *message)
{
    g_set_error (err, domain, code, "%s", message);
}

void
g_error_free (GError *err)
{
    if (!err)
        return;
    g_free (err->message);
    g_free (err);
}

void
g_clear_error (GError **err)
{
    if (err && *err)
    {
        g_error_free (*err);
        *err = NULL;
    }
}

void
g_propagate_error (GError **dest, GError *src)
{
    if (dest && !*dest)
        *dest = src;
    else
        g_error_free (src);
}

/* ------------------------------------------------------------------ */
/* once, mutex, thread pool                                            */

static pthread_mutex_t once_mutex = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t once_cond = PTHREAD_COND_INITIALIZER;

gpointer
g_once_impl (GOnce *once, GThreadFunc func, gpointer arg)
{
    pthread_mutex_lock (&once_mutex);
    while (once->status == G_ONCE_STATUS_PROGRESS)
        pthread_cond_wait (&once_cond, &once_mutex);
    if (once->status != G_ONCE_STATUS_READY)
    {
        gpointer r;
        once->status = G_ONCE_STATUS_PROGRESS;
        pthread_mutex_unlock (&once_mutex);
        r = func (arg);
        pthread_mutex_lock (&once_mutex);
        once->retval = r;
        __atomic_thread_fence (__ATOMIC_SEQ_CST);
        once->status = G_ONCE_STATUS_READY;
        pthread_cond_broadcast (&once_cond);
    }
    pthread_mutex_unlock (&once_mutex);
    return once->retval;
}

void g_mutex_init (GMutex *m) { pthread_mutex_t *p = malloc (sizeof *p); pthread_mutex_init (p, NULL); m->impl [0] = p; }
void g_mutex_clear (GMutex *m) { pthread_mutex_destroy (m->impl [0]); free (m->impl [0]); m->impl [0] = NULL; }
static pthread_mutex_t *
mutex_get (GMutex *m)
{
    /* Statically zero-initialised GMutex is legal in GLib; create lazily. */
    pthread_mutex_t *p = __atomic_load_n ((pthread_mutex_t **) &m->impl [0], __ATOMIC_ACQUIRE);
    if (!p)
    {
        pthread_mutex_t *n = malloc (sizeof *n), *expected = NULL;
        pthread_mutex_init (n, NULL);
        if (__atomic_compare_exchange_n ((pthread_mutex_t **) &m->impl [0], &expected, n, FALSE,
                                         __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE))
            p = n;
        else
        {
            pthread_mutex_destroy (n);
            free (n);
            p = expected;
        }
    }
    return p;
}
void g_mutex_lock (GMutex *m) { pthread_mutex_lock (mutex_get (m)); }
void g_mutex_unlock (GMutex *m) { pthread_mutex_unlock (mutex_get (m)); }
void g_cond_init (GCond *c) { pthread_cond_t *p = malloc (sizeof *p); pthread_cond_init (p, NULL); c->impl [0] = p; }
void g_cond_clear (GCond *c) { pthread_cond_destroy (c->impl [0]); free (c->impl [0]); c->impl [0] = NULL; }
void g_cond_wait (GCond *c, GMutex *m) { pthread_cond_wait (c->impl [0], mutex_get (m)); }
void g_cond_signal (GCond *c) { pthread_cond_signal (c->impl [0]); }
void g_cond_broadcast (GCond *c) { pthread_cond_broadcast (c->impl [0]); }

typedef struct PoolItem
{
    struct PoolItem *next;
    gpointer data;
}
PoolItem;

struct _GThreadPool
{
    GFunc func;
    gpointer user_data;
    gint n_threads;
    pthread_t *threads;
    pthread_mutex_t mutex;
    pthread_cond_t cond;
    PoolItem *head, *tail;
    gboolean closing;
};

static void *
pool_worker (void *arg)
{
    GThreadPool *pool = arg;

    for (;;)
    {
        PoolItem *item;

        pthread_mutex_lock (&pool->mutex);
        while (!pool->head && !pool->closing)
            pthread_cond_wait (&pool->cond, &pool->mutex);
        item = pool->head;
        if (item)
        {
            pool->head = item->next;
            if (!pool->head)
                pool->tail = NULL;
        }
        pthread_mutex_unlock (&pool->mutex);

        if (!item)
            break;
        pool->func (item->data, pool->user_data);
        free (item);
    }
    return NULL;
}

GThreadPool *
g_thread_pool_new (GFunc func, gpointer user_data, gint max_threads, gboolean exclusive, GError **error)
{
    GThreadPool *pool = g_new0 (GThreadPool, 1);
    gint i;

    (void) exclusive;
    (void) error;
    if (max_threads < 1)
        max_threads = g_get_num_processors ();
    pool->func = func;
    pool->user_data = user_data;
    pool->n_threads = max_threads;
    pool->threads = g_new0 (pthread_t, max_threads);
    pthread_mutex_init (&pool->mutex, NULL);
    pthread_cond_init (&pool->cond, NULL);
    for (i = 0; i < max_threads; i++)
        pthread_create (&pool->threads [i], NULL, pool_worker, pool);
    return pool;
}

gboolean
g_thread_pool_push (GThreadPool *pool, gpointer data, GError **error)
{
    PoolItem *item = calloc (1, sizeof *item);

    (void) error;
    item->data = data;
    pthread_mutex_lock (&pool->mutex);
    if (pool->tail)
        pool->tail->next = item;
    else
        pool->head = item;
    pool->tail = item;
    pthread_cond_signal (&pool->cond);
    pthread_mutex_unlock (&pool->mutex);
    return TRUE;
}

void
g_thread_pool_free (GThreadPool *pool, gboolean immediate, gboolean wait_)
{
    gint i;

    (void) wait_;
    pthread_mutex_lock (&pool->mutex);
    if (immediate)
    {
        while (pool->head)
        {
            PoolItem *n = pool->head->next;
            free (pool->head);
            pool->head = n;
        }
        pool->tail = NULL;
    }
    pool->closing = TRUE;
    pthread_cond_broadcast (&pool->cond);
    pthread_mutex_unlock (&pool->mutex);
    for (i = 0; i < pool->n_threads; i++)
        pthread_join (pool->threads [i], NULL);
    pthread_mutex_destroy (&pool->mutex);
    pthread_cond_destroy (&pool->cond);
    g_free (pool->threads);
    g_free (pool);
}

guint
g_get_num_processors (void)
{
    long n = sysconf (_SC_NPROCESSORS_ONLN);
    const char *env = getenv ("GLIBSHIM_NPROC");
    if (env && atoi (env) > 0)
        return atoi (env);
    return n > 0 ? n : 1;
}

/* ------------------------------------------------------------------ */
/* environment                                                         */

gchar **
g_get_environ (void)
{
    gsize n = 0, i;
    gchar **v;
    while (environ && environ [n])
        n++;
    v = g_new (gchar *, n + 1);
    for (i = 0; i < n; i++)
        v [i] = g_strdup (environ [i]);
    v [n] = NULL;
    return v;
}

const gchar *
g_environ_getenv (gchar **envp, const gchar *variable)
{
    gsize l = strlen (variable);
    if (!envp)
        return NULL;
    for (; *envp; envp++)
        if (!strncmp (*envp, variable, l) && (*envp) [l] == '=')
            return *envp + l + 1;
    return NULL;
}

const gchar *g_getenv (const gchar *variable) { return getenv (variable); }

/* ------------------------------------------------------------------ */
/* hash table: open addressing as GLib >= 2.58 documents it            */

#define HT_MIN_SHIFT 3
#define HT_UNUSED 0u
#define HT_TOMBSTONE 1u
#define HT_IS_REAL(h) ((h) >= 2u)

static const guint ht_prime_mod [] =
{
    1, 2, 3, 7, 13, 31, 61, 127, 251, 509, 1021, 2039, 4093, 8191, 16381, 32749, 65521,
    131071, 262139, 524287, 1048573, 2097143, 4194301, 8388593, 16777213, 33554393,
    67108859, 134217689, 268435399, 536870909, 1073741789, 2147483647
};

struct _GHashTable
{
    gsize size;
    guint mod;
    guint mask;
    guint nnodes;
    guint noccupied;
    gpointer *keys;
    gpointer *values;
    guint *hashes;
    GHashFunc hash_func;
    GEqualFunc key_equal_func;
    GDestroyNotify key_destroy_func;
    GDestroyNotify value_destroy_func;
    gint refs;
};

typedef struct
{
    GHashTable *ht;
    gpointer dummy1;
    gpointer dummy2;
    int position;
    gboolean dummy3;
    gpointer dummy4;
}
RealIter;

guint g_direct_hash (gconstpointer v) { return GPOINTER_TO_UINT (v); }
gboolean g_direct_equal (gconstpointer a, gconstpointer b) { return a == b; }

guint
g_str_hash (gconstpointer v)
{
    const signed char *p;
    guint32 h = 5381;
    for (p = v; *p; p++)
        h = (h << 5) + h + *p;
    return h;
}

gboolean g_str_equal (gconstpointer a, gconstpointer b) { return !strcmp (a, b); }

static void
ht_set_shift (GHashTable *ht, gint shift)
{
    ht->size = (gsize) 1 << shift;
    ht->mod = ht_prime_mod [shift];
    ht->mask = ht->size - 1;
}

static void
ht_set_shift_from_size (GHashTable *ht, gint size)
{
    gint shift = 0;
    while (size)
    {
        size >>= 1;
        shift++;
    }
    ht_set_shift (ht, MAX (shift, HT_MIN_SHIFT));
}

static inline guint
ht_hash_to_index (const GHashTable *ht, guint hash)
{
    return (hash * 11u) % ht->mod;
}

GHashTable *
g_hash_table_new_full (GHashFunc hash_func, GEqualFunc key_equal_func,
                       GDestroyNotify key_destroy_func, GDestroyNotify value_destroy_func)
{
    GHashTable *ht = g_new0 (GHashTable, 1);

    ht->hash_func = hash_func ? hash_func : g_direct_hash;
    ht->key_equal_func = key_equal_func;
    ht->key_destroy_func = key_destroy_func;
    ht->value_destroy_func = value_destroy_func;
    ht->refs = 1;
    ht_set_shift (ht, HT_MIN_SHIFT);
    ht->keys = g_new0 (gpointer, ht->size);
    ht->values = g_new0 (gpointer, ht->size);
    ht->hashes = g_new0 (guint, ht->size);
    return ht;
}

GHashTable *
g_hash_table_new (GHashFunc hash_func, GEqualFunc key_equal_func)
{
    return g_hash_table_new_full (hash_func, key_equal_func, NULL, NULL);
}

static guint
ht_lookup_node (GHashTable *ht, gconstpointer key, guint *hash_return)
{
    guint node_index, node_hash, hash_value;
    guint first_tombstone = 0;
    gboolean have_tombstone = FALSE;
    guint step = 0;

    hash_value = ht->hash_func (key);
    if (!HT_IS_REAL (hash_value))
        hash_value = 2;
    *hash_return = hash_value;

    node_index = ht_hash_to_index (ht, hash_value);
    node_hash = ht->hashes [node_index];

    while (node_hash != HT_UNUSED)
    {
        if (node_hash == hash_value)
        {
            gpointer node_key = ht->keys [node_index];
            if (ht->key_equal_func)
            {
                if (ht->key_equal_func (node_key, key))
                    return node_index;
            }
            else if (node_key == key)
                return node_index;
        }
        else if (node_hash == HT_TOMBSTONE && !have_tombstone)
        {
            first_tombstone = node_index;
            have_tombstone = TRUE;
        }

        step++;
        node_index += step;
        node_index &= ht->mask;
        node_hash = ht->hashes [node_index];
    }

    if (have_tombstone)
        return first_tombstone;
    return node_index;
}

static void
ht_resize (GHashTable *ht)
{
    gsize old_size = ht->size, bm_words;
    guint32 *moved;
    gsize i;

    ht_set_shift_from_size (ht, (gint) (ht->nnodes * 1.333));

    if (ht->size > old_size)
    {
        ht->keys = g_renew (gpointer, ht->keys, ht->size);
        ht->values = g_renew (gpointer, ht->values, ht->size);
        ht->hashes = g_renew (guint, ht->hashes, ht->size);
        memset (&ht->hashes [old_size], 0, (ht->size - old_size) * sizeof (guint));
        bm_words = (ht->size + 31) / 32;
    }
    else
        bm_words = (old_size + 31) / 32;

    moved = g_new0 (guint32, bm_words);

#define BM_GET(i_) ((moved [(i_) / 32] >> ((i_) % 32)) & 1)
#define BM_SET(i_) (moved [(i_) / 32] |= 1u << ((i_) % 32))

    for (i = 0; i < old_size; i++)
    {
        guint node_hash = ht->hashes [i];
        gpointer key, value;

        if (!HT_IS_REAL (node_hash))
        {
            ht->hashes [i] = HT_UNUSED;
            continue;
        }
        if (BM_GET (i))
            continue;

        ht->hashes [i] = HT_UNUSED;
        key = ht->keys [i];
        value = ht->values [i];
        ht->keys [i] = NULL;
        ht->values [i] = NULL;

        for (;;)
        {
            guint idx, replaced_hash, step = 0;
            gpointer k2, v2;

            idx = ht_hash_to_index (ht, node_hash);
            while (BM_GET (idx))
            {
                step++;
                idx += step;
                idx &= ht->mask;
            }
            BM_SET (idx);

            replaced_hash = ht->hashes [idx];
            ht->hashes [idx] = node_hash;
            if (!HT_IS_REAL (replaced_hash))
            {
                ht->keys [idx] = key;
                ht->values [idx] = value;
                break;
            }

            node_hash = replaced_hash;
            k2 = ht->keys [idx];
            v2 = ht->values [idx];
            ht->keys [idx] = key;
            ht->values [idx] = value;
            key = k2;
            value = v2;
        }
    }

#undef BM_GET
#undef BM_SET

    g_free (moved);

    if (ht->size < old_size)
    {
        ht->keys = g_renew (gpointer, ht->keys, ht->size);
        ht->values = g_renew (gpointer, ht->values, ht->size);
        ht->hashes = g_renew (guint, ht->hashes, ht->size);
    }

    ht->noccupied = ht->nnodes;
}

static void
ht_maybe_resize (GHashTable *ht)
{
    gsize noccupied = ht->noccupied, size = ht->size;

    if ((size > (gsize) ht->nnodes * 4 && size > (1u << HT_MIN_SHIFT))
        || (size <= noccupied + (noccupied / 16)))
        ht_resize (ht);
}

static gboolean
ht_insert_internal (GHashTable *ht, gpointer key, gpointer value, gboolean keep_new_key)
{
    guint key_hash, node_index, old_hash;
    gboolean already_exists;

    node_index = ht_lookup_node (ht, key, &key_hash);
    old_hash = ht->hashes [node_index];
    already_exists = HT_IS_REAL (old_hash);

    if (already_exists)
    {
        gpointer old_value = ht->values [node_index];
        gpointer key_to_free;

        if (keep_new_key)
        {
            key_to_free = ht->keys [node_index];
            ht->keys [node_index] = key;
        }
        else
            key_to_free = key;

        ht->values [node_index] = value;
        if (ht->key_destroy_func)
            ht->key_destroy_func (key_to_free);
        if (ht->value_destroy_func)
            ht->value_destroy_func (old_value);
    }
    else
    {
        ht->hashes [node_index] = key_hash;
        ht->keys [node_index] = key;
        ht->values [node_index] = value;
        ht->nnodes++;
        if (old_hash == HT_UNUSED)
        {
            ht->noccupied++;
            ht_maybe_resize (ht);
        }
    }

    return !already_exists;
}

gboolean g_hash_table_insert (GHashTable *ht, gpointer key, gpointer value) { return ht_insert_internal (ht, key, value, FALSE); }
gboolean g_hash_table_replace (GHashTable *ht, gpointer key, gpointer value) { return ht_insert_internal (ht, key, value, TRUE); }

static void
ht_remove_node (GHashTable *ht, guint i, gboolean notify)
{
    gpointer key = ht->keys [i], value = ht->values [i];

    ht->hashes [i] = HT_TOMBSTONE;
    ht->keys [i] = NULL;
    ht->values [i] = NULL;
    ht->nnodes--;
    if (notify && ht->key_destroy_func)
        ht->key_destroy_func (key);
    if (notify && ht->value_destroy_func)
        ht->value_destroy_func (value);
}

gboolean
g_hash_table_remove (GHashTable *ht, gconstpointer key)
{
    guint hash, i = ht_lookup_node (ht, key, &hash);

    if (!HT_IS_REAL (ht->hashes [i]))
        return FALSE;
    ht_remove_node (ht, i, TRUE);
    ht_maybe_resize (ht);
    return TRUE;
}

void
g_hash_table_remove_all (GHashTable *ht)
{
    gsize i;

    for (i = 0; i < ht->size; i++)
    {
        if (HT_IS_REAL (ht->hashes [i]))
        {
            if (ht->key_destroy_func)
                ht->key_destroy_func (ht->keys [i]);
            if (ht->value_destroy_func)
                ht->value_destroy_func (ht->values [i]);
        }
    }
    g_free (ht->keys);
    g_free (ht->values);
    g_free (ht->hashes);
    ht->nnodes = ht->noccupied = 0;
    ht_set_shift (ht, HT_MIN_SHIFT);
    ht->keys = g_new0 (gpointer, ht->size);
    ht->values = g_new0 (gpointer, ht->size);
    ht->hashes = g_new0 (guint, ht->size);
}

void
g_hash_table_unref (GHashTable *ht)
{
    if (!ht)
        return;
    if (__atomic_fetch_sub (&ht->refs, 1, __ATOMIC_SEQ_CST) != 1)
        return;
    g_hash_table_remove_all (ht);
    g_free (ht->keys);
    g_free (ht->values);
    g_free (ht->hashes);
    g_free (ht);
}

void g_hash_table_destroy (GHashTable *ht) { g_hash_table_unref (ht); }

gpointer
g_hash_table_lookup (GHashTable *ht, gconstpointer key)
{
    guint hash, i = ht_lookup_node (ht, key, &hash);
    return HT_IS_REAL (ht->hashes [i]) ? ht->values [i] : NULL;
}

gboolean
g_hash_table_contains (GHashTable *ht, gconstpointer key)
{
    guint hash, i = ht_lookup_node (ht, key, &hash);
    return HT_IS_REAL (ht->hashes [i]);
}

guint g_hash_table_size (GHashTable *ht) { return ht->nnodes; }

void
g_hash_table_iter_init (GHashTableIter *iter, GHashTable *ht)
{
    RealIter *ri = (RealIter *) iter;
    ri->ht = ht;
    ri->position = -1;
}

gboolean
g_hash_table_iter_next (GHashTableIter *iter, gpointer *key, gpointer *value)
{
    RealIter *ri = (RealIter *) iter;
    gint position = ri->position;

    do
    {
        position++;
        if (position >= (gint) ri->ht->size)
        {
            ri->position = position;
            return FALSE;
        }
    }
    while (!HT_IS_REAL (ri->ht->hashes [position]));

    if (key)
        *key = ri->ht->keys [position];
    if (value)
        *value = ri->ht->values [position];
    ri->position = position;
    return TRUE;
}

/* ------------------------------------------------------------------ */
/* array                                                               */

typedef struct
{
    gchar *data;
    guint len;
    guint alloc;
    guint elt_size;
    guint zero_terminated : 1;
    guint clear : 1;
    gint refs;
}
RealArray;

static void
array_maybe_expand (RealArray *a, guint extra)
{
    guint want = a->len + extra + a->zero_terminated;
    if (want > a->alloc)
    {
        guint n = a->alloc ? a->alloc : 16;
        while (n < want)
            n *= 2;
        a->data = g_realloc (a->data, (gsize) n * a->elt_size);
        if (a->clear)
            memset (a->data + (gsize) a->alloc * a->elt_size, 0, (gsize) (n - a->alloc) * a->elt_size);
        a->alloc = n;
    }
}

GArray *
g_array_sized_new (gboolean zero_terminated, gboolean clear_, guint element_size, guint reserved)
{
    RealArray *a = g_new0 (RealArray, 1);
    a->elt_size = element_size;
    a->zero_terminated = !!zero_terminated;
    a->clear = !!clear_;
    a->refs = 1;
    if (reserved || zero_terminated)
    {
        array_maybe_expand (a, reserved);
        if (zero_terminated)
            memset (a->data, 0, element_size);
    }
    return (GArray *) a;
}

GArray *
g_array_new (gboolean zero_terminated, gboolean clear_, guint element_size)
{
    return g_array_sized_new (zero_terminated, clear_, element_size, 0);
}

gchar *
g_array_free (GArray *array, gboolean free_segment)
{
    RealArray *a = (RealArray *) array;
    gchar *d;
    if (!a)
        return NULL;
    d = a->data;
    if (free_segment)
    {
        g_free (d);
        d = NULL;
    }
    g_free (a);
    return d;
}

void
g_array_unref (GArray *array)
{
    RealArray *a = (RealArray *) array;
    if (__atomic_fetch_sub (&a->refs, 1, __ATOMIC_SEQ_CST) == 1)
        g_array_free (array, TRUE);
}

GArray *
g_array_append_vals (GArray *array, gconstpointer data, guint len)
{
    RealArray *a = (RealArray *) array;
    if (!len)
        return array;
    array_maybe_expand (a, len);
    memcpy (a->data + (gsize) a->len * a->elt_size, data, (gsize) len * a->elt_size);
    a->len += len;
    if (a->zero_terminated)
        memset (a->data + (gsize) a->len * a->elt_size, 0, a->elt_size);
    return array;
}

GArray *
g_array_set_size (GArray *array, guint length)
{
    RealArray *a = (RealArray *) array;
    if (length > a->len)
    {
        array_maybe_expand (a, length - a->len);
        if (a->clear)
            memset (a->data + (gsize) a->len * a->elt_size, 0, (gsize) (length - a->len) * a->elt_size);
    }
    a->len = length;
    if (a->zero_terminated && a->data)
        memset (a->data + (gsize) a->len * a->elt_size, 0, a->elt_size);
    return array;
}

GArray *
g_array_remove_index (GArray *array, guint index_)
{
    RealArray *a = (RealArray *) array;
    if (index_ != a->len - 1)
        memmove (a->data + (gsize) index_ * a->elt_size, a->data + (gsize) (index_ + 1) * a->elt_size,
                 (gsize) (a->len - index_ - 1) * a->elt_size);
    a->len--;
    if (a->zero_terminated)
        memset (a->data + (gsize) a->len * a->elt_size, 0, a->elt_size);
    return array;
}

/* ------------------------------------------------------------------ */
/* time                                                                */

gint64
g_get_monotonic_time (void)
{
    struct timespec ts;
    clock_gettime (CLOCK_MONOTONIC, &ts);
    return (gint64) ts.tv_sec * 1000000 + ts.tv_nsec / 1000;
}

gint64
g_get_real_time (void)
{
    struct timespec ts;
    clock_gettime (CLOCK_REALTIME, &ts);
    return (gint64) ts.tv_sec * 1000000 + ts.tv_nsec / 1000;
}

void g_usleep (gulong microseconds) { usleep (microseconds); }
